#!/bin/bash
# same-box A/B of the committed build (MOE_LIB_VARIANT=base: build HEAD with that variant name first) against the working tree
for v in base "" base ""; do
  echo "== lib variant '${v:-worktree}'"
  MOE_LIB_VARIANT=$v FUSED_ONLY=1 python tools/sweep_fused.py "$@" 2>&1 | grep fused
done
