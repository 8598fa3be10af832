#!/bin/bash
# A/B of the phase-1 schedules of the fused kernel (default / resident-x with 2-k-block slots / resident-x with whole-tile slots)
for v in 0 1 2 0 2; do
  echo "== MOE_FUSED_ARES=$v"
  MOE_FUSED_ARES=$v FUSED_ONLY=1 python tools/sweep_fused.py 320 8192 320 65536 2>&1 | grep fused
done
