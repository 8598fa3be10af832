#!/bin/bash
# A/B of one environment switch of the fused kernel: tools/ab_env.sh VAR "v1 v2 ..." d T [d T ...]
var=$1; vals=$2; shift 2
for v in $vals $vals; do
  echo "== $var=$v"
  env $var=$v FUSED_ONLY=1 python tools/sweep_fused.py "$@" 2>&1 | grep fused
done
