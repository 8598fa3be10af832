#!/bin/bash
# phase-3 tiling / routing-lane sensitivity of the fused kernel (env overrides MOE_FUSED_BN / _SPLIT / _LANES); run under gpurun
run() { echo -n "$1 | "; env $1 python tools/sweep_fused.py $2 $3 2>&1 | grep fused | sed 's/| split.*//'; }
for cfg in "X=1" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=2" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=4" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=2" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=1" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=1"; do run "$cfg" 640 2048; done
for cfg in "X=1" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=4" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=8" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=2" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=4" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=1"; do run "$cfg" 1280 512; done
for cfg in "X=1" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=8" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=4" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=8" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=2"; do run "$cfg" 1280 128; done
for cfg in "X=1" "MOE_FUSED_BN=80 MOE_FUSED_SPLIT=1" "MOE_FUSED_BN=160 MOE_FUSED_SPLIT=2" "MOE_FUSED_LANES=4" "MOE_FUSED_LANES=16" "MOE_FUSED_LANES=32"; do run "$cfg" 320 8192; done
