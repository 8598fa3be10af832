#!/bin/bash
# One-box sweep of the fused kernel's environment knobs on the four UNet-batch-2 layer shapes (us per layer call).
shapes="320 8192 640 2048 1280 512 1280 128"
run() { echo "== $*"; env "$@" FUSED_ONLY=1 python tools/sweep_fused.py $shapes 2>&1 | grep fused | awk '{printf "%s %s %s  ", $1, $2, $5} END {print ""}'; }
run X=default
for v in 2 3 4 6; do run MOE_FUSED_PUB=$v; done
for v in 4 8 16 32; do run MOE_FUSED_LANES=$v; done
for v in 2 4 8 16; do run MOE_FUSED_ROUTE_WARPS=$v; done
for v in 1 2 3 4; do run MOE_FUSED_SPLIT=$v; done
for v in 80 160; do run MOE_FUSED_BN=$v; done
run MOE_FUSED_KS=1
run MOE_FUSED_PREFETCH=0
run MOE_FUSED_REV3=1
run X=default
