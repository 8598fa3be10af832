#!/bin/bash
# fused kernel per layer shape under environment switches (A/B runs); run under gpurun
run() { echo -n "$1 | "; env $1 python tools/sweep_fused.py 320 8192 640 2048 1280 512 1280 128 2>&1 | grep fused | sed 's/| split.*//' | sed 's/es=20: fused//' | tr '\n' ' '; echo; }
for cfg in "$@"; do run "$cfg"; done
