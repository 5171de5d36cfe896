#!/bin/bash
# DL-SCL env-knob A/B on the GPU box: each entry of the list is a set of env assignments tried with scripts/dl_leg_probe.py
mkdir -p gpurun_out
for v in "X=0" "PB200_DL_WPC=10" "PB200_DL_WPC=8" "PB200_DL_WPC=6" "PB200_DL_WPC=5"; do
  echo "== $v"
  env $v timeout 300 python scripts/dl_leg_probe.py 2>&1 | grep -E " 4 4.0 beta| 8 4.0 beta| 4 5.0 beta| 8 5.0 beta"
done
