#!/bin/bash
# DL-SCL env-knob A/B on the GPU box: each line of VARIANTS is a set of env assignments tried with scripts/dl_leg_probe.py
mkdir -p gpurun_out
for v in "X=0" "PB200_L2_HIST=1" "X=1" "PB200_L2_HIST=1 PB200_DL_WPC=32" ; do
  echo "== $v"
  env $v timeout 300 python scripts/dl_leg_probe.py 2>&1 | tail -8
done
