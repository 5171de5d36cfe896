#!/bin/bash
# DL-SCL retry kernel A/B on the GPU box: binned (default) vs frame-per-group (PB200_DL_BINNED=0); tests first.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "dl or DL or retr or sweep or published or mirror or generic" > gpurun_out/dl_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/dl_tests.log
for mode in 1 0; do
  echo "== PB200_DL_BINNED=$mode"
  PB200_DL_BINNED=$mode timeout 300 python scripts/dl_probe.py 2097152 2>&1 | tail -12
done
