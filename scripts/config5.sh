#!/bin/bash
# BASELINE config 5: DL-SCL M=8, 8 retries, beta_M8, 1e9 frames at one SNR point, sharded over the visible GPUs.
set -e
N=${1:-8}; FRAMES=${2:-1000000000}
mkdir -p gpurun_out/c5
python -c "import numpy as np; g=np.load('tests/golden/scl_p128.npz'); np.save('gpurun_out/c5/beta_M8.npy', g['beta_M8'])"
ls -la gpurun_out/c5/beta_M8.npy
time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  -m dl_scl_polar.eval.run_fer_sweep --M 8 --frames $FRAMES --snr_lo 5.0 --snr_step 0 --retries 8 \
  --beta gpurun_out/c5/beta_M8.npy --out_dir gpurun_out/c5 --plot_dir gpurun_out/c5 > gpurun_out/c5/log.txt 2>&1 || { tail -20 gpurun_out/c5/log.txt; exit 1; }
grep -E "SNR=|Saved" gpurun_out/c5/log.txt
cat gpurun_out/c5/fer_M8.csv
