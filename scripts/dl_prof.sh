#!/bin/bash
# binned retry kernel: timing in both modes, then one --set full capture of dl_bin_kernel (M=4, 4 dB, 1 Mi frames)
mkdir -p gpurun_out
for mode in 1 0; do echo "mode=$mode"; PB200_DL_BINNED=$mode python scripts/prof_decode.py dl 4 4.0 2>&1 | tail -3; PB200_DL_BINNED=$mode python scripts/prof_decode.py dl 4 5.0 2>&1 | tail -2; PB200_DL_BINNED=$mode python scripts/prof_decode.py dl 8 4.0 2>&1 | tail -2; done
python scripts/prof_decode.py dl 4 4.0 > gpurun_out/dlbin_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dl_bin_kernel -s 3 -c 1 -f -o gpurun_out/r02_dlbin_a python scripts/prof_decode.py dl 4 4.0 > gpurun_out/dlbin_ncu.log 2>&1
ls -la gpurun_out/r02_dlbin_a.ncu-rep
