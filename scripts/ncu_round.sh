#!/bin/bash
# ncu evidence of the current build (one GPU): --set full captures of the four hot kernels on 1 Mi frames, and the launch
# list of the bench command.  usage: bash scripts/ncu_round.sh <tag>   -> gpurun_out/<tag>_{decode,decode_M8,sweep,tracesweep,retry}.ncu-rep
tag=${1:-r02_final}
mkdir -p gpurun_out
P="python scripts/prof_decode.py"
NCU="ncu --set full --clock-control none --import-source on -s 3 -c 1 -f"
$P decode 4 5.0 > gpurun_out/${tag}_plain.log 2>&1 && $NCU -k regex:decode_kernel -o gpurun_out/${tag}_decode $P decode 4 5.0 > gpurun_out/${tag}_ncu_decode.log 2>&1
$P decode 8 5.0 >> gpurun_out/${tag}_plain.log 2>&1 && $NCU -k regex:decode_kernel -o gpurun_out/${tag}_decode_M8 $P decode 8 5.0 > gpurun_out/${tag}_ncu_decode_M8.log 2>&1
$P sweep 4 5.0 >> gpurun_out/${tag}_plain.log 2>&1 && $NCU -k regex:sweep_kernel -o gpurun_out/${tag}_sweep $P sweep 4 5.0 > gpurun_out/${tag}_ncu_sweep.log 2>&1
PB200_DL_REPLAY=0 $P dl 4 4.0 >> gpurun_out/${tag}_plain.log 2>&1 && PB200_DL_REPLAY=0 $NCU -k regex:sweep_kernel -o gpurun_out/${tag}_tracesweep $P dl 4 4.0 > gpurun_out/${tag}_ncu_tracesweep.log 2>&1
$P dl 4 4.0 >> gpurun_out/${tag}_plain.log 2>&1 && $NCU -k regex:dl_bin_kernel -o gpurun_out/${tag}_retry $P dl 4 4.0 > gpurun_out/${tag}_ncu_retry.log 2>&1
B="python bench.py --steps 2 --warmup 3 --frames 1048576 --e2e-frames 262144 --cpu-sample 60000"
$B > gpurun_out/${tag}_bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
cat gpurun_out/${tag}_plain.log | tail -20
ls -la gpurun_out/${tag}_*
