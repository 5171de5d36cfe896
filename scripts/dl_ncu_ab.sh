#!/bin/bash
# retry kernel duration + instruction count, binned vs frame-per-group (ncu metrics pass, 1 Mi frames, M=4 @4 dB and @5 dB)
mkdir -p gpurun_out
for mode in 1 0; do
 for snr in 4.0 5.0; do
  PB200_DL_BINNED=$mode python scripts/prof_decode.py dl 4 $snr > gpurun_out/dlab_plain_${mode}_$snr.log 2>&1 &&
  PB200_DL_BINNED=$mode ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:dl_ -s 3 -c 2 --csv --log-file gpurun_out/dlab_${mode}_$snr.csv python scripts/prof_decode.py dl 4 $snr > /dev/null 2>&1
  echo "mode=$mode snr=$snr"; tail -2 gpurun_out/dlab_plain_${mode}_$snr.log; grep -v "^==" gpurun_out/dlab_${mode}_$snr.csv | cut -d, -f5,13- | tail -10
 done
done
