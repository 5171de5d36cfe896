#!/bin/bash
# Turn the captures of `scripts/ncu_round.sh <tag>` (gpurun_out/<tag>_*.ncu-rep) into the tracked summaries under profiles/
# and refresh profiles/r02_constants.json (run in the build container: ncu reads the reports without a GPU).
#   bash scripts/ncu_publish.sh r02_v17
tag=${1:?tag}
F=1048576
pub() {  # <rep suffix> <profiles name> <constants key>
  rep=gpurun_out/${tag}_$1.ncu-rep
  [ -f "$rep" ] || { echo "missing $rep"; return; }
  python scripts/ncu_summary.py "$rep" --json "$3" $F > profiles/${tag}_$2_metrics.txt
  ncu -i "$rep" --page source --csv --print-source cuda,sass > /tmp/${tag}_$1_src.csv 2>/dev/null
  python scripts/ncu_lines.py /tmp/${tag}_$1_src.csv 60 > profiles/${tag}_$2_lines.txt
  python scripts/opmix.py /tmp/${tag}_$1_src.csv > profiles/${tag}_$2_opmix.txt
}
pub decode decode_kernel decode_kernel_M4
pub decode_M8 decode_kernel_M8 decode_kernel_M8
pub sweep sweep_kernel sweep_kernel_M4
pub tracesweep trace_sweep_kernel trace_sweep_kernel_M4
pub retry dl_bin_kernel dl_bin_kernel_M4_4dB_beta
cp gpurun_out/${tag}_launches.csv profiles/${tag}_launches.csv
ls -la profiles/${tag}_*
python scripts/launch_shares.py profiles/${tag}_launches.csv > profiles/${tag}_launch_shares.txt
