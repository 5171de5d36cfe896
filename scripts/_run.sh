timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -2 gpurun_out/gpu_tests.log
for i in 1 2; do python scripts/dl_stats.py 4 4.0 1048576 | cut -c60-400; done
python scripts/dl_stats.py 8 4.0 1048576 | cut -c1-400
python scripts/dl_stats.py 4 5.0 2097152 | cut -c60-400
PB200_DL_BINNED=0 python scripts/dl_stats.py 4 4.0 1048576 | cut -c1-100
PB200_DL_BINNED=0 python scripts/dl_stats.py 8 4.0 1048576 | cut -c1-100
