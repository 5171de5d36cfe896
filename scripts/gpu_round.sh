#!/bin/bash
# One GPU-box pass: the -m gpu tests, smoke, the bench line (N=1) -- logs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "bench rc=$?"
python scripts/quick_show.py gpurun_out/bench_now.json
