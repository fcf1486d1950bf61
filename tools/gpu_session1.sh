#!/bin/bash
# First GPU session: parity tests, smoke, small + full bench.  Every step under its own timeout.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --n 2000000 --k 500 --cpu-sample 100000 > gpurun_out/bench_2Mx500.log 2>&1
echo "bench small exit $?" >> gpurun_out/bench_2Mx500.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_10Mx1000.log 2>&1
echo "bench full exit $?" >> gpurun_out/bench_10Mx1000.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_2Mx500.log gpurun_out/bench_10Mx1000.log
