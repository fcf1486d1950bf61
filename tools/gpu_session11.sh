#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread -k "complex or csr or projection or vecops or chol" > gpurun_out/pytest_s11.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s11.log
timeout 600 python tools/bench_spmm.py > gpurun_out/spmm_warm.log 2>&1
timeout 900 python tools/bench_project.py --nx 2500 --ny 2500 --k 2000 --reps 2 > gpurun_out/project_6Mx2000_v3.log 2>&1
timeout 900 python tools/bench_project.py --nx 2000 --ny 1000 --k 1000 --check > gpurun_out/project_2Mx1000_v3.log 2>&1
tail -n 2 gpurun_out/pytest_s11.log; grep -E '^\{' gpurun_out/spmm_warm.log | cut -c1-200; for f in gpurun_out/project_6Mx2000_v3.log gpurun_out/project_2Mx1000_v3.log; do grep -E '^\{|Error' $f | cut -c1-800; done
