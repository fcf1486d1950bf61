#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/spmm_ab.log
for v in 0 1; do PMC_CSR_VARIANT=$v timeout 600 python tools/bench_spmm.py >> gpurun_out/spmm_ab.log 2>&1; done
timeout 600 python tools/bench_spmm.py --ks 64 --reps 2 > gpurun_out/spmm_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_apply" -s 1 -c 1 -o gpurun_out/prof_spmm_tiled python tools/bench_spmm.py --ks 64 --reps 2 > gpurun_out/ncu_spmm.log 2>&1
PMC_CSR_VARIANT=1 timeout 600 python tools/bench_spmm.py --ks 64 --reps 2 > gpurun_out/spmm_plain1.log 2>&1 &&
PMC_CSR_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"csr_apply" -s 1 -c 1 -o gpurun_out/prof_spmm_simple python tools/bench_spmm.py --ks 64 --reps 2 > gpurun_out/ncu_spmm1.log 2>&1
timeout 900 python tools/bench_project.py --nx 2500 --ny 2500 --k 2000 --reps 2 > gpurun_out/project_6Mx2000_v2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread -k "csr or projection or vecops" > gpurun_out/pytest_s10.log 2>&1
grep -E '^\{' gpurun_out/spmm_ab.log | cut -c1-300; grep -E '^\{|Error' gpurun_out/project_6Mx2000_v2.log | cut -c1-700; tail -n 2 gpurun_out/pytest_s10.log
