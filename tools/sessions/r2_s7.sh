#!/bin/bash
# Round 2, GPU session 7 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/sessions/r2_s7.sh'
# what the driver runs at round end, on the current tree: smoke(), the -m gpu suite, the bench with the driver's
# arguments; plus config 2 (2M x 500) and the vectorised amax under ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider"
timeout 300 python __graft_entry__.py smoke > $O/r02_s7_smoke.log 2>&1; echo "smoke exit $?" >> $O/r02_s7_smoke.log; tail -n 2 $O/r02_s7_smoke.log | cut -c1-200
timeout 2400 $PT tests -m gpu -x > $O/r02_s7_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s7_pytest_gpu.log
tail -n 5 $O/r02_s7_pytest_gpu.log | cut -c1-300
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r02_s7_bench_1gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s7_bench_1gpu.log
grep -E '^\{' $O/r02_s7_bench_1gpu.log | cut -c1-700; tail -n 6 $O/r02_s7_bench_1gpu.log | cut -c1-200
timeout 600 python bench.py --dofs 2000000 --k 500 --steps 5 --warmup 3 --no-csr-pod > $O/r02_s7_bench_2Mx500.log 2>&1
grep -E '^\{' $O/r02_s7_bench_2Mx500.log | cut -c1-900
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'amax_kernel' -c 2 --csv \
    --log-file $O/r02_s7_amax.csv python tools/bench_misc_kernels.py > /dev/null 2>&1
grep -E '^"[0-9]' $O/r02_s7_amax.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ' '; echo
