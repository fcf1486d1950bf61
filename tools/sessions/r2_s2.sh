#!/bin/bash
# Round 2, GPU session 2 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/sessions/r2_s2.sh'
# gram2 with the weight applied in registers vs in shared memory vs gram_dmma_kernel; the new CSR kernels; the
# rewritten bench (pyMOR's pod, kernel-time roofline, parity block); DRAM bytes of the MGS loop as it runs
# (ncu --cache-control none: the vector being orthogonalised stays in L2 between launches).
set -u
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider"
timeout 900 $PT tests/test_gpu_parity.py -m gpu -x -k "gram2 or csr or kernel_timing or sweep or device_cg" > $O/r02_s2_pytest_new.log 2>&1
echo "pytest new exit $?" >> $O/r02_s2_pytest_new.log; tail -n 4 $O/r02_s2_pytest_new.log
timeout 900 python tools/bench_kernels.py --shapes 2000000x500,10000000x1000,1250000x1000 --no-lincomb --opts 0,16,272 \
    > $O/r02_s2_kernels_gram2_wreg.jsonl 2>&1
grep -E '^\{' $O/r02_s2_kernels_gram2_wreg.jsonl | cut -c1-420
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram2_dmma --launch-skip 2 -c 1 \
    -o $O/r02_prof_gram2_wreg_10Mx1000 -f python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts 16 --reps 1 \
    > $O/r02_s2_ncu_gram2.log 2>&1
ncu -i $O/r02_prof_gram2_wreg_10Mx1000.ncu-rep --page details > $O/r02_ncu_gram2_wreg_10Mx1000_details.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 > $O/r02_s2_bench_1gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s2_bench_1gpu.log
grep -E '^\{' $O/r02_s2_bench_1gpu.log | cut -c1-3000; tail -n 3 $O/r02_s2_bench_1gpu.log | cut -c1-600
PMC_GRAM_OPTS=16 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > $O/r02_s2_bench_1gpu_gram2.log 2>&1
grep -E '^\{' $O/r02_s2_bench_1gpu_gram2.log | cut -c1-1500
# DRAM bytes of the Gram-Schmidt loop as it runs: no cache flush between kernels
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    -k regex:'reduce_cols|axpy_dot|elementwise|mgs_sweep' --launch-skip 30 -c 120 --csv --log-file $O/r02_blas1_live_raw.csv \
    python tools/bench_gs.py --dofs 10000000 --k 12 > $O/r02_s2_ncu_blas1_live.log 2>&1
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    -k regex:'mgs_sweep' -c 20 --csv --log-file $O/r02_sweep_live_raw.csv \
    python tools/bench_gs.py --dofs 10000000 --k 12 --sweep > $O/r02_s2_ncu_sweep_live.log 2>&1
timeout 2400 $PT tests -m gpu -x > $O/r02_s2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s2_pytest_gpu.log
tail -n 5 $O/r02_s2_pytest_gpu.log
