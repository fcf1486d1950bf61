#!/bin/bash
# First GPU session of round 2 (1 GPU, ~25 min box time).  Round 1 ran out of GPU budget before these could
# be measured; everything here was prepared and CPU-checked in round 1.
#   gpurun --timeout 1700 -- 'bash tools/gpu_session_r2_first.sh'
#  1. gram2_dmma_kernel (opt-in Gram tile kernel, csrc/gram2.cu): gated parity test under a short timeout
#     (a barrier-protocol bug would hang, the model test cannot see those), then the A/B against
#     gram_dmma_kernel with its ablation bits (32 no triangular blocks, 64 no K-split, 128 no ragged
#     specialisation).
#  2. ncu evidence for the HBM-bound kernels of the Gram-Schmidt loop.
#  3. the full -m gpu suite and the contract bench.
set -u
mkdir -p gpurun_out
export PMC_TEST_GRAM2=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method thread -k gram2 \
    > gpurun_out/r02_pytest_gram2.log 2>&1
echo "pytest gram2 exit $?" >> gpurun_out/r02_pytest_gram2.log
tail -n 4 gpurun_out/r02_pytest_gram2.log
if grep -q "pytest gram2 exit 0" gpurun_out/r02_pytest_gram2.log; then
    timeout 900 python tools/bench_kernels.py --shapes 2000000x500,4000000x1000,4000000x256 --no-lincomb \
        --opts 0,16,48,80,144 > gpurun_out/r02_kernels_gram2_ab.jsonl 2>&1
    grep -E '^\{' gpurun_out/r02_kernels_gram2_ab.jsonl | cut -c1-400
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram2_dmma -c 2 \
        -o gpurun_out/r02_prof_gram2 -f python tools/bench_kernels.py --shapes 2000000x500 --no-lincomb --opts 16 --reps 1 \
        > gpurun_out/r02_ncu_gram2.log 2>&1
fi
unset PMC_TEST_GRAM2
# fused Gram-Schmidt sweep (csrc/mgs.cu): gated parity test under a short timeout (the kernel's waits give up
# after 5 s on their own), then config 3 pair-by-pair vs swept
PMC_TEST_SWEEP=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --timeout-method thread \
    -k "mgs_sweep or device_cg" > gpurun_out/r02_pytest_sweep.log 2>&1
echo "pytest sweep exit $?" >> gpurun_out/r02_pytest_sweep.log
tail -n 4 gpurun_out/r02_pytest_sweep.log
if grep -q "pytest sweep exit 0" gpurun_out/r02_pytest_sweep.log; then
    timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/r02_gs_10Mx256_pairs.log 2>&1
    timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 --sweep > gpurun_out/r02_gs_10Mx256_sweep.log 2>&1
    grep -hE '^\{' gpurun_out/r02_gs_10Mx256_pairs.log gpurun_out/r02_gs_10Mx256_sweep.log | cut -c1-500
fi
timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 --variant cgs2 > gpurun_out/r02_gs_10Mx256_cgs2.log 2>&1
grep -hE '^\{' gpurun_out/r02_gs_10Mx256_cgs2.log | cut -c1-500
timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 > gpurun_out/r02_gs_10Mx256_bcgs2.log 2>&1
timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 --panel 128,16 > gpurun_out/r02_gs_10Mx256_bcgs2_128_16.log 2>&1
timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 --panel 32,8 > gpurun_out/r02_gs_10Mx256_bcgs2_32_8.log 2>&1
grep -hE '^\{' gpurun_out/r02_gs_10Mx256_bcgs2.log | cut -c1-500
grep -hE '^\{' gpurun_out/r02_gs_10Mx256_bcgs2_128_16.log | cut -c1-500
grep -hE '^\{' gpurun_out/r02_gs_10Mx256_bcgs2_32_8.log | cut -c1-500
timeout 600 python tools/bench_pod_qr.py > gpurun_out/r02_pod_qr_4Mx500.log 2>&1
grep -hE '^\{' gpurun_out/r02_pod_qr_4Mx500.log | cut -c1-600
timeout 300 python tools/bench_cholqr.py --dofs 10000000 --k 256 > gpurun_out/r02_cholqr_10Mx256.log 2>&1
grep -hE '^\{' gpurun_out/r02_cholqr_10Mx256.log | cut -c1-400
# HBM-bound kernels: 10M-DOF vectors, 12 columns -> a few hundred launches; capture a handful of each kind
timeout 300 python tools/bench_gs.py --dofs 10000000 --k 12 > gpurun_out/r02_gs_10Mx12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'reduce_cols|axpy_dot|elementwise' \
    --launch-skip 40 -c 12 -o gpurun_out/r02_prof_blas1 -f python tools/bench_gs.py --dofs 10000000 --k 12 \
    > gpurun_out/r02_ncu_blas1.log 2>&1
ncu -i gpurun_out/r02_prof_blas1.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    > gpurun_out/r02_blas1_raw.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method thread > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r02_pytest_gpu.log
tail -n 3 gpurun_out/r02_pytest_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_10Mx1000.log 2>&1
grep -E '^\{' gpurun_out/r02_bench_10Mx1000.log | cut -c1-600
