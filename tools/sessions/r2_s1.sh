#!/bin/bash
# Round 2, GPU session 1 (1 GPU):  gpurun --timeout 2400 -- 'bash tools/sessions/r2_s1.sh'
# pyMOR from baseline/_ref on the box; first runs of gram2_dmma_kernel / mgs_sweep_kernel / device CG (gated tests
# under short timeouts), A/B + ncu captures at the headline shape, ncu of the HBM-bound kernels, full -m gpu suite.
set -u
mkdir -p gpurun_out
O=gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; python -c "import pymor_b200.interface as i, pymor; print('pymor', pymor.__version__, pymor.__file__)"; } > $O/r02_s1_info.txt 2>&1
cat $O/r02_s1_info.txt
PT="python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider"

export PMC_TEST_GRAM2=1
timeout 700 $PT tests/test_gpu_parity.py -m gpu -k gram2 > $O/r02_pytest_gram2.log 2>&1; echo "pytest gram2 exit $?" >> $O/r02_pytest_gram2.log
tail -n 4 $O/r02_pytest_gram2.log
if grep -q "pytest gram2 exit 0" $O/r02_pytest_gram2.log; then
    timeout 900 python tools/bench_kernels.py --shapes 2000000x500,10000000x1000 --no-lincomb --opts 0,16,48,80,144 \
        > $O/r02_kernels_gram2_ab.jsonl 2>&1
    grep -E '^\{' $O/r02_kernels_gram2_ab.jsonl | cut -c1-420
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram2_dmma --launch-skip 2 -c 1 \
        -o $O/r02_prof_gram2_10Mx1000 -f python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts 16 --reps 1 \
        > $O/r02_ncu_gram2.log 2>&1
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_dmma --launch-skip 2 -c 1 \
    -o $O/r02_prof_gram_10Mx1000 -f python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts 0 --reps 1 \
    > $O/r02_ncu_gram.log 2>&1
unset PMC_TEST_GRAM2

PMC_TEST_SWEEP=1 timeout 700 $PT tests/test_gpu_parity.py -m gpu -k "mgs_sweep or device_cg" > $O/r02_pytest_sweep.log 2>&1
echo "pytest sweep exit $?" >> $O/r02_pytest_sweep.log
tail -n 4 $O/r02_pytest_sweep.log
gs() { tag=$1; shift; timeout 300 python tools/bench_gs.py --dofs 10000000 --k 256 "$@" > $O/r02_gs_10Mx256_$tag.log 2>&1; grep -hE '^\{' $O/r02_gs_10Mx256_$tag.log | cut -c1-500; }
gs pairs
if grep -q "pytest sweep exit 0" $O/r02_pytest_sweep.log; then gs sweep --sweep; fi
gs cgs2 --variant cgs2
gs bcgs2_32 --variant bcgs2
gs bcgs2_128_16 --variant bcgs2 --panel 128,16
gs bcgs2_32_8 --variant bcgs2 --panel 32,8
timeout 300 python tools/bench_cholqr.py --dofs 10000000 --k 256 > $O/r02_cholqr_10Mx256.log 2>&1
grep -hE '^\{' $O/r02_cholqr_10Mx256.log | cut -c1-400

# HBM-bound kernels: 10M-DOF vectors, 12 columns -> a few hundred launches; capture a handful of each kind
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'reduce_cols|axpy_dot|elementwise|amax' \
    --launch-skip 40 -c 14 -o $O/r02_prof_blas1 -f python tools/bench_gs.py --dofs 10000000 --k 12 \
    > $O/r02_ncu_blas1.log 2>&1
ncu -i $O/r02_prof_blas1.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    > $O/r02_blas1_raw.csv 2>&1
for f in gram2_10Mx1000 gram_10Mx1000; do
  [ -f $O/r02_prof_$f.ncu-rep ] && ncu -i $O/r02_prof_$f.ncu-rep --page details > $O/r02_ncu_${f}_details.txt 2>&1
done

timeout 2400 $PT tests -m gpu -x > $O/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_pytest_gpu.log
tail -n 5 $O/r02_pytest_gpu.log
