#!/bin/bash
# Round 2, GPU session 4 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/sessions/r2_s4.sh'
# register-weight variants of gram2 (bit 512: ni-major), split granularity (gram_waves) vs DRAM over-fetch, the full
# -m gpu suite, the bench with the CSR-product POD block.
set -u
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider"
timeout 900 python tools/bench_kernels.py --shapes 10000000x1000,2000000x500 --no-lincomb --opts 16,528 > $O/r02_s4_kernels_nimajor.jsonl 2>&1
grep -E '^\{' $O/r02_s4_kernels_nimajor.jsonl | cut -c1-420
for W in 32 64 128 256; do
  PMC_GRAM_WAVES=$W timeout 300 python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts 16 > $O/r02_s4_kernels_waves$W.jsonl 2>&1
  grep -E '^\{"shape' $O/r02_s4_kernels_waves$W.jsonl | cut -c1-420
  PMC_GRAM_WAVES=$W timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
     -k regex:gram2_dmma --launch-skip 2 -c 1 --csv --log-file $O/r02_s4_ncu_waves$W.csv \
     python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts 16 --reps 1 > /dev/null 2>&1
  grep -E '^"[0-9]' $O/r02_s4_ncu_waves$W.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ' '; echo
done
timeout 2400 $PT tests -m gpu -x > $O/r02_s4_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s4_pytest_gpu.log
tail -n 5 $O/r02_s4_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 2 > $O/r02_s4_bench_1gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s4_bench_1gpu.log
grep -E '^\{' $O/r02_s4_bench_1gpu.log | cut -c1-6000; tail -n 2 $O/r02_s4_bench_1gpu.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_s4_bench_ref.log 2>&1
grep -E '^\{' $O/r02_s4_bench_ref.log | cut -c1-1500
