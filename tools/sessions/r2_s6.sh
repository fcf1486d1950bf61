#!/bin/bash
# Round 2, GPU session 6 (1 GPU):  gpurun --timeout 1800 -- 'bash tools/sessions/r2_s6.sh'
# full -m gpu suite on the current tree; warp-specialised weight mode of gram2 (bit 1024) vs register scaling;
# ncu --set full of the Gram kernel as configured by default (L2-aware splits); ncu of csr_apply / amax / gather.
set -u
mkdir -p gpurun_out
O=gpurun_out
PT="python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider"
timeout 600 $PT tests/test_gpu_parity.py -m gpu -x -k "gram2" > $O/r02_s6_pytest_gram2.log 2>&1; echo "pytest gram2 exit $?" >> $O/r02_s6_pytest_gram2.log
tail -n 3 $O/r02_s6_pytest_gram2.log | cut -c1-300
timeout 900 python tools/bench_kernels.py --shapes 10000000x1000,2000000x500,1250000x1000 --no-lincomb --opts 16,1040 > $O/r02_s6_kernels_warpspec.jsonl 2>&1
grep -E '^\{' $O/r02_s6_kernels_warpspec.jsonl | cut -c1-420
BEST=16
python - <<'PY' > $O/r02_s6_best.txt
import json
rows=[json.loads(l) for l in open('gpurun_out/r02_s6_kernels_warpspec.jsonl') if l.startswith('{"shape')]
a=[r for r in rows if r['shape']=='10000000x1000']
best=min(a,key=lambda r:r['syrk_w']['ms'])
print(best['gram_opts'])
PY
BEST=$(cat $O/r02_s6_best.txt); echo "best gram_opts: $BEST"
PMC_GRAM_OPTS=$BEST timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram2_dmma --launch-skip 2 -c 1 \
    -o $O/r02_prof_gram2_final_10Mx1000 -f python tools/bench_kernels.py --shapes 10000000x1000 --no-lincomb --opts $BEST --reps 1 > $O/r02_s6_ncu_gram2.log 2>&1
ncu -i $O/r02_prof_gram2_final_10Mx1000.ncu-rep --page details > $O/r02_ncu_gram2_final_10Mx1000_details.txt 2>&1
ncu -i $O/r02_prof_gram2_final_10Mx1000.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct > $O/r02_ncu_gram2_final_10Mx1000_dram.csv 2>&1
tail -n 2 $O/r02_ncu_gram2_final_10Mx1000_dram.csv | cut -c1-400
# CSR SpMM / pairwise, amax, gather at n = 10M: projection tool at 10M DOF x 64 + a short vecops script
timeout 600 ncu --set full --clock-control none -k regex:'csr_apply_kernel|csr_pairwise|amax_kernel|gather_cols|shift_cols|gather_rows' -c 12 \
    -o $O/r02_prof_csr_misc -f python tools/bench_misc_kernels.py > $O/r02_s6_ncu_misc.log 2>&1
ncu -i $O/r02_prof_csr_misc.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct > $O/r02_ncu_csr_misc_raw.csv 2>&1
tail -n 14 $O/r02_ncu_csr_misc_raw.csv | cut -c1-330
timeout 2400 $PT tests -m gpu -x > $O/r02_s6_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r02_s6_pytest_gpu.log
tail -n 5 $O/r02_s6_pytest_gpu.log | cut -c1-300
