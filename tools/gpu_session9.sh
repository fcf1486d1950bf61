#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/bench_project.py --nx 2000 --ny 1000 --k 1000 --check > gpurun_out/project_2Mx1000.log 2>&1
timeout 900 python tools/bench_project.py --nx 2500 --ny 2500 --k 2000 --reps 2 > gpurun_out/project_6Mx2000.log 2>&1
rm -f gpurun_out/kernels_waves2.log
for wv in 16 32 64; do PMC_GRAM_WAVES=$wv timeout 600 python tools/bench_kernels.py --shapes 2000000x500,4000000x1000 --reps 3 >> gpurun_out/kernels_waves2.log 2>&1; done
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_10Mx1000_v4.log 2>&1
CMD="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 $CMD > gpurun_out/plain_full.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k regex:"lincomb_dmma" -s 0 -c 1 -o gpurun_out/prof_r01_full_lincomb $CMD > gpurun_out/ncu_full_lincomb.log 2>&1
for f in gpurun_out/project_2Mx1000.log gpurun_out/project_6Mx2000.log; do grep -E '^\{|Error' $f | cut -c1-900; done
python - <<'PY'
import json
for l in open('gpurun_out/kernels_waves2.log'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['shape'], {k:(round(v['ms'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
try:
    d=json.loads([l for l in open('gpurun_out/bench_10Mx1000_v4.log') if l.startswith('{')][-1]); print(round(d['value'],2), round(d['ms_per_step'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['roofline']['gram_dmma_kernel'], d['e2e']['value'])
except Exception as e: print('ERR', e)
PY
tail -n 3 gpurun_out/ncu_full_lincomb.log | cut -c1-300
