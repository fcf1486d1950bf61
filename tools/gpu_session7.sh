#!/bin/bash
# Single-GPU session: full parity suite, final round-1 numbers, ncu re-capture of the tuned kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
rm -f gpurun_out/kernels_v3.log
for o in 0 1; do PMC_GRAM_OPTS=$o timeout 600 python tools/bench_kernels.py --shapes 2000000x500,4000000x1000 --reps 3 >> gpurun_out/kernels_v3.log 2>&1; done
timeout 600 python bench.py --steps 3 --warmup 3 --dofs 2000000 --k 500 --cpu-sample 100000 > gpurun_out/bench_2Mx500_v3.log 2>&1
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_10Mx1000_v3.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_10Mx1000_v3.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.log 2>&1
timeout 600 python tools/bench_gs.py --dofs 10000000 --k 256 --cpu-n 100000 > gpurun_out/gs_10Mx256_v3.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --dofs 2000000 --k 500 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lincomb_dmma|gram_dmma" -s 2 -c 3 -o gpurun_out/prof_r01b $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
python - <<'PY'
import json
for l in open('gpurun_out/kernels_v3.log'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['gram_opts'], d['shape'], {k:(round(v['ms'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
for f in ('gpurun_out/bench_2Mx500_v3.log','gpurun_out/bench_10Mx1000_v3.log'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['phases_ms'], round(d['roofline']['frac'],3), d['roofline']['gram_dmma_kernel'], d['e2e'], d['cpu_baseline'])
    except Exception as e: print(f, 'ERR', e)
PY
cut -c1-600 gpurun_out/gs_10Mx256_v3.log; cut -c1-400 gpurun_out/bench_reference_arm.log
