#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread -k "gram_schmidt_golden or chol_qr or gram_parity" > gpurun_out/pytest_gpu_s4.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_s4.log
for o in 0 1 3 7; do
  PMC_GRAM_OPTS=$o timeout 600 python tools/bench_kernels.py --shapes 2000000x500,4000000x1000 --reps 3 >> gpurun_out/kernels_ab.log 2>&1
done
tail -4 gpurun_out/pytest_gpu_s4.log
python - <<'PY'
import json
for l in open('gpurun_out/kernels_ab.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['gram_opts'], d['shape'], {k:(round(v['ms'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
PY
