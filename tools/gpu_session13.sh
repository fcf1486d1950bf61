#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread -k "uploading or gram_parity or fused" > gpurun_out/pytest_s13.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s13.log
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10Mx1000_v5.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --dofs 2000000 --k 500 --no-cpu-baseline > gpurun_out/bench_2Mx500_v5.log 2>&1
tail -n 3 gpurun_out/pytest_s13.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_10Mx1000_v5.log','gpurun_out/bench_2Mx500_v5.log'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, round(d['value'],2), round(d['ms_per_step'],1), d['phases_ms'], d['e2e'])
    except Exception as e: print(f, 'ERR', e, open(f).read()[-1500:])
PY
