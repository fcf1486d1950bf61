#!/bin/bash
# Third GPU session: parity after kernel changes, L2-keep A/B on Gram-Schmidt, re-measure POD.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_gs.py --n 10000000 --k 256 > gpurun_out/gs_10Mx256_keep.log 2>&1
PMC_NO_L2_KEEP=1 timeout 600 python tools/bench_gs.py --n 10000000 --k 256 > gpurun_out/gs_10Mx256_nokeep.log 2>&1
timeout 600 python tools/bench_gs.py --n 5000000 --k 256 > gpurun_out/gs_5Mx256_keep.log 2>&1
PMC_NO_L2_KEEP=1 timeout 600 python tools/bench_gs.py --n 5000000 --k 256 > gpurun_out/gs_5Mx256_nokeep.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --n 2000000 --k 500 --no-cpu-baseline --no-e2e > gpurun_out/bench_2Mx500_v2.log 2>&1
timeout 1500 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_10Mx1000_v2.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; for f in gpurun_out/gs_*keep.log; do echo $f; cut -c1-420 $f; done
python - <<'PY'
import json
for f in ('gpurun_out/bench_2Mx500_v2.log','gpurun_out/bench_10Mx1000_v2.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'], d['roofline']['gram_dmma_kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
