#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --timeout-method thread -k "fused or gram_schmidt or streaming or gram_parity" > gpurun_out/pytest_s8.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s8.log
timeout 600 python tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/gs_10Mx256_fused.log 2>&1
PMC_NO_FUSE_AXPY=1 timeout 600 python tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/gs_10Mx256_unfused.log 2>&1
timeout 600 python tools/bench_gs.py --dofs 2000000 --k 256 > gpurun_out/gs_2Mx256_fused.log 2>&1
rm -f gpurun_out/kernels_waves.log
for wv in 4 8 16; do PMC_GRAM_WAVES=$wv timeout 600 python tools/bench_kernels.py --shapes 2000000x500,4000000x1000 --reps 3 >> gpurun_out/kernels_waves.log 2>&1; done
tail -n 3 gpurun_out/pytest_s8.log
for f in gpurun_out/gs_10Mx256_fused.log gpurun_out/gs_10Mx256_unfused.log gpurun_out/gs_2Mx256_fused.log; do cut -c1-520 $f; done
python - <<'PY'
import json
for l in open('gpurun_out/kernels_waves.log'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['shape'], {k:(round(v['ms'],2), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)})
PY
