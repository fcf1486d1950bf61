#!/bin/bash
# Second GPU session: full parity suite, full-size bench, GS bench, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 600 --timeout-method thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_10Mx1000.log 2>&1
echo "bench full exit $?" >> gpurun_out/bench_10Mx1000.log
timeout 900 python tools/bench_gs.py --n 10000000 --k 256 --cpu-n 100000 > gpurun_out/gs_10Mx256.log 2>&1
echo "gs exit $?" >> gpurun_out/gs_10Mx256.log
CMD="python bench.py --steps 2 --warmup 1 --n 2000000 --k 500 --no-cpu-baseline --no-e2e"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"lincomb_dmma|gram_dmma" -s 2 -c 3 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/bench_10Mx1000.log gpurun_out/gs_10Mx256.log gpurun_out/ncu_launches.log gpurun_out/ncu_full.log | cut -c1-600
