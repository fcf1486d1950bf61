#!/bin/bash
# Round 2, session 15 (1 GPU): the bench with the driver's arguments on the final tree (wide warp tiles).
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/r02_s15_bench_1gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s15_bench_1gpu.log
grep -hE '^\{' $O/r02_s15_bench_1gpu.log | cut -c1-400; tail -n 5 $O/r02_s15_bench_1gpu.log | cut -c1-100
