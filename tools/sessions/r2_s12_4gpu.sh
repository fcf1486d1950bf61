#!/bin/bash
# Round 2, session 12 (4 GPUs): the bench as the driver launches it at N = 4 (the one N not yet run this round).
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( time timeout 800 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/r02_s12_bench_${N}gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s12_bench_${N}gpu.log
grep -hE '^\{' $O/r02_s12_bench_${N}gpu.log | cut -c1-500; tail -n 5 $O/r02_s12_bench_${N}gpu.log | cut -c1-200
