#!/bin/bash
# Round 2, session 14 (4 GPUs, short): bitwise symmetry of Gramians after the cross-GPU all-reduce (pmc_mirror_lower) --
# the sharded worker on both all-reduce paths and a short bench (parity block).
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tests/sharded_worker.py --real > $O/r02_s14_worker_${N}gpu.log 2>&1; echo "exit $?" >> $O/r02_s14_worker_${N}gpu.log; tail -n 3 $O/r02_s14_worker_${N}gpu.log | cut -c1-200
PMC_NATIVE_COMM=0 timeout 300 $TR tests/sharded_worker.py --real > $O/r02_s14_worker_torchdist_${N}gpu.log 2>&1; echo "exit $?" >> $O/r02_s14_worker_torchdist_${N}gpu.log; tail -n 2 $O/r02_s14_worker_torchdist_${N}gpu.log | cut -c1-200
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-weak --no-gs --c5 off > $O/r02_s14_bench_${N}gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s14_bench_${N}gpu.log
grep -hE '^\{' $O/r02_s14_bench_${N}gpu.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['parity']['full_size'], d['parity']['sample']['ok'], d['roofline']['frac'])"
