#!/bin/bash
# Round 2, session 8 (2 GPUs):  gpurun --gpus 2 --timeout 1200 -- 'bash tools/sessions/r2_s8_multi.sh'
# the bench as the driver launches it at N = 2 (rank pinning, config-3 block, config-5 block forced on) and the
# NCCL test of the -m gpu suite.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $O/r02_s8_pytest_multi.log 2>&1; echo "pytest exit $?" >> $O/r02_s8_pytest_multi.log
tail -n 3 $O/r02_s8_pytest_multi.log | cut -c1-200
( time timeout 1200 $TR bench.py --gpus $N --steps 20 --warmup 5 --c5 on ) > $O/r02_s8_bench_${N}gpu.log 2>&1; echo "bench exit $?" >> $O/r02_s8_bench_${N}gpu.log
grep -hE '^\{' $O/r02_s8_bench_${N}gpu.log | cut -c1-600; tail -n 5 $O/r02_s8_bench_${N}gpu.log | cut -c1-200
( time timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > $O/r02_s8_bench_ref_${N}gpu.log 2>&1
grep -hE '^\{' $O/r02_s8_bench_ref_${N}gpu.log | cut -c1-400; tail -n 4 $O/r02_s8_bench_ref_${N}gpu.log | cut -c1-100
