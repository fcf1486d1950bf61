#!/bin/bash
# Multi-GPU session (run with gpurun --gpus N): NCCL parity test, sharded POD bench, sharded GS.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" > gpurun_out/multi_info.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 --timeout-method thread > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 2 --dofs 2000000 --k 500 --no-e2e > gpurun_out/bench_${N}gpu_2Mx500.log 2>&1
echo "exit $?" >> gpurun_out/bench_${N}gpu_2Mx500.log
timeout 1500 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_10Mx1000.log 2>&1
echo "exit $?" >> gpurun_out/bench_${N}gpu_10Mx1000.log
timeout 900 $TR tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/gs_${N}gpu_10Mx256.log 2>&1
echo "exit $?" >> gpurun_out/gs_${N}gpu_10Mx256.log
tail -3 gpurun_out/pytest_multi.log
for f in gpurun_out/bench_${N}gpu_2Mx500.log gpurun_out/bench_${N}gpu_10Mx1000.log gpurun_out/gs_${N}gpu_10Mx256.log; do echo "== $f"; grep -E '^\{|exit|Error' $f | cut -c1-900; done
