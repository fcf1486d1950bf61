#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 --timeout-method thread > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/gs_${N}gpu_10Mx256_fused.log 2>&1
echo "exit $?" >> gpurun_out/gs_${N}gpu_10Mx256_fused.log
timeout 900 $TR tools/bench_project.py --nx 2500 --ny 5000 --k 2000 --reps 2 > gpurun_out/project_${N}gpu_12Mx2000.log 2>&1
echo "exit $?" >> gpurun_out/project_${N}gpu_12Mx2000.log
timeout 1500 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_10Mx1000.log 2>&1
echo "exit $?" >> gpurun_out/bench_${N}gpu_10Mx1000.log
tail -n 3 gpurun_out/pytest_multi.log
for f in gpurun_out/gs_${N}gpu_10Mx256_fused.log gpurun_out/project_${N}gpu_12Mx2000.log gpurun_out/bench_${N}gpu_10Mx1000.log; do echo "== $f"; grep -E '^\{|exit|Error' $f | cut -c1-1300; done
