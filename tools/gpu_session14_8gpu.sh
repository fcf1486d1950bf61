#!/bin/bash
# 8-GPU session: weak-scaling POD bench (contract line), strong-scaling GS, sharded projection.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 1200 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu_10Mx1000.log 2>&1
echo "exit $?" >> gpurun_out/bench_${N}gpu_10Mx1000.log
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/gs_${N}gpu_10Mx256_fused.log 2>&1
echo "exit $?" >> gpurun_out/gs_${N}gpu_10Mx256_fused.log
timeout 900 $TR tools/bench_project.py --nx 2000 --ny 8000 --k 1000 --reps 2 > gpurun_out/project_${N}gpu_16Mx1000.log 2>&1
echo "exit $?" >> gpurun_out/project_${N}gpu_16Mx1000.log
for f in gpurun_out/bench_${N}gpu_10Mx1000.log gpurun_out/gs_${N}gpu_10Mx256_fused.log gpurun_out/project_${N}gpu_16Mx1000.log; do echo "== $f"; grep -E '^\{|exit|Error' $f | cut -c1-1500; done
