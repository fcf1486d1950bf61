#!/bin/bash
# Multi-GPU session of round 2 (start with `gpurun --gpus 2`, repeat with --gpus 8 once green):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_session_r2_multi.sh'
# The fused Gram-Schmidt sweep sums its coefficients across the GPUs inside the kernel (peer inboxes over
# NVLink); its waits give up after 5 s, every command runs under `timeout`.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
PMC_TEST_SWEEP=1 timeout 600 $TR tests/sharded_worker.py --real > gpurun_out/r02_sharded_worker_${N}gpu.log 2>&1
echo "exit $?" >> gpurun_out/r02_sharded_worker_${N}gpu.log
tail -n 5 gpurun_out/r02_sharded_worker_${N}gpu.log
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 > gpurun_out/r02_gs_${N}gpu_pairs.log 2>&1
if grep -q "exit 0" gpurun_out/r02_sharded_worker_${N}gpu.log; then
    timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 --sweep > gpurun_out/r02_gs_${N}gpu_sweep.log 2>&1
fi
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 --variant cgs2 > gpurun_out/r02_gs_${N}gpu_cgs2.log 2>&1
grep -hE '^\{' gpurun_out/r02_gs_${N}gpu_cgs2.log | cut -c1-500
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 > gpurun_out/r02_gs_${N}gpu_bcgs2.log 2>&1
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 --panel 128,16 > gpurun_out/r02_gs_${N}gpu_bcgs2_128_16.log 2>&1
timeout 600 $TR tools/bench_gs.py --dofs 10000000 --k 256 --variant bcgs2 --panel 32,8 > gpurun_out/r02_gs_${N}gpu_bcgs2_32_8.log 2>&1
grep -hE '^\{' gpurun_out/r02_gs_${N}gpu_bcgs2.log | cut -c1-500
grep -hE '^\{' gpurun_out/r02_gs_${N}gpu_bcgs2_128_16.log | cut -c1-500
grep -hE '^\{' gpurun_out/r02_gs_${N}gpu_bcgs2_32_8.log | cut -c1-500
timeout 600 $TR tools/bench_cholqr.py --dofs 10000000 --k 256 > gpurun_out/r02_cholqr_${N}gpu.log 2>&1
grep -hE '^\{' gpurun_out/r02_cholqr_${N}gpu.log | cut -c1-400
# config 5 (Galerkin projection V^T (A V) of a 5-point CSR operator, halo exchange): 6.25M DOF x 2000 basis per GPU
timeout 900 $TR tools/bench_project.py --nx 2500 --ny $((2500 * N)) --k 2000 --reps 2 > gpurun_out/r02_project_${N}gpu.log 2>&1
grep -hE '^\{' gpurun_out/r02_project_${N}gpu.log | cut -c1-600
timeout 1200 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.log 2>&1
grep -hE '^\{' gpurun_out/r02_gs_${N}gpu_pairs.log gpurun_out/r02_gs_${N}gpu_sweep.log gpurun_out/r02_bench_${N}gpu.log | cut -c1-600
