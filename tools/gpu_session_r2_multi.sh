#!/bin/bash
# Multi-GPU session of round 2.  A call on N GPUs is charged N x its box time, so:
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_session_r2_multi.sh'          # everything, first
#   gpurun --gpus 8 --timeout 600  -- 'QUICK=1 bash tools/gpu_session_r2_multi.sh'  # once green: the essentials only
# The fused Gram-Schmidt sweep sums its coefficients across the GPUs inside the kernel (peer inboxes over
# NVLink); its waits give up after 5 s, every command runs under `timeout`.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
QUICK=${QUICK:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
show() { for f in "$@"; do grep -hE '^\{' "$f" 2>/dev/null | cut -c1-600; done; }
gs() { tag=$1; shift; timeout 300 $TR tools/bench_gs.py --dofs 10000000 --k 256 "$@" > gpurun_out/r02_gs_${N}gpu_${tag}.log 2>&1; show gpurun_out/r02_gs_${N}gpu_${tag}.log; }

if [ "$QUICK" != "1" ]; then
    PMC_TEST_SWEEP=1 timeout 600 $TR tests/sharded_worker.py --real > gpurun_out/r02_sharded_worker_${N}gpu.log 2>&1
    echo "exit $?" >> gpurun_out/r02_sharded_worker_${N}gpu.log
    tail -n 5 gpurun_out/r02_sharded_worker_${N}gpu.log
    gs pairs
fi
# config 3 (10M x 256, rows sharded): swept MGS (in-kernel sums over NVLink), classical GS with re-orthogonalisation
# (two syncs per pass), block GS on the tensor cores, shifted CholeskyQR
if [ "$QUICK" = "1" ] || grep -q "exit 0" gpurun_out/r02_sharded_worker_${N}gpu.log; then
    gs sweep --sweep
fi
gs cgs2 --variant cgs2
gs bcgs2_32_8 --variant bcgs2 --panel 32,8
if [ "$QUICK" != "1" ]; then
    gs bcgs2 --variant bcgs2
    gs bcgs2_128_16 --variant bcgs2 --panel 128,16
fi
timeout 300 $TR tools/bench_cholqr.py --dofs 10000000 --k 256 > gpurun_out/r02_cholqr_${N}gpu.log 2>&1
show gpurun_out/r02_cholqr_${N}gpu.log
# config 5 (Galerkin projection V^T (A V) of a 5-point CSR operator, halo exchange): 6.25M DOF x 2000 basis per GPU
timeout 600 $TR tools/bench_project.py --nx 2500 --ny $((2500 * N)) --k 2000 --reps 2 > gpurun_out/r02_project_${N}gpu.log 2>&1
show gpurun_out/r02_project_${N}gpu.log
# the contract line (weak scaling, 10M x 1000 per GPU)
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_${N}gpu.log 2>&1
show gpurun_out/r02_bench_${N}gpu.log
