#!/bin/bash
# Round 2, multi-GPU session:  gpurun --gpus N --timeout 1500 -- 'bash tools/sessions/r2_s3_multi.sh'
# (charged N x its box time).  NCCL path on hardware: the sharded worker (parity of every sharded operation incl.
# the fused sweep over peer inboxes, halo exchange, per-rank CSR blocks, block CG), the bench at this N (strong
# scaling, parity block, weak extra, config-5 secondary), Gram-Schmidt variants at config 3.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=$(nvidia-smi -L | wc -l)
QUICK=${QUICK:-0}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
show() { for f in "$@"; do grep -hE '^\{' "$f" 2>/dev/null | cut -c1-${W:-700}; done; }
{ nproc; free -g | head -2; nvidia-smi topo -m | head -12; numactl -H 2>/dev/null | head -8; lscpu | grep -E "NUMA|Model name|Socket" ; } > $O/r02_multi_info_${N}gpu.txt 2>&1
PMC_NATIVE_COMM=0 timeout 900 $TR tests/sharded_worker.py --real > $O/r02_sharded_worker_${N}gpu.log 2>&1; echo "exit $?" >> $O/r02_sharded_worker_${N}gpu.log
tail -n 4 $O/r02_sharded_worker_${N}gpu.log | cut -c1-300
# the same worker with the library-side NCCL communicator (pmc_allreduce_sum) instead of torch.distributed
PMC_NATIVE_COMM=1 timeout 900 $TR tests/sharded_worker.py --real > $O/r02_sharded_worker_native_${N}gpu.log 2>&1; echo "exit $?" >> $O/r02_sharded_worker_native_${N}gpu.log
tail -n 3 $O/r02_sharded_worker_native_${N}gpu.log | cut -c1-300
C5=on; [ "$N" = "8" ] && C5=auto
timeout 1200 $TR bench.py --gpus $N --steps ${STEPS:-4} --warmup 3 --c5 $C5 --accelerate > $O/r02_bench_${N}gpu.log 2>&1; echo "bench exit $?" >> $O/r02_bench_${N}gpu.log
W=6000 show $O/r02_bench_${N}gpu.log; tail -n 2 $O/r02_bench_${N}gpu.log | cut -c1-400
if [ "$QUICK" != "1" ]; then
    PMC_NATIVE_COMM=0 timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-4} --warmup 3 --c5 off --no-e2e --no-weak > $O/r02_bench_${N}gpu_torchdist.log 2>&1
    W=1500 show $O/r02_bench_${N}gpu_torchdist.log
fi
gs() { tag=$1; shift; timeout 300 $TR tools/bench_gs.py --dofs 10000000 --k 256 "$@" > $O/r02_gs_${N}gpu_${tag}.log 2>&1; show $O/r02_gs_${N}gpu_${tag}.log; }
gs sweep --sweep
gs bcgs2_32 --variant bcgs2
if [ "$QUICK" != "1" ]; then
    gs pairs
    gs cgs2 --variant cgs2
fi
timeout 300 $TR tools/bench_cholqr.py --dofs 10000000 --k 256 > $O/r02_cholqr_${N}gpu.log 2>&1; show $O/r02_cholqr_${N}gpu.log
