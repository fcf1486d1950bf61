#!/usr/bin/env python
"""Config 3's orthonormalisation through the GEMM-bound driver: shifted CholeskyQR ("next" row N1,
reference src/pymor/algorithms/chol_qr.py:16) of k vectors with a diagonal mass-matrix product -- per
iteration one weighted Gramian (SYRK on the FP64 tensor pipe) and one block lincomb, three host syncs in
total instead of k(k-1) as in re-iterated Gram-Schmidt (tools/bench_gs.py), and the only cross-GPU exchange
is the all-reduce of the k x k Gramians.  Same input family as bench_gs.py.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dofs', dest='n', type=int, default=10_000_000)
    ap.add_argument('--k', type=int, default=256)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    import torch

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    n, k = args.n, args.k
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    comm = None
    if world > 1:                                   # torchrun: rows sharded over ranks (strong scaling)
        import torch.distributed as dist

        from pymor_b200.dist import ProcessGroupComm
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = ProcessGroupComm()
    space = CudaVectorSpace(n, device=torch.device('cuda', torch.cuda.current_device()), comm=comm)
    gen = torch.Generator(device=space.device).manual_seed(3 + rank)
    w = torch.rand(space.local_dim, generator=gen, device=space.device, dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, space)

    def make():
        A = space.random_device(k, seed=1, scale=1.0 / np.sqrt(n))
        A.axpy(2.0, space.random_device(1, seed=2, scale=1.0 / np.sqrt(n)))
        return A

    alg.shifted_chol_qr(make()[:8].copy(), product=prod, product_norm=float(np.sqrt(1.5)), copy=False)   # warm-up
    times, launches = [], 0
    for _ in range(args.reps):
        A = make()
        torch.cuda.synchronize()
        l0 = space.ctx.launch_count
        t0 = time.perf_counter()
        Q, R = alg.shifted_chol_qr(A, product=prod, return_R=True, product_norm=float(np.sqrt(1.5)), copy=False)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = space.ctx.launch_count - l0
    err = float(np.max(np.abs(Q.gramian(prod) - np.eye(len(Q)))))
    dt = min(times)
    iters = 3
    flop = iters * (n * k * (k + 1.0) + 2.0 * n * k * k)          # Gramian (SYRK) + lincomb per iteration
    out = {'workload': f'shifted_chol_qr {n} x {k}, diag-W', 'n_gpus': world, 'seconds': dt, 'all_seconds': times,
           'algorithmic_flop': flop, 'tflops': flop / dt * 1e-12, 'gpu_launches': launches, 'orth_err': err,
           'len_Q': len(Q), 'note': 'flop assume the default 3 iterations; compare seconds with tools/bench_gs.py'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
