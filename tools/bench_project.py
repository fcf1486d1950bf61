#!/usr/bin/env python
"""Config 5 of BASELINE.json at per-GPU scale: Galerkin projection V^T (A V) of a sparse operator
(5-point stencil, CSR) onto a k-vector basis plus the reconstruction lincomb -- `project()` +
`lincomb`, the reductor hot path (reference: src/pymor/algorithms/projection.py:118-143,
src/pymor/reductors/basic.py:125-127, :177-186).

The fused two-form never materialises the n x k product A V: row panels of it live in the workspace
(L2-sized) and are contracted by the TMA + DMMA Gram kernel.  Prints one JSON line (quoted in
DESIGN.md / profiles; not the driver's bench contract).  Works under torchrun (row-sharded, halo
exchange of the grid lines at the shard boundaries).
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault('PYTORCH_CUDA_ALLOC_CONF', 'expandable_segments:True')
import numpy as np  # noqa: E402
import scipy.sparse as sps  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def stencil(nx, ny):
    """5-point Laplacian + identity on an nx x ny grid (SPD), CSR."""
    ex, ey = np.ones(nx), np.ones(ny)
    tx = sps.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
    ty = sps.diags([-ey[:-1], 2 * ey, -ey[:-1]], [-1, 0, 1])
    return (sps.kron(sps.eye(ny), tx) + sps.kron(ty, sps.eye(nx)) + sps.eye(nx * ny)).tocsr()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=2000)
    ap.add_argument('--ny', type=int, default=1000)
    ap.add_argument('--k', type=int, default=1000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--check', action='store_true', help='compare a few entries with SciPy')
    args = ap.parse_args()
    import torch

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    comm = None
    if world > 1:
        import torch.distributed as dist

        from pymor_b200.dist import ProcessGroupComm
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm = ProcessGroupComm()
    n, k = args.nx * args.ny, args.k
    M = stencil(args.nx, args.ny)
    space = CudaVectorSpace(n, device=torch.device('cuda', torch.cuda.current_device()), comm=comm)
    op = CudaCsrOperator(M, space)
    V = space.random_device(k, seed=11, scale=1.0 / np.sqrt(n))
    torch.cuda.synchronize()

    def timed(fn, warm=2):
        # A FIXED number of warm-up calls: every rank must issue the same sequence of collectives
        # (a time-based warm-up loop deadlocked the 8-rank run: ranks did different numbers of halo
        # exchanges).  Short runs start from idle clocks, hence at least two calls.
        for _ in range(warm):
            fn()
            torch.cuda.synchronize()
        if comm:
            comm.barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.reps, out

    t_proj, P = timed(lambda: alg.project(op, V, V).matrix)
    l0 = space.ctx.launch_count
    alg.project(op, V, V).matrix
    launches = space.ctx.launch_count - l0
    coeff = np.random.default_rng(0).standard_normal((k, 8))
    t_rec, _ = timed(lambda: V.lincomb(coeff), warm=5)
    t_apply, _ = timed(lambda: op.apply(V[:64]), warm=20)
    peak = space.ctx.measure_fp64_peak(True, 200.0)
    nnz = M.nnz
    flop = 2.0 * n * k * k + 2.0 * nnz * k
    out = {'workload': f'project(V^T A V): {args.nx}x{args.ny} grid = {n} DOF, 5-point CSR ({nnz} nnz), {k} basis vectors',
           'n_gpus': world, 'project_seconds': t_proj, 'project_tflops': flop / t_proj * 1e-12,
           'frac_of_dmma_peak': flop / t_proj * 1e-12 / (peak * world), 'dmma_peak_tflops': peak,
           'launches_per_project': launches, 'reconstruct_lincomb_k_x_8_seconds': t_rec,
           'reconstruct_gbs': 8.0 * n * (k + 8) / t_rec * 1e-9, 'apply_64_vectors_seconds': t_apply,
           'apply_gbs': (12.0 * nnz + 16.0 * n * 64) / t_apply * 1e-9, 'symmetry_defect': float(np.max(np.abs(P - P.T)))}
    if args.check and world == 1:
        idx = [0, 1, k // 2, k - 1]
        Vh = V[idx].to_numpy()
        ref = Vh.T @ (M @ Vh)
        out['max_abs_err_vs_scipy'] = float(np.max(np.abs(P[np.ix_(idx, idx)] - ref)))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
