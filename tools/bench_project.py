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


def stencil_rows(nx, ny, lo, hi):
    """Rows [lo, hi) of :func:`stencil` as a CSR block of shape (hi - lo, nx * ny) with GLOBAL column numbers,
    assembled without ever forming the global matrix (what a rank hands to CudaCsrOperator.from_local_rows)."""
    n = nx * ny
    i = np.arange(lo, hi, dtype=np.int64)
    ix = i % nx
    rows, cols, vals = [i - lo], [i], [np.full(len(i), 5.0)]
    for off, ok in ((-1, ix > 0), (1, ix < nx - 1), (-nx, i >= nx), (nx, i < n - nx)):
        rows.append((i - lo)[ok]); cols.append((i + off)[ok]); vals.append(np.full(int(ok.sum()), -1.0))
    return sps.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(hi - lo, n))


def run_projection(space, nx, ny, k, reps=2, comm=None):
    """Config 5 on `space` (dim = nx * ny): `project(op, V, V)` of pyMOR (projection.py:28, rule
    action_apply_basis :118-143) with the 5-point CSR operator built per rank from its own rows.  Returns the
    result dict (host wall-clock around synchronised calls after a barrier, max over the ranks)."""
    import torch

    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator
    n = nx * ny
    lo, hi = space.offset, space.offset + space.local_dim
    local = stencil_rows(nx, ny, lo, hi)
    nnz_local = local.nnz
    op = CudaCsrOperator.from_local_rows(local, space, symmetric=True)
    full = CudaCsrOperator.from_local_rows(local, space, symmetric=False)      # same matrix, full (non-mirrored) product
    del local
    V = space.random_device(k, seed=11, scale=1.0 / np.sqrt(n))
    torch.cuda.synchronize()

    def timed(fn, warm=2, reps=reps):
        # A FIXED number of warm-up calls: every rank must issue the same sequence of collectives
        for _ in range(warm):
            fn()
            torch.cuda.synchronize()
        if comm:
            comm.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, out

    t_proj, P = timed(lambda: alg.project(op, V, V).matrix)
    l0 = space.ctx.launch_count
    alg.project(op, V, V).matrix
    launches = space.ctx.launch_count - l0
    t_full, Pf = timed(lambda: alg.project(full, V[:min(k, 512)], V[:min(k, 512)]).matrix, warm=1, reps=1)
    coeff = np.random.default_rng(0).standard_normal((k, 8))
    t_rec, _ = timed(lambda: V.lincomb(coeff), warm=3)
    t_apply, _ = timed(lambda: op.apply(V[:64]), warm=5)
    peak = space.ctx.measure_fp64_peak(True, 200.0)
    world = comm.size if comm else 1
    nnz = nnz_local
    if comm:                                                # totals / the slowest rank's times
        gathered = comm.allgather_object((int(nnz_local), t_proj, t_rec, t_apply))
        nnz = int(sum(g[0] for g in gathered))
        t_proj, t_rec, t_apply = (max(g[i] for g in gathered) for i in (1, 2, 3))
    flop_gemm = 2.0 * n * k * k + 2.0 * nnz * k             # BASELINE's count: full product
    flop_alg = n * k * (k + 1.0) + 2.0 * nnz * k            # symmetric result: lower triangle only
    # parity: (1) the mirrored lower-only result against the full product of the same kernels on a sub-basis;
    # (2) a few entries against an independent evaluation through the SpMM and streaming-reduction kernels
    kc = Pf.shape[0]
    err_full = float(np.max(np.abs(P[:kc, :kc] - Pf)) / np.max(np.abs(Pf)))
    idx = [0, 1, k // 2, k - 1]
    MV = op.apply(V[idx])
    ref = V[idx].inner(MV)
    err_indep = float(np.max(np.abs(P[np.ix_(idx, idx)] - ref)) / np.max(np.abs(ref)))
    return {'workload': f'pymor project(op, V, V) = V^T (A V): {nx} x {ny} grid = {n} DOF, 5-point CSR ({nnz} nnz, symmetric), '
                        f'{k} basis vectors, rows sharded over {world} GPU(s) with halo exchange',
            'n_gpus': world, 'n': n, 'k': k, 'project_seconds': t_proj,
            'project_tflops_gemm_equivalent': flop_gemm / t_proj * 1e-12,
            'project_tflops_algorithmic': flop_alg / t_proj * 1e-12,
            'frac_of_dmma_peak_algorithmic': flop_alg / t_proj * 1e-12 / (peak * world), 'dmma_peak_tflops': peak,
            'launches_per_project': launches, 'reconstruct_lincomb_k_x_8_seconds': t_rec,
            'reconstruct_gbs': 8.0 * n * (k + 8) / t_rec * 1e-9, 'apply_64_vectors_seconds': t_apply,
            'apply_gbs': (12.0 * nnz + 16.0 * n * 64) / t_apply * 1e-9,
            'parity': {'lower_only_vs_full_product_rel': err_full, 'entries_vs_spmm_plus_inner_rel': err_indep,
                       'ok': bool(err_full < 1e-12 and err_indep < 1e-12)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=2000)
    ap.add_argument('--ny', type=int, default=1000)
    ap.add_argument('--k', type=int, default=1000)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    import torch

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
    space = CudaVectorSpace(args.nx * args.ny, device=torch.device('cuda', torch.cuda.current_device()), comm=comm)
    out = run_projection(space, args.nx, args.ny, args.k, reps=args.reps, comm=comm)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
