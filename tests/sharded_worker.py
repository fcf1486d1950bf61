"""Worker of tests/test_sharded_gloo.py: one process per rank (gloo, CPU).

Every rank runs the SAME program on its row shard (SPMD) -- the multi-GPU execution model of the
backend (pymor_b200/dist.py) -- with the C ABI served by the NumPy emulation.  Results are compared
with the oracle evaluated on the global data, and with the oracle's restatement of the reference's
gather-and-sum (src/pymor/vectorarrays/mpi.py:394-445).
"""
import os
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sps
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))

from emulated_lib import EmulatedLib  # noqa: E402
from oracle import numpy_path as O  # noqa: E402

from pymor_b200 import _lib, algorithms as alg  # noqa: E402
from pymor_b200.dist import ProcessGroupComm, partition  # noqa: E402
from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator  # noqa: E402
from pymor_b200.vectorarray import CudaVectorSpace  # noqa: E402


def main():
    real = '--real' in sys.argv          # real kernels + NCCL, one GPU per rank (tests/test_gpu_multi.py)
    if real:
        import torch
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        device = f'cuda:{local}'
        dist.init_process_group('nccl', device_id=torch.device(device))
        _lib.set_library_for_testing(None)
    else:
        device = 'cpu'
        dist.init_process_group('gloo')
        emu = EmulatedLib()

        def _sum_over_ranks(x):            # what pmc_mgs_sweep does over NVLink peer memory on the GPUs
            import torch
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t[0])
        emu.sweep_allreduce = _sum_over_ranks
        _lib.set_library_for_testing(emu)
    comm = ProcessGroupComm()
    rank, size = comm.rank, comm.size
    for dim in (1, 5, 103) + ((40013,) if real else ()):    # dim < size: some ranks own zero rows
        rng = np.random.default_rng(11)          # identical global data on every rank
        k = 9
        A = rng.standard_normal((dim, k)) + 2 * rng.standard_normal((dim, 1))
        B = rng.standard_normal((dim, k))
        w = rng.uniform(0.5, 1.5, dim)
        sp = CudaVectorSpace(dim, device=device, comm=comm)
        dims, offs = partition(dim, size)
        assert sp.local_dim == dims[rank] and sp.offset == offs[rank] and sp.dim == dim
        VA, VB = sp.from_numpy(A), sp.from_numpy(B)
        prod = CudaDiagonalOperator(w, sp)
        lo, hi = offs[rank], offs[rank] + dims[rank]

        # reductions equal the global result and the reference's "sum of local results"
        G = VA.inner(VB)
        np.testing.assert_allclose(G, O.inner(A, B), atol=1e-12)
        parts = [O.inner(A[o:o + d], B[o:o + d]) for o, d in zip(offs, dims)]
        np.testing.assert_allclose(G, O.sharded_reduce(parts), atol=1e-12)
        Gs = VA.gramian(prod)
        np.testing.assert_allclose(Gs, O.gramian(A, w), atol=1e-12)
        np.testing.assert_array_equal(Gs, Gs.T)      # bitwise symmetric also after the cross-rank sum
        if dim >= 103:
            # more than 2048 entries: the partials are summed by the all-reduce on the device (not by the host
            # reducer), whose chunks are added in different rank orders -- the lower triangle is mirrored afterwards
            W64 = rng.standard_normal((dim, 64))
            G64 = sp.from_numpy(W64).gramian(prod)
            np.testing.assert_allclose(G64, O.gramian(W64, w), atol=1e-11)
            np.testing.assert_array_equal(G64, G64.T)
        np.testing.assert_allclose(VA.pairwise_inner(VB, prod), O.pairwise_inner(A, B, w), atol=1e-12)
        np.testing.assert_allclose(VA.norm(), O.norm(A), atol=1e-12)
        np.testing.assert_allclose(VA[2:7].norm2(prod), O.norm2(A[:, 2:7], w), atol=1e-12)
        # no-communication ops act on the shard only
        Y = VA.copy(); Y.axpy(-0.5, VB); Y.scal(np.arange(k))
        np.testing.assert_allclose(Y.to_numpy(), (A - 0.5 * B) * np.arange(k), atol=1e-13)
        np.testing.assert_allclose(Y.to_torch().cpu().numpy().T, ((A - 0.5 * B) * np.arange(k))[lo:hi], atol=1e-13)
        C = rng.standard_normal((k, 4))
        np.testing.assert_allclose(VA.lincomb(C).to_numpy(), O.lincomb(A, C), atol=1e-12)
        # dofs: owner lookup + sum (mpi.py:421-432); amax: lowest shard wins ties (mpi.py:341-347)
        dofs = [0, dim - 1, dim // 2]
        np.testing.assert_array_equal(VA.dofs(dofs), O.dofs(A, dofs))
        T = A.copy()
        T[0, 0], T[dim - 1, 0] = 7.0, -7.0        # tie between first and last shard
        idx, val = sp.from_numpy(T).amax()
        oi, ov = O.amax(T)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(val, ov)
        inds = [O.amax(T[o:o + d])[0] if d else np.zeros(k, dtype=np.int64) for o, d in zip(offs, dims)]
        vals = [O.amax(T[o:o + d])[1] if d else np.full(k, -1.0) for o, d in zip(offs, dims)]
        si, sv = O.sharded_amax(inds, vals, offs)
        np.testing.assert_array_equal(idx, si)
        np.testing.assert_array_equal(val, sv)

        if dim >= k:
            # the algorithms take identical control-flow decisions on every rank
            Q, R = alg.gram_schmidt(VA, product=prod, return_R=True)
            Qo, Ro = O.gram_schmidt(A, w, return_R=True)
            np.testing.assert_allclose(R, Ro, atol=1e-11)
            np.testing.assert_allclose(Q.to_numpy(), Qo, atol=1e-11)
            # one fused sweep per pass, the sums over the ranks inside the kernel (peer inboxes); on the GPUs
            # only on request until the kernel has been seen green there (tools/gpu_session_r2_first.sh)
            if True:
                Qs, Rs = alg.gram_schmidt(VA, product=prod, return_R=True, sweep=True)
                np.testing.assert_allclose(Rs, Ro, atol=1e-11)
                np.testing.assert_allclose(Qs.to_numpy(), Qo, atol=1e-11)
            if True:
                Qc, Rc = alg.gram_schmidt(VA, product=prod, return_R=True, variant='cgs2')    # two syncs per pass
                np.testing.assert_allclose(Rc, Ro, atol=1e-10)
                np.testing.assert_allclose(Qc.to_numpy(), Qo, atol=1e-10)
                Qb, Rb = alg.block_gram_schmidt(VA, product=prod, return_R=True, panel=4)     # GEMM-bound panels
                np.testing.assert_allclose(Rb, Ro, atol=1e-10)
                np.testing.assert_allclose(Qb.to_numpy(), Qo, atol=1e-10)
            U, s = alg.pod(VA, product=prod, method='method_of_snapshots')
            Uo, so = O.pod(A, w)
            np.testing.assert_allclose(s, so, rtol=1e-10)
            assert O.subspace_angle(Uo, U.to_numpy(), w) < 1e-8
            # block-diagonal CSR operator (no halo): two-form and projection
            blocks = [sps.random(int(d), int(d), density=min(0.3, 20.0 / max(int(d), 1)), random_state=3 + i)
                      + sps.eye(int(d)) for i, d in enumerate(dims)]
            M = sps.block_diag(blocks).tocsr()
            op = CudaCsrOperator(M, sp)
            np.testing.assert_allclose(alg.project(op, VA, VB).matrix, O.project(M, A, B), atol=1e-11)
    # CSR operators coupling rows of different shards: halo exchange of the boundary rows
    # (what the reference delegates to "MPI-aware operators", src/pymor/operators/mpi.py:22-26)
    for dim in (7, 64, 301):
        rng = np.random.default_rng(5)
        M = (sps.diags([-np.ones(dim - 1), 2.5 * np.ones(dim), -0.7 * np.ones(dim - 1)], [-1, 0, 1])
             + sps.random(dim, dim, density=0.05, random_state=9)).tocsr()     # non-symmetric, long-range entries
        P = sps.diags([np.ones(dim - 1), 4 * np.ones(dim), np.ones(dim - 1)], [-1, 0, 1]).tocsr() / 6
        sp = CudaVectorSpace(dim, device=device, comm=comm)
        kv, ku = 3, 5
        V, U = rng.standard_normal((dim, kv)), rng.standard_normal((dim, ku))
        VV, UU = sp.from_numpy(V), sp.from_numpy(U)
        op, prod = CudaCsrOperator(M, sp), CudaCsrOperator(P, sp)
        np.testing.assert_allclose(op.apply(UU).to_numpy(), M @ U, atol=1e-12)
        np.testing.assert_allclose(op.apply(UU[[4, 0]]).to_numpy(), M @ U[:, [4, 0]], atol=1e-12)
        np.testing.assert_allclose(op.apply_adjoint(VV).to_numpy(), M.T @ V, atol=1e-12)
        np.testing.assert_allclose(op.apply2(VV, UU), V.T @ (M @ U), atol=1e-11)
        np.testing.assert_allclose(op.pairwise_apply2(VV, UU[:kv]), O.pairwise_inner(V, U[:, :kv], M), atol=1e-11)
        np.testing.assert_allclose(alg.project(op, VV, UU, product=prod).matrix, O.project(M, V, U, P), atol=1e-11)
        # per-rank construction: every rank hands over ITS row block only (global column numbers); the ghost map is
        # derived from the column indices -- no rank holds the global matrix
        lo, hi = sp.offset, sp.offset + sp.local_dim
        loc = CudaCsrOperator.from_local_rows(M[lo:hi], sp)
        np.testing.assert_allclose(loc.apply(UU).to_numpy(), M @ U, atol=1e-12)
        np.testing.assert_allclose(loc.apply2(VV, UU), V.T @ (M @ U), atol=1e-11)
        locP = CudaCsrOperator.from_local_rows(P[lo:hi], sp, symmetric=True)
        np.testing.assert_allclose(locP.apply2(UU, UU), U.T @ (P @ U), atol=1e-11)
        np.testing.assert_allclose(locP.apply_adjoint(VV).to_numpy(), P @ V, atol=1e-12)
        np.testing.assert_allclose(UU.norm(locP), O.norm(U, P), atol=1e-12)
        np.testing.assert_allclose(locP.apply_inverse(locP.apply(UU)).to_numpy(), U, atol=1e-7)
        if True:
            # block CG with the halo-exchanging SpMM: Riesz representatives on row-sharded arrays
            import scipy.sparse.linalg as spla
            X = prod.apply_inverse(UU)
            np.testing.assert_allclose(X.to_numpy(), spla.spsolve(P.tocsc(), U), atol=1e-8 * np.abs(U).max() * 6)
        if dim >= ku:
            Q, R = alg.gram_schmidt(UU, product=prod, return_R=True)
            Qo, Ro = O.gram_schmidt(U, P, return_R=True)
            np.testing.assert_allclose(R, Ro, atol=1e-11)
            np.testing.assert_allclose(Q.to_numpy(), Qo, atol=1e-11)
    dist.barrier()
    dist.destroy_process_group()
    print(f'rank {rank}: ok')


if __name__ == '__main__':
    main()
