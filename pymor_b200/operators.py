"""Product / system operators on a :class:`CudaVectorSpace`.

pyMOR evaluates ``product.apply2(V, U)`` as ``V.inner(product.apply(U))`` and
``pairwise_apply2`` as ``V.pairwise_inner(product.apply(U))`` (reference:
src/pymor/operators/interface.py:79-143), materialising the n x k array ``W U`` first
(src/pymor/operators/numpy.py:232-237).  The operators here override both with fused device
kernels so that weighted Gramians, norms and Galerkin projections never allocate that temporary:

* :class:`CudaDiagonalOperator` -- diagonal (lumped-mass / FV L2) products: the weight is applied to
  the B-fragment inside the TMA + DMMA Gram kernel resp. inside the streaming reduction.
* :class:`CudaCsrOperator` -- sparse (FEM mass/stiffness) matrices in CSR: ``apply2`` forms row panels
  of ``M U`` in an L2-sized scratch and contracts them at once; on row-sharded spaces the rows of
  ``U`` owned by other ranks are fetched by a neighbour (halo) exchange first.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import torch

from pymor_b200.complexarray import CudaComplexVectorArrayImpl
from pymor_b200.interface import Operator
from pymor_b200.solvers import CgSolver, DiagonalSolver
from pymor_b200._lib import PMC_LOWER_ONLY
from pymor_b200.vectorarray import CudaVectorArray, CudaVectorArrayImpl, CudaVectorSpace


def _is_complex(U):
    return type(U.impl) is CudaComplexVectorArrayImpl


def _apply_by_parts(op, U):
    """A real linear operator acts on the real and imaginary parts separately."""
    fn = op.apply if hasattr(op, 'apply') else op
    re, im = fn(U.real), fn(U.imag)
    return CudaVectorArray(U.space, CudaComplexVectorArrayImpl(U.space, re.impl, im.impl))


class CudaDiagonalOperator(Operator):
    """``x -> w * x`` with a device-resident weight vector (this rank's rows).

    Parameters
    ----------
    weights
        1-D fp64 array-like of the GLOBAL diagonal (NumPy) or a CUDA tensor of this rank's
        ``space.local_dim`` entries.
    space
        The :class:`CudaVectorSpace` (source and range).
    """

    linear = True
    fused_apply2 = True        # apply2 / pairwise_apply2 weigh inside the reduction kernels: no temporary W U

    def __init__(self, weights, space, solver=None, name=None):
        assert isinstance(space, CudaVectorSpace)
        self.weights = weights
        self.space = space
        self.name = name
        self.solver = solver if solver is not None else DiagonalSolver()     # apply_inverse: reciprocal weights
        self.source = self.range = space
        if isinstance(weights, torch.Tensor):
            w = weights.to(device=space.device, dtype=torch.float64).contiguous()
            assert w.shape == (space.local_dim,)
        else:
            w = np.asarray(weights, dtype=np.float64).reshape(-1)
            assert w.shape == (space.dim,)
            w = torch.from_numpy(np.ascontiguousarray(w[space.offset:space.offset + space.local_dim])).to(space.device)
        if w.numel() == 0:
            w = torch.zeros(16, dtype=torch.float64, device=space.device)
        self._w = w

    @classmethod
    def from_scipy(cls, matrix, space, name=None):
        """From a SciPy sparse matrix that is diagonal (e.g. the FV ``l2_product``,
        src/pymor/discretizers/builtin/fv.py:516-539)."""
        m = sps.coo_matrix(matrix)
        assert np.all(m.row == m.col), 'matrix is not diagonal'
        return cls(np.asarray(matrix.diagonal(), dtype=np.float64), space, name=name)

    def apply(self, U, mu=None):
        assert U in self.source
        if _is_complex(U):
            return _apply_by_parts(self, U)
        pu, csu, k, keep = U.impl._panel(U.ind)
        out = CudaVectorArrayImpl.allocate(self.space, k)
        if k and self.space.local_dim:
            self.space.ctx.call('pmc_diag_apply', self._w.data_ptr(), pu, csu, out._ptr(0), out.ld, k,
                                self.space.local_dim, 0)
        del keep
        return CudaVectorArray(self.space, out)

    def apply_adjoint(self, V, mu=None):
        return self.apply(V, mu=mu)

    def _assemble_lincomb(self, operators, coefficients, identity_shift=0., name=None):
        """See :meth:`CudaCsrOperator._assemble_lincomb` (this operator may be the first of the combination)."""
        return CudaCsrOperator._assemble_lincomb(self, operators, coefficients, identity_shift=identity_shift, name=name)

    def inverse(self):
        """The diagonal operator with the reciprocal weights (cached)."""
        inv = self.__dict__.get('_inverse')
        if inv is None:
            inv = CudaDiagonalOperator(1.0 / self._w[:max(self.space.local_dim, 1)] if self.space.local_dim
                                       else torch.zeros(0, dtype=torch.float64, device=self.space.device), self.space)
            object.__setattr__(self, '_inverse', inv)
        return inv

    def apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        return V.impl.inner(U.impl, V.ind, U.ind, weight_ptr=self._w.data_ptr())

    def pairwise_apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        assert len(V) == len(U)
        return V.impl.pairwise_inner(U.impl, V.ind, U.ind, weight_ptr=self._w.data_ptr())

    def assemble(self, mu=None):
        return self


class _ShardedCsr:
    """This rank's row block of a sparse matrix on the device, plus the halo plan.

    Columns owned by this rank are renumbered to ``[0, n_local)``; columns owned by other ranks
    ("ghost" DOFs) to ``n_local + position`` in the sorted list of ghost indices.  Before every
    product the ghost rows of the operand are fetched from their owners: each owner packs the rows a
    neighbour asked for (``pmc_gather_rows``) and sends them point-to-point (NCCL over NVLink).  This
    is the "MPI-aware operator" the reference leaves to external solvers
    (src/pymor/operators/mpi.py:22-26)."""

    def __init__(self, csr, space, rows_are_local=False):
        comm = space.comm
        lo, n = space.offset, space.local_dim
        hi = lo + n
        if rows_are_local:          # this rank's row block only (n_local x dim, GLOBAL column numbers)
            assert csr.shape == (n, space.dim)
            local = sps.csr_matrix(csr)
        else:                       # the global matrix: keep this rank's rows
            local = sps.csr_matrix(csr[lo:hi])
        local.sort_indices()
        # this rank's part of the matrix diagonal (Jacobi preconditioner of the device CG)
        self.diag_local = np.asarray(local.diagonal(k=lo) if local.shape[1] > lo else np.zeros(0), dtype=np.float64)
        cols = local.indices.astype(np.int64)
        mine = (cols >= lo) & (cols < hi)
        ghosts = np.unique(cols[~mine])
        new_cols = np.where(mine, cols - lo, n + np.searchsorted(ghosts, cols))
        assert n + len(ghosts) < 2 ** 31
        starts = np.concatenate((np.cumsum(space.local_dims) - space.local_dims, [space.dim]))
        owner = np.searchsorted(starts, ghosts, side='right') - 1
        need = {int(p): ghosts[owner == p] for p in np.unique(owner)}      # peer -> global rows I need
        all_need = comm.allgather_object(need)
        dev = space.device
        self.n_ghost = len(ghosts)
        self.recv = {}                                                     # peer -> (first ghost slot, count)
        pos = 0
        for p in sorted(need):
            self.recv[p] = (pos, len(need[p]))
            pos += len(need[p])
        self.send = {}                                                     # peer -> device int64 local rows
        for r, wanted in enumerate(all_need):
            rows = wanted.get(comm.rank)
            if r != comm.rank and rows is not None and len(rows):
                self.send[r] = torch.from_numpy(np.ascontiguousarray(rows - lo, dtype=np.int64)).to(dev)
        self.rowptr = torch.from_numpy(local.indptr.astype(np.int64)).to(dev)
        self.colidx = torch.from_numpy(new_cols.astype(np.int32)).to(dev)
        self.vals = torch.from_numpy(local.data.astype(np.float64)).to(dev)
        self.needs_exchange = comm.size > 1 and any(len(w) for d in all_need for w in d.values())

    @classmethod
    def from_device(cls, rowptr, colidx, vals):
        self = cls.__new__(cls)
        self.rowptr, self.colidx, self.vals = rowptr.contiguous(), colidx.contiguous(), vals.contiguous()
        self.n_ghost, self.recv, self.send, self.needs_exchange = 0, {}, {}, False
        self.diag_local = None
        return self

    def halo(self, space, ptr, cs, k):
        """Fetch the ghost rows of the panel ``(ptr, cs, k)``; returns the (n_ghost, k) buffer or None.
        Collective over the communicator when any rank has ghosts."""
        if not self.needs_exchange:
            return None
        ghost = torch.empty((max(self.n_ghost, 1), k), dtype=torch.float64, device=space.device)
        sends, recvs = {}, {}
        for p, idx in self.send.items():
            buf = torch.empty((idx.numel(), k), dtype=torch.float64, device=space.device)
            space.ctx.call('pmc_gather_rows', ptr, cs, k, space.local_dim, idx.data_ptr(), idx.numel(),
                           buf.data_ptr(), 0)
            sends[p] = buf
        for p, (first, count) in self.recv.items():
            recvs[p] = ghost[first:first + count]
        if space.device.type == 'cuda':
            # The packing kernels ran on the library's stream; NCCL orders its transfers after the work queued on
            # torch's CURRENT stream.  Normally those are the same stream (nothing to do); otherwise chain them with
            # an event -- the host never waits.
            cur = torch.cuda.current_stream(space.device)
            lib_stream = space.ctx.stream_ptr
            if lib_stream != cur.cuda_stream:
                cur.wait_stream(torch.cuda.ExternalStream(lib_stream, device=space.device))
        space.comm.exchange(sends, recvs)
        if space.device.type == 'cuda':
            cur = torch.cuda.current_stream(space.device)
            if space.ctx.stream_ptr != cur.cuda_stream:
                torch.cuda.ExternalStream(space.ctx.stream_ptr, device=space.device).wait_stream(cur)
        return ghost


class CudaCsrOperator(Operator):
    """Sparse matrix operator with the CSR arrays resident on the device.

    Parameters
    ----------
    matrix
        SciPy sparse matrix (any format; converted to CSR with int64 row pointers and int32 column
        indices) of shape ``(space.dim, space.dim)`` -- the GLOBAL matrix; every rank keeps its row
        block.  Couplings across shard boundaries are served by a halo exchange.
    space
        The :class:`CudaVectorSpace` (source and range).
    symmetric
        ``True`` / ``False``: the matrix is (not) symmetric; ``None``: checked on the host matrix on first use.
        Symmetric operators compute ``apply2(V, V)`` on the lower tile triangle only (half the flop), use
        ``apply`` for ``apply_adjoint`` and default to the device CG solver for ``apply_inverse``.

    Use :meth:`from_local_rows` when the global matrix should not exist on every rank (each rank passes its own
    row block), :meth:`from_device_csr` for CSR arrays that are already on the device.
    """

    linear = True

    def __init__(self, matrix, space, solver=None, name=None, symmetric=None):
        assert isinstance(space, CudaVectorSpace)
        self.matrix = matrix
        self.space = space
        self.name = name
        self.solver = solver
        self.symmetric = symmetric
        self.source = self.range = space
        self._dev = None
        self._dev_T = None

        if matrix is not None:
            assert matrix.shape == (space.dim, space.dim)
            self._dev = _ShardedCsr(sps.csr_matrix(matrix), space)

    @classmethod
    def from_local_rows(cls, local_rows, space, symmetric=False, solver=None, name=None):
        """From THIS RANK's row block only: a SciPy sparse matrix of shape ``(space.local_dim, space.dim)`` holding
        rows ``[space.offset, space.offset + space.local_dim)`` with GLOBAL column numbers.  No rank ever holds the
        global matrix; the ghost map (which remote DOFs each rank needs, who sends what) is derived from the column
        indices and agreed on once (one ``all_gather_object`` of the request lists).  ``apply_adjoint`` is available
        for ``symmetric=True`` only (the transpose of a row-distributed matrix would need a redistribution)."""
        op = cls(None, space, solver=solver, name=name, symmetric=bool(symmetric))
        object.__setattr__(op, '_dev', _ShardedCsr(local_rows, space, rows_are_local=True))
        return op

    def with_(self, **kwargs):
        """pyMOR's copy-with-changes (``ImmutableObject.with_`` re-runs ``__init__``): operators built from per-rank
        rows or device CSR arrays have no host matrix to rebuild from, so their device block is carried over."""
        new = super().with_(**kwargs)
        if new._dev is None and self._dev is not None and new.space == self.space:
            object.__setattr__(new, '_dev', self._dev)
            object.__setattr__(new, '_dev_T', self._dev_T)
        return new

    def is_symmetric(self):
        """Whether the matrix is symmetric: the ``symmetric`` argument, else a host check cached on first use
        (``False`` without a host matrix)."""
        if self.symmetric is not None:
            return bool(self.symmetric)
        if '_symmetric' not in self.__dict__:
            sym = False
            if self.matrix is not None and self.matrix.shape[0]:
                csr = sps.csr_matrix(self.matrix)
                asym = abs(csr - csr.T)
                sym = bool(asym.nnz == 0 or asym.max() <= 1e-14 * abs(csr).max())
            object.__setattr__(self, '_symmetric', sym)
        return self.__dict__['_symmetric']

    @classmethod
    def from_device_csr(cls, rowptr, colidx, vals, space, name=None, symmetric=False):
        """Adopt this rank's CSR block already on the device (int64 rowptr, int32 colidx with LOCAL
        column numbers, fp64 vals); no halo."""
        op = cls(None, space, name=name, symmetric=symmetric)
        assert rowptr.dtype == torch.int64 and colidx.dtype == torch.int32 and vals.dtype == torch.float64
        object.__setattr__(op, '_dev', _ShardedCsr.from_device(rowptr, colidx, vals))
        return op

    def _apply_csr(self, blk, U):
        if _is_complex(U):
            return _apply_by_parts(lambda W: self._apply_csr(blk, W), U)
        pu, csu, k, keep = U.impl._panel(U.ind)
        out = CudaVectorArrayImpl.allocate(self.space, k)
        n = self.space.local_dim
        ghost = blk.halo(self.space, pu, csu, k) if k else None
        if k and n:
            self.space.ctx.call('pmc_csr_apply', n, blk.rowptr.data_ptr(), blk.colidx.data_ptr(),
                                blk.vals.data_ptr(), pu, csu, ghost.data_ptr() if ghost is not None else 0,
                                out._ptr(0), out.ld, k, 0)
        del keep
        return CudaVectorArray(self.space, out)

    def apply(self, U, mu=None):
        assert U in self.source
        return self._apply_csr(self._dev, U)

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        if self.is_symmetric():
            return self._apply_csr(self._dev, V)
        if self._dev_T is None:
            if self.matrix is None:
                raise NotImplementedError('apply_adjoint needs the host matrix to build the transpose')
            self._dev_T = _ShardedCsr(sps.csr_matrix(self.matrix.T), self.space)
        return self._apply_csr(self._dev_T, V)

    def apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        if _is_complex(V) or _is_complex(U):
            return V.inner(self.apply(U))
        pv, csv, kv, keep_v = V.impl._panel(V.ind)
        pu, csu, ku, keep_u = U.impl._panel(U.ind)
        if kv == 0 or ku == 0:
            return np.zeros((kv, ku), order='F')
        sp, blk = self.space, self._dev
        ghost = blk.halo(sp, pu, csu, ku)
        out = torch.zeros(kv * ku, dtype=torch.float64, device=sp.device)   # accumulated on the device
        # V^T (M V) with symmetric M is symmetric: contract the lower tile triangle only (SYRK flop count)
        flags = sp.kernel_flags
        if (pv, csv, kv) == (pu, csu, ku) and kv > 128 and self.is_symmetric():
            flags |= PMC_LOWER_ONLY
        if sp.local_dim:
            sp.ctx.call('pmc_csr_gram', sp.local_dim, blk.rowptr.data_ptr(), blk.colidx.data_ptr(),
                        blk.vals.data_ptr(), pv, csv, kv, pu, csu, ku,
                        ghost.data_ptr() if ghost is not None else 0, out.data_ptr(), kv, flags)
        if sp.comm.size > 1:
            sp.comm.allreduce_sum_(out)
            if (flags & PMC_LOWER_ONLY) and sp.comm.size > 2:    # the collective may break the mirrored halves' equality
                sp.ctx.call('pmc_mirror_lower', out.data_ptr(), kv, kv, 0)
        del keep_v, keep_u
        return out.to('cpu').numpy().reshape((kv, ku), order='F')

    def pairwise_apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        assert len(V) == len(U)
        if _is_complex(V) or _is_complex(U):
            return V.pairwise_inner(self.apply(U))
        # fused: one pass over M, U and V; the n x k product M U (what operators/interface.py:111-143 materialises)
        # is never written
        pv, csv, kv, keep_v = V.impl._panel(V.ind)
        pu, csu, ku, keep_u = U.impl._panel(U.ind)
        if kv == 0:
            return np.zeros(0)
        sp, blk = self.space, self._dev
        ghost = blk.halo(sp, pu, csu, ku)
        out, flags = sp.reduction_target(kv)
        sp.ctx.call('pmc_csr_pairwise', sp.local_dim, blk.rowptr.data_ptr(), blk.colidx.data_ptr(),
                    blk.vals.data_ptr(), pv, csv, pu, csu, ghost.data_ptr() if ghost is not None else 0, kv,
                    out.data_ptr(), flags)
        del keep_v, keep_u
        return sp.reduction_result(out, kv)

    def _assemble_lincomb(self, operators, coefficients, identity_shift=0., name=None):
        """pyMOR's hook for ``LincombOperator.assemble`` (reference counterpart: operators/numpy.py:243-287): a real
        linear combination of CSR / diagonal device operators on this space with host matrices becomes ONE
        ``CudaCsrOperator`` (assembled with SciPy on the host, uploaded once), so that ``fom.solve(mu)`` of a device
        model is a single device solve instead of a chain of operator applications."""
        mats = []
        for op in operators:
            if isinstance(op, CudaCsrOperator) and op.matrix is not None and op.space == self.space:
                mats.append(sps.csr_matrix(op.matrix))
            elif isinstance(op, CudaDiagonalOperator) and op.space == self.space and not isinstance(op.weights, torch.Tensor):
                mats.append(sps.diags(np.asarray(op.weights, dtype=np.float64).reshape(-1)).tocsr())
            else:
                return None
        coeffs = list(coefficients) + [identity_shift]
        if any(np.iscomplexobj(c) and np.imag(c) != 0 for c in coeffs):
            return None
        total = None
        for m, c in zip(mats, coefficients):
            term = m * float(np.real(c))
            total = term if total is None else total + term
        if identity_shift != 0:
            total = total + sps.eye(total.shape[0]) * float(np.real(identity_shift))
        solver = self.solver if isinstance(self, CudaCsrOperator) else None
        return CudaCsrOperator(total.tocsr(), self.space, solver=solver, name=name)

    def default_solver(self):
        """Solver used by ``apply_inverse`` when none was given: conjugate gradients on the device for symmetric
        matrices (mass / stiffness / energy products -- Riesz representatives ``product.apply_inverse(residual)``
        then stay on the device), ``None`` (pyMOR's default) otherwise.  Symmetry is checked on first use."""
        if '_default_solver' not in self.__dict__:
            object.__setattr__(self, '_default_solver', CgSolver() if self.is_symmetric() else None)
        return self.__dict__['_default_solver']

    def apply_inverse(self, V, mu=None, initial_guess=None, return_info=False, solver=None):
        solver = solver or self.solver or self.default_solver()
        return super().apply_inverse(V, mu=mu, initial_guess=initial_guess, return_info=return_info, solver=solver)

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, return_info=False, solver=None):
        solver = solver or self.solver or self.default_solver()
        return super().apply_inverse_adjoint(U, mu=mu, initial_guess=initial_guess, return_info=return_info,
                                             solver=solver)

    def jacobi_preconditioner(self):
        """Inverse of the matrix diagonal as a :class:`CudaDiagonalOperator` (``None`` with a zero on the diagonal or
        for operators adopted from device CSR arrays); cached.  Collective on sharded spaces."""
        if '_jacobi' not in self.__dict__:
            jac = None
            d = self._dev.diag_local if self._dev is not None else None
            if d is not None and len(d) == self.space.local_dim:
                ok = bool(np.all(d != 0))
                if self.space.comm.size > 1:                       # all ranks must take the same branch
                    ok = all(self.space.comm.allgather_object(ok))
                if ok and self.space.dim:
                    jac = CudaDiagonalOperator(torch.from_numpy(1.0 / d).to(self.space.device), self.space)
            object.__setattr__(self, '_jacobi', jac)
        return self.__dict__['_jacobi']

    def assemble(self, mu=None):
        return self
