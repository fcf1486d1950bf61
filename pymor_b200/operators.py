"""Product / system operators on a :class:`CudaVectorSpace`.

pyMOR evaluates ``product.apply2(V, U)`` as ``V.inner(product.apply(U))`` and
``pairwise_apply2`` as ``V.pairwise_inner(product.apply(U))`` (reference:
src/pymor/operators/interface.py:79-143), materialising the n x k array ``W U`` first
(src/pymor/operators/numpy.py:232-237).  The operators here override both with fused device
kernels so that weighted Gramians, norms and Galerkin projections never allocate that temporary:

* :class:`CudaDiagonalOperator` -- diagonal (lumped-mass / FV L2) products: the weight is applied to
  the B-fragment inside the TMA + DMMA Gram kernel resp. inside the streaming reduction.
* :class:`CudaCsrOperator` -- sparse (FEM mass/stiffness) matrices in CSR: ``apply2`` forms row panels
  of ``M U`` in an L2-sized scratch and contracts them at once.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps
import torch

from pymor_b200._lib import PMC_SYNC
from pymor_b200.interface import Operator
from pymor_b200.vectorarray import CudaVectorArray, CudaVectorArrayImpl, CudaVectorSpace, _same_index


def _same_columns(V, U):
    return V.impl is U.impl and _same_index(V.ind, U.ind)


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

    def __init__(self, weights, space, name=None):
        assert isinstance(space, CudaVectorSpace)
        self.weights = weights
        self.space = space
        self.name = name
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
        pu, csu, k, keep = U.impl._panel(U.ind)
        out = CudaVectorArrayImpl.allocate(self.space, k)
        if k and self.space.local_dim:
            self.space.ctx.call('pmc_diag_apply', self._w.data_ptr(), pu, csu, out._ptr(0), out.ld, k,
                                self.space.local_dim, 0)
        del keep
        return CudaVectorArray(self.space, out)

    def apply_adjoint(self, V, mu=None):
        return self.apply(V, mu=mu)

    def apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        return V.impl.inner(U.impl, V.ind, U.ind, weight_ptr=self._w.data_ptr())

    def pairwise_apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        assert len(V) == len(U)
        return V.impl.pairwise_inner(U.impl, V.ind, U.ind, weight_ptr=self._w.data_ptr())

    def assemble(self, mu=None):
        return self


class CudaCsrOperator(Operator):
    """Sparse matrix operator with the CSR arrays resident on the device.

    Parameters
    ----------
    matrix
        SciPy sparse matrix (any format; converted to CSR with int64 row pointers and int32 column
        indices) of shape ``(space.dim, space.dim)``.
    space
        The :class:`CudaVectorSpace` (source and range).  Sharded spaces (more than one rank) are
        supported for block-diagonal matrices only (no halo exchange yet).
    """

    linear = True

    def __init__(self, matrix, space, name=None):
        assert isinstance(space, CudaVectorSpace)
        self.matrix = matrix
        self.space = space
        self.name = name
        self.source = self.range = space
        self._dev = None
        self._dev_T = None
        if matrix is not None:
            assert matrix.shape == (space.dim, space.dim)
            self._dev = self._upload(sps.csr_matrix(matrix))

    @classmethod
    def from_device_csr(cls, rowptr, colidx, vals, space, name=None):
        """Adopt CSR arrays already on the device (int64 rowptr, int32 colidx, fp64 vals)."""
        op = cls(None, space, name=name)
        assert rowptr.dtype == torch.int64 and colidx.dtype == torch.int32 and vals.dtype == torch.float64
        object.__setattr__(op, '_dev', (rowptr.contiguous(), colidx.contiguous(), vals.contiguous()))
        return op

    def _upload(self, csr):
        space = self.space
        lo, hi = space.offset, space.offset + space.local_dim
        local = csr[lo:hi]
        if space.comm.size > 1:
            cols = local.indices
            if cols.size and (cols.min() < lo or cols.max() >= hi):
                raise NotImplementedError('sharded CSR operators need a halo exchange (block-diagonal only)')
            local = sps.csr_matrix((local.data, local.indices - lo, local.indptr),
                                   shape=(space.local_dim, space.local_dim))
        local.sort_indices()
        dev = space.device
        return (torch.from_numpy(local.indptr.astype(np.int64)).to(dev),
                torch.from_numpy(local.indices.astype(np.int32)).to(dev),
                torch.from_numpy(local.data.astype(np.float64)).to(dev))

    def _apply_csr(self, csr_dev, U):
        pu, csu, k, keep = U.impl._panel(U.ind)
        out = CudaVectorArrayImpl.allocate(self.space, k)
        n = self.space.local_dim
        if k and n:
            rowptr, colidx, vals = csr_dev
            self.space.ctx.call('pmc_csr_apply', n, rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr(),
                                pu, csu, out._ptr(0), out.ld, k, 0)
        del keep
        return CudaVectorArray(self.space, out)

    def apply(self, U, mu=None):
        assert U in self.source
        return self._apply_csr(self._dev, U)

    def apply_adjoint(self, V, mu=None):
        assert V in self.range
        if self._dev_T is None:
            if self.matrix is None:
                raise NotImplementedError('apply_adjoint needs the host matrix to build the transpose')
            self._dev_T = self._upload(sps.csr_matrix(self.matrix.T))
        return self._apply_csr(self._dev_T, V)

    def apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        pv, csv, kv, keep_v = V.impl._panel(V.ind)
        pu, csu, ku, keep_u = U.impl._panel(U.ind)
        if kv == 0 or ku == 0:
            return np.zeros((kv, ku), order='F')
        sp = self.space
        out = torch.empty(kv * ku, dtype=torch.float64, device=sp.device)   # accumulated on the device
        rowptr, colidx, vals = self._dev
        sp.ctx.call('pmc_csr_gram', sp.local_dim, rowptr.data_ptr(), colidx.data_ptr(), vals.data_ptr(),
                    pv, csv, kv, pu, csu, ku, out.data_ptr(), kv, sp.kernel_flags)
        if sp.comm.size > 1:
            sp.comm.allreduce_sum_(out)
        del keep_v, keep_u
        return out.to('cpu').numpy().reshape((kv, ku), order='F')

    def pairwise_apply2(self, V, U, mu=None):
        assert V in self.range and U in self.source
        assert len(V) == len(U)
        return V.pairwise_inner(self.apply(U))

    def assemble(self, mu=None):
        return self
