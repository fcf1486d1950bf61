"""Complex vector arrays as a pair of real device arrays (real part, imaginary part).

The hot path of this project is real fp64; complex data only has to be CORRECT, because pyMOR's
VectorArray contract (and its hypothesis test-suite) includes complex arrays and in-place promotion
(``v.scal(1j)``, reference: src/pymor/vectorarrays/numpy.py:85-100).  Every complex operation is
composed from the real kernels acting on the two parts:

    <a, b>       = (ar.br + ai.bi) + i (ar.bi - ai.br)          (conjugate-linear in a, numpy.py:155-163)
    y += alpha x : yr += alr xr - ali xi ;  yi += ali xr + alr xi
    A C          = (Ar Cr - Ai Ci) + i (Ar Ci + Ai Cr)

``_im is None`` stands for an all-zero imaginary part (a real array viewed as complex, no storage).
"""
from __future__ import annotations

from numbers import Number

import numpy as np

from pymor_b200.interface import VectorArrayImpl


def _parts(alpha, count):
    """Real and imaginary parts of a scalar / per-vector coefficient as float64 arrays (or floats)."""
    a = np.asarray(alpha)
    if a.ndim == 0:
        c = complex(a)
        return c.real, c.imag
    a = a.astype(np.complex128).reshape(-1)
    assert a.size == count
    return np.ascontiguousarray(a.real), np.ascontiguousarray(a.imag)


def _nonzero(v):
    return bool(np.any(np.asarray(v) != 0))


class CudaComplexVectorArrayImpl(VectorArrayImpl):
    """``_re`` / ``_im``: real :class:`~pymor_b200.vectorarray.CudaVectorArrayImpl` of equal length."""

    def __init__(self, space, re, im):
        self.space = space
        self._re = re
        self._im = im

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def wrap(impl):
        """View any backend impl as complex (no copy)."""
        if type(impl) is CudaComplexVectorArrayImpl:
            return impl
        return CudaComplexVectorArrayImpl(impl.space, impl, None)

    def _ensure_im(self):
        if self._im is None:
            from pymor_b200.vectorarray import CudaVectorArrayImpl
            self._im = CudaVectorArrayImpl.allocate(self.space, len(self._re), self._re._buf.shape[0], zero=True)
        return self._im

    def __len__(self):
        return len(self._re)

    @property
    def _len(self):
        return len(self._re)

    # -- storage -------------------------------------------------------------------------------
    def copy(self, deep, ind):
        return CudaComplexVectorArrayImpl(self.space, self._re.copy(deep, ind),
                                          None if self._im is None else self._im.copy(deep, ind))

    def real(self, ind):
        return self._re.copy(True, ind)

    def imag(self, ind):
        if self._im is None:
            return self._re.imag(ind)
        return self._im.copy(True, ind)

    def conj(self, ind):
        out = self.copy(True, ind)
        if out._im is not None:
            out._im.scal(-1.0, None)
        return out

    def to_numpy(self, ensure_copy, ind):
        re = self._re.to_numpy(True, ind)
        if self._im is None:
            return re.astype(np.complex128)
        return re + 1j * self._im.to_numpy(True, ind)

    def delete(self, ind):
        self._re.delete(ind)
        if self._im is not None:
            self._im.delete(ind)

    def append(self, other, remove_from_other, oind):
        o = self.wrap(other)
        count = o._re.len_ind(oind)
        if count == 0:
            return
        if o is self:
            o = self.copy(True, oind)
            oind = None
        im = self._ensure_im()
        self._re.append(o._re, False, oind)
        if o._im is None:
            from pymor_b200.vectorarray import CudaVectorArrayImpl
            im.append(CudaVectorArrayImpl.allocate(self.space, count, zero=True), False, None)
        else:
            im.append(o._im, False, oind)
        if remove_from_other:
            other.delete(oind)

    # -- BLAS-1 --------------------------------------------------------------------------------
    def scal(self, alpha, ind):
        ar, ai = _parts(alpha, self._re.len_ind(ind))
        if not _nonzero(ai):
            self._re.scal(ar, ind)
            if self._im is not None:
                self._im.scal(ar, ind)
            return
        im = self._ensure_im()
        old_re = self._re.copy(True, ind)
        self._re.scal(ar, ind)
        self._re.axpy(-ai, im, ind, ind)
        im.scal(ar, ind)
        im.axpy(ai, old_re, ind, None)

    def axpy(self, alpha, x, ind, xind):
        count = self._re.len_ind(ind)
        ar, ai = _parts(alpha, count)
        x = self.wrap(x)
        if x._re is self._re or (x._im is not None and x._im is self._im):
            x = x.copy(True, xind)          # right-hand side first (numpy.py:138)
            xind = None
        complex_alpha = _nonzero(ai)
        self._re.axpy(ar, x._re, ind, xind)
        if x._im is not None and complex_alpha:
            self._re.axpy(-ai, x._im, ind, xind)
        if x._im is not None or complex_alpha:
            im = self._ensure_im()
            if x._im is not None:
                im.axpy(ar, x._im, ind, xind)
            if complex_alpha:
                im.axpy(ai, x._re, ind, xind)

    # -- reductions ----------------------------------------------------------------------------
    def inner(self, other, ind, oind, weight_ptr=0):
        o = self.wrap(other)
        rr = self._re.inner(o._re, ind, oind, weight_ptr=weight_ptr)
        out = rr.astype(np.complex128)
        if self._im is not None and o._im is not None:
            out += self._im.inner(o._im, ind, oind, weight_ptr=weight_ptr)
        if o._im is not None:
            out += 1j * self._re.inner(o._im, ind, oind, weight_ptr=weight_ptr)
        if self._im is not None:
            out -= 1j * self._im.inner(o._re, ind, oind, weight_ptr=weight_ptr)
        return np.asfortranarray(out)

    def gramian(self, ind):
        g = self.inner(self, ind, ind)
        return (g + g.conj().T) / 2 if g.size else g       # exactly Hermitian

    def pairwise_inner(self, other, ind, oind, weight_ptr=0):
        o = self.wrap(other)
        out = self._re.pairwise_inner(o._re, ind, oind, weight_ptr=weight_ptr).astype(np.complex128)
        if self._im is not None and o._im is not None:
            out += self._im.pairwise_inner(o._im, ind, oind, weight_ptr=weight_ptr)
        if o._im is not None:
            out += 1j * self._re.pairwise_inner(o._im, ind, oind, weight_ptr=weight_ptr)
        if self._im is not None:
            out -= 1j * self._im.pairwise_inner(o._re, ind, oind, weight_ptr=weight_ptr)
        return out

    def norm2(self, ind, weight_ptr=0):
        out = self._re.norm2(ind, weight_ptr=weight_ptr)
        if self._im is not None:
            out = out + self._im.norm2(ind, weight_ptr=weight_ptr)
        return out

    def norm(self, ind):
        return np.sqrt(self.norm2(ind))

    def lincomb(self, coefficients, ind):
        c = np.asarray(coefficients)
        cr = np.ascontiguousarray(c.real, dtype=np.float64)
        ci = np.ascontiguousarray(c.imag, dtype=np.float64) if np.iscomplexobj(c) else None
        has_ci = ci is not None and _nonzero(ci)
        re = self._re.lincomb(cr, ind)
        im = None
        if self._im is not None and has_ci:
            re.axpy(-1.0, self._im.lincomb(ci, ind), None, None)
        if self._im is not None:
            im = self._im.lincomb(cr, ind)
        if has_ci:
            t = self._re.lincomb(ci, ind)
            if im is None:
                im = t
            else:
                im.axpy(1.0, t, None, None)
        return CudaComplexVectorArrayImpl(self.space, re, im)

    def dofs(self, dof_indices, ind):
        out = self._re.dofs(dof_indices, ind).astype(np.complex128)
        if self._im is not None:
            out += 1j * self._im.dofs(dof_indices, ind)
        return np.asfortranarray(out)

    def amax(self, ind):
        if self._im is None:
            return self._re.amax(ind)
        # |z|^2 = re^2 + im^2, built column by column with the elementwise kernel (w .* x with w = x)
        from pymor_b200.vectorarray import CudaVectorArrayImpl
        re, im = self._re.copy(True, ind), self._im.copy(True, ind)
        k, n, ctx = len(re), re.n, self.space.ctx
        sq = CudaVectorArrayImpl.allocate(self.space, k)
        sq2 = CudaVectorArrayImpl.allocate(self.space, k)
        for c in range(k if n else 0):
            ctx.call('pmc_diag_apply', re._ptr(c), re._ptr(c), re.ld, sq._ptr(c), sq.ld, 1, n, 0)
            ctx.call('pmc_diag_apply', im._ptr(c), im._ptr(c), im.ld, sq2._ptr(c), sq2.ld, 1, n, 0)
        sq.axpy(1.0, sq2, None, None)
        idx, val = sq.amax(None)
        return idx, np.sqrt(val)
