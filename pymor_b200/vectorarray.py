"""CudaVectorSpace / CudaVectorArray: device-resident fp64 snapshot arrays behind pyMOR's
VectorArray interface.

Storage (mirrors ``NumpyVectorArrayImpl``'s Fortran-ordered ``(dim, len)`` ndarray, reference:
src/pymor/vectorarrays/numpy.py:16-18, :264): one torch CUDA tensor of shape ``(capacity, ld)``,
row ``i`` = vector ``i``, the ``dim`` DOFs of a vector contiguous, ``ld`` = ``dim`` rounded up to 16
doubles (128 B) so that every vector is TMA- and 128-bit-load aligned.  torch is used for allocation,
stream and DLPack interop only; all arithmetic runs in libpymor_b200.so (hand-written sm_100a
kernels) through the C ABI of include/pymor_b200.h.  There is no CPU fallback.

With a communicator the rows are sharded over ranks (one process per GPU) exactly like
``MPIVectorArrayAutoComm`` (src/pymor/vectorarrays/mpi.py:316-445).

Complex numbers: the performance path is real fp64.  Complex arrays (pyMOR's contract includes them and
in-place promotion such as ``v.scal(1j)``) are a pair of real arrays, every operation composed from the
real kernels -- see pymor_b200/complexarray.py.
"""
from __future__ import annotations

import os
from numbers import Number

import numpy as np
import torch

from pymor_b200 import _lib
from pymor_b200._lib import PMC_FORCE_SIMPLE, PMC_LOCAL_ONLY, PMC_SYMMETRIC, PMC_SYNC
from pymor_b200.dist import SingleRank, merge_amax, partition
from pymor_b200.interface import VectorArray, VectorArrayImpl, VectorSpace, create_random_values

_ALIGN = 16  # doubles
_FLOAT_TYPES = (float, np.float64, int)
_HOST_REDUCE_MAX = 2048  # sharded reductions of at most this many doubles are summed on the host
_STREAMED_UPLOAD_BYTES = 256 << 20   # pinned shards at least this large are uploaded on a side stream
_UPLOAD_BLOCKS = 8                   # ... in this many column blocks
_SKINNY = 4  # inner products with at most this many vectors on one side use the streaming kernel


class _RawTarget:
    """A bare device-accessible address standing in for a tensor as reduction output."""
    __slots__ = ('ptr',)

    def __init__(self, ptr):
        self.ptr = ptr

    def data_ptr(self):
        return self.ptr


def _round_up(a, b):
    return (a + b - 1) // b * b


def _progression(ind, length):
    """(start, step, count) if ``ind`` addresses an arithmetic progression of columns, else None."""
    if ind is None:
        return 0, 1, length
    if type(ind) is slice:
        r = range(*ind.indices(length))
        return r.start, r.step, len(r)
    count = len(ind)
    if count == 0:
        return 0, 1, 0
    first = int(ind[0])
    if count == 1:
        return first, 1, 1
    step = int(ind[1]) - first
    prev = int(ind[1])
    for v in ind[2:]:
        if int(v) - prev != step:
            return None
        prev = int(v)
    return first, step, count


def _columns(ind, length):
    """The addressed columns as a ``range`` or list."""
    if ind is None:
        return range(length)
    if type(ind) is slice:
        return range(*ind.indices(length))
    return [int(i) for i in ind]


def _overlap(a, b):
    if len(a) == 0 or len(b) == 0:
        return False
    if len(a) == 1 and len(b) == 1:
        return a[0] == b[0]
    return not set(a).isdisjoint(b)


def _runs(cols):
    """Split a column list into maximal arithmetic progressions: (offset, start, step, count)."""
    out, i, n = [], 0, len(cols)
    while i < n:
        if i + 1 == n:
            out.append((i, cols[i], 1, 1))
            break
        step = cols[i + 1] - cols[i]
        j = i + 1
        while j + 1 < n and cols[j + 1] - cols[j] == step:
            j += 1
        if step == 0:           # never merge repeated columns into one writable run
            out.append((i, cols[i], 1, 1))
            i += 1
            continue
        out.append((i, cols[i], step, j - i + 1))
        i = j + 1
    return out


def _consecutive_runs(cols):
    """Maximal runs of consecutive integers in the sorted list ``cols``: (start, count)."""
    out = []
    for c in cols:
        if out and out[-1][0] + out[-1][1] == c:
            out[-1][1] += 1
        else:
            out.append([c, 1])
    return out


def _is_complex_value(alpha):
    """True for a Python/NumPy complex scalar or complex ndarray (whatever its imaginary part): the
    reference promotes on the coefficient's dtype, not on its value (numpy.py:90-95)."""
    return isinstance(alpha, (complex, np.complexfloating)) or (type(alpha) is np.ndarray and np.iscomplexobj(alpha))


def _real_scalars(alpha, count):
    """alpha (Number | ndarray (count,)) -> contiguous float64 host array of 1 or ``count`` values."""
    a = np.asarray(alpha)
    if np.iscomplexobj(a):
        if np.any(a.imag != 0):
            raise NotImplementedError('CudaVectorArray is real fp64; complex coefficients are not supported')
        a = a.real
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert a.size in (1, count)
    return a


class CudaVectorArrayImpl(VectorArrayImpl):
    """Backend object owning one device buffer (reference counterpart: numpy.py:14-211)."""

    def __init__(self, space, buf, length):
        self.space = space
        self._buf = buf
        self._len = int(length)
        self._arrivals = None        # [(col_begin, col_end, event)] of a host -> device upload in flight
        self._arrival_src = None

    # -- storage helpers -------------------------------------------------------------------
    @classmethod
    def allocate(cls, space, count, reserve=0, zero=False):
        cap = max(int(count), int(reserve), 1)
        alloc = torch.zeros if zero else torch.empty
        buf = alloc((cap, space.ld), dtype=torch.float64, device=space.device)
        return cls(space, buf, count)

    def _promote(self):
        """In-place promotion to complex (the reference does `self._array = self._array.astype(...)`,
        numpy.py:90-95): this object becomes a CudaComplexVectorArrayImpl whose real part adopts the
        buffer; views and copy-on-write references keep pointing at the same impl object."""
        from pymor_b200.complexarray import CudaComplexVectorArrayImpl
        self.space.ctx.flush()
        re = CudaVectorArrayImpl(self.space, self._buf_t, self._len)
        for name in ('_buf_t', '_base', '_ld', '_len', '_arrivals', '_arrival_src'):
            self.__dict__.pop(name, None)
        self.__class__ = CudaComplexVectorArrayImpl
        self._re, self._im = re, None
        return self

    # the buffer's address and leading dimension are cached: they are read on every kernel call
    @property
    def _buf(self):
        return self._buf_t

    @_buf.setter
    def _buf(self, t):
        self._buf_t = t
        self._base = t.data_ptr()
        self._ld = t.stride(0)

    @property
    def ld(self):
        return self._ld

    @property
    def n(self):
        return self.space.local_dim

    def __len__(self):
        return self._len

    def _ptr(self, col=0):
        return self._base + col * self._ld * 8

    def _panel(self, ind):
        """(ptr, column stride, count, keepalive) of a READ-ONLY panel; arbitrary index lists are
        gathered into a temporary (numpy.py:21 fancy indexing does the same)."""
        self.space.ctx.flush()
        prog = _progression(ind, self._len)
        if prog is not None:
            start, step, count = prog
            return self._ptr(start) if count else self._ptr(0), step * self.ld, count, self
        tmp = self._gathered(_columns(ind, self._len))
        return tmp._ptr(0), tmp.ld, tmp._len, tmp

    def _gathered(self, cols, reserve=0):
        self.space.ctx.flush()
        new = CudaVectorArrayImpl.allocate(self.space, len(cols), reserve)
        if len(cols) and self.n:
            idx = np.ascontiguousarray(cols, dtype=np.int64)
            self.space.ctx.call('pmc_gather_cols', new._ptr(0), new.ld, self._ptr(0), self.ld,
                                idx.ctypes.data, len(cols), self.n, 0)
        return new

    def _copied(self, ind, reserve=0):
        self.space.ctx.flush()
        prog = _progression(ind, self._len)
        if prog is None:
            return self._gathered(_columns(ind, self._len), reserve)
        start, step, count = prog
        new = CudaVectorArrayImpl.allocate(self.space, count, reserve)
        if count and self.n:
            self.space.ctx.call('pmc_copy', new._ptr(0), new.ld, self._ptr(start), step * self.ld,
                                count, self.n, 0)
        return new

    # -- VectorArrayImpl interface ------------------------------------------------------------
    def copy(self, deep, ind):
        return self._copied(ind)

    def real(self, ind):
        return self._copied(ind)

    def conj(self, ind):
        return self._copied(ind)

    def imag(self, ind):
        return CudaVectorArrayImpl.allocate(self.space, self.len_ind(ind), zero=True)

    def to_numpy(self, ensure_copy, ind):
        self.space.ctx.flush()
        prog = _progression(ind, self._len)
        if prog is not None and prog[1] == 1:
            block = self._buf[prog[0]:prog[0] + prog[2], :self.n]
        else:
            tmp = self._copied(ind)
            block = tmp._buf[:tmp._len, :self.n]
        comm = self.space.comm
        if comm.size > 1:
            nmax = int(self.space.local_dims.max())
            padded = torch.zeros((block.shape[0], nmax), dtype=torch.float64, device=self.space.device)
            padded[:, :self.n] = block
            parts = comm.allgather(padded)
            block = torch.cat([p[:, :int(d)] for p, d in zip(parts, self.space.local_dims)], dim=1)
        host = block.to('cpu').contiguous().numpy()      # (k, dim) C-order == (dim, k) F-order
        return host.T

    def delete(self, ind):
        self.space.ctx.flush()
        if ind is None:
            self._len = 0
            return
        # `ind` arrives un-normalised here (interface.py:252-259): int, slice or list, negatives allowed
        if isinstance(ind, Number):
            ind = [int(ind)]
        drop = {i % self._len for i in _columns(ind, self._len)} if self._len else set()
        keep = [c for c in range(self._len) if c not in drop]
        ctx, pos = self.space.ctx, 0
        # in-place compaction: one launch per run of kept columns (the kernel walks the columns of a run in
        # increasing order row by row, so overlapping source / destination ranges are safe)
        for start, count in _consecutive_runs(keep):
            shift = start - pos
            if shift and self.n:
                ctx.call('pmc_shift_cols', self._ptr(0), self.ld, start, count, shift, self.n, 0)
            pos += count
        self._len = len(keep)

    def append(self, other, remove_from_other, oind):
        if type(other) is not CudaVectorArrayImpl:            # complex data arrives: promote in place
            return self._promote().append(other, remove_from_other, oind)
        self.space.ctx.flush()
        k_other = other.len_ind(oind)
        if k_other == 0:
            return
        if (remove_from_other and self._len == 0 and other is not self
                and _progression(oind, other._len) == (0, 1, other._len)
                and other._buf.shape[0] >= self._buf.shape[0] and other.ld == self.ld):
            # moving a whole array into an empty one: adopt its buffer instead of copying
            # (`remove_from_other` exists for exactly this, interface.py:279-282)
            self._buf, other._buf = other._buf, self._buf
            self._len, other._len = other._len, 0
            return
        if other is self:
            src = self._copied(oind)
            ptr, cs, keep = src._ptr(0), src.ld, src
        else:
            ptr, cs, _, keep = other._panel(oind)
        need = self._len + k_other
        if need > self._buf.shape[0]:
            cap = max(need, int(self._buf.shape[0] * 1.5) + 1)
            grown = torch.empty((cap, self.ld), dtype=torch.float64, device=self.space.device)
            if self._len and self.n:
                self.space.ctx.call('pmc_copy', grown.data_ptr(), self.ld, self._ptr(0), self.ld,
                                    self._len, self.n, 0)
            self._buf = grown
        if self.n:
            self.space.ctx.call('pmc_copy', self._ptr(self._len), self.ld, ptr, cs, k_other, self.n, 0)
        del keep
        self._len = need
        if remove_from_other:
            other.delete(oind)

    def scal(self, alpha, ind):
        if _is_complex_value(alpha):
            return self._promote().scal(alpha, ind)
        self.space.ctx.flush()
        cols = _columns(ind, self._len)
        if len(cols) == 0 or self.n == 0:
            _real_scalars(alpha, len(cols))
            return
        a = _real_scalars(alpha, len(cols))
        ctx = self.space.ctx
        for off, start, step, count in self._write_runs(ind, cols):
            if a.size == 1:
                ctx.call('pmc_scal', self._ptr(start), step * self.ld, count, self.n, a.ctypes.data, 1, 0)
            else:
                ctx.call('pmc_scal', self._ptr(start), step * self.ld, count, self.n,
                         a.ctypes.data + 8 * off, count, 0)

    def _write_runs(self, ind, cols):
        prog = _progression(ind, self._len)
        if prog is not None and (prog[1] != 0 or prog[2] <= 1):
            return [(0, prog[0], prog[1] if prog[2] > 1 else 1, prog[2])]
        return _runs(list(cols))

    def axpy(self, alpha, x, ind, xind):
        # fast path of the Gram-Schmidt inner step: one vector += scalar * another vector, both
        # addressed by the unit slices ``A[i]`` / ``A[j]`` produce (interface.py:219-223)
        if type(x) is not CudaVectorArrayImpl or _is_complex_value(alpha):
            return self._promote().axpy(alpha, x, ind, xind)
        if (type(ind) is slice and type(xind) is slice and type(alpha) in _FLOAT_TYPES
                and ind.step in (None, 1) and xind.step in (None, 1)
                and ind.stop is not None and xind.stop is not None
                and ind.stop - ind.start == 1 and xind.stop - xind.start == 1
                and 0 <= ind.start < self._len and 0 <= xind.start < x._len and self.n
                and not (x._buf is self._buf and ind.start == xind.start)):
            ctx = self.space.ctx
            ctx.flush()                     # also orders the stream after uploads in flight
            if ctx.fuse_axpy:
                # deferred: merged into the next dot / norm that reads this vector (pmc_axpy_dot)
                ctx.pending = (self, ind.start, float(alpha), x, xind.start)
            else:
                ctx.scalar[0] = alpha
                ctx.call('pmc_axpy', self._ptr(ind.start), self.ld, x._ptr(xind.start), x.ld, 1, self.n,
                         ctx.scalar_ptr, 1, 0)
            return
        self.space.ctx.flush()
        cols = _columns(ind, self._len)
        k = len(cols)
        a = _real_scalars(alpha, k)
        if k == 0 or self.n == 0:
            return
        xcols = _columns(xind, x._len)
        if x._buf is self._buf and _overlap(cols, xcols):
            # NumPy evaluates the right-hand side first (numpy.py:138): snapshot aliased operands
            x = x._copied(xind)
            xind = None
        xptr, xcs, kx, keep = x._panel(xind)
        if kx == 1 and k > 1:
            xcs = 0
        ctx = self.space.ctx
        for off, start, step, count in self._write_runs(ind, cols):
            xp = xptr + 8 * off * xcs
            if a.size == 1:
                ctx.call('pmc_axpy', self._ptr(start), step * self.ld, xp, xcs, count, self.n,
                         a.ctypes.data, 1, 0)
            else:
                ctx.call('pmc_axpy', self._ptr(start), step * self.ld, xp, xcs, count, self.n,
                         a.ctypes.data + 8 * off, count, 0)
        del keep

    def inner(self, other, ind, oind, weight_ptr=0):
        if type(other) is not CudaVectorArrayImpl:
            from pymor_b200.complexarray import CudaComplexVectorArrayImpl
            return CudaComplexVectorArrayImpl.wrap(self).inner(other, ind, oind, weight_ptr=weight_ptr)
        if (self._arrivals is not None and other is self and _same_index(ind, oind)
                and _progression(ind, self._len) == (0, 1, self._len) and not self.space.kernel_flags):
            return self._gramian_while_uploading(weight_ptr)
        pa, csa, ka, keep_a = self._panel(ind)
        symmetric = other is self and _same_index(ind, oind)
        if symmetric:
            pb, csb, kb, keep_b = pa, csa, ka, keep_a
        else:
            pb, csb, kb, keep_b = other._panel(oind)
        if ka == 0 or kb == 0:
            return np.zeros((ka, kb), order='F')
        sp = self.space
        out, flags = sp.reduction_target(ka * kb)
        if min(ka, kb) <= _SKINNY and not sp.kernel_flags:
            # skinny products (projection of a right-hand side, Gramian of a few vectors) are
            # HBM-bound: stream the wide panel once per narrow column through the reduction kernel,
            # the narrow column broadcast (column stride 0) and served from L2
            last = min(ka, kb) - 1
            for c in range(last + 1):
                f = flags if c == last else 0
                if kb <= ka:      # column c of the result: <a_i, b_c> for all i
                    sp.ctx.call('pmc_pairwise_inner', pa, csa, pb + 8 * c * csb, 0, ka, self.n, weight_ptr,
                                out.data_ptr() + 8 * c * ka, f)
                else:             # row c of the result, written as a contiguous scratch row
                    sp.ctx.call('pmc_pairwise_inner', pa + 8 * c * csa, 0, pb, csb, kb, self.n, weight_ptr,
                                out.data_ptr() + 8 * c * kb, f)
            res = sp.reduction_result(out, ka * kb)
            res = res.reshape((ka, kb), order='F') if kb <= ka else np.asfortranarray(res.reshape((ka, kb)))
            if symmetric:
                low = np.tril_indices(ka, -1)
                res[low[1], low[0]] = res[low]          # bitwise symmetric like the SYRK path
            del keep_a, keep_b
            return res
        if symmetric:
            flags |= PMC_SYMMETRIC
        sp.ctx.call('pmc_gram', pa, csa, ka, pb, csb, kb, self.n, weight_ptr, out.data_ptr(), ka,
                    flags | sp.kernel_flags)
        del keep_a, keep_b
        return sp.reduction_result(out, ka * kb, symmetric_k=ka if symmetric else 0).reshape((ka, kb), order='F')

    def gramian(self, ind):
        return self.inner(self, ind, ind)

    def _gramian_while_uploading(self, weight_ptr):
        """Gramian of an array whose host -> device upload is still in flight (from_local_numpy of
        pinned data): column block b is contracted with blocks 0..b as soon as it has landed, so the
        tensor-core work hides behind the PCIe transfer of the following blocks.  Only the upper block
        triangle is computed (same flop as SYRK) and mirrored on the host."""
        sp, ctx = self.space, self.space.ctx
        arrivals, self._arrivals = self._arrivals, None
        if self in ctx.uploads:
            ctx.uploads.remove(self)
        ctx.flush()
        k = self._len
        out, flags = sp.reduction_target(k * k)
        cur = torch.cuda.current_stream(sp.device)
        for c0, c1, ev in arrivals:
            cur.wait_event(ev)
            base = out.data_ptr() + 8 * c0 * k
            ctx.call('pmc_gram', self._ptr(c0), self.ld, c1 - c0, self._ptr(c0), self.ld, c1 - c0, self.n,
                     weight_ptr, base + 8 * c0, k, PMC_SYMMETRIC)
            if c0:
                ctx.call('pmc_gram', self._ptr(0), self.ld, c0, self._ptr(c0), self.ld, c1 - c0, self.n,
                         weight_ptr, base, k, 0)
        if flags & PMC_SYNC:
            ctx.sync()
        # the strictly lower blocks were never written: only the upper triangle of the scratch is used
        raw = sp.reduction_result(out, k * k).reshape((k, k), order='F')
        self._arrival_src = None          # the host has synchronised: every block has left the staging buffer
        res = np.triu(raw)
        return res + np.triu(raw, 1).T

    def pairwise_inner(self, other, ind, oind, weight_ptr=0):
        if type(other) is not CudaVectorArrayImpl:
            from pymor_b200.complexarray import CudaComplexVectorArrayImpl
            return CudaComplexVectorArrayImpl.wrap(self).pairwise_inner(other, ind, oind, weight_ptr=weight_ptr)
        if (type(ind) is slice and type(oind) is slice and ind.step in (None, 1) and oind.step in (None, 1)
                and ind.stop is not None and oind.stop is not None
                and ind.stop - ind.start == 1 and oind.stop - oind.start == 1
                and 0 <= ind.start < self._len and 0 <= oind.start < other._len):
            # fast path: a single pair of vectors (Gram-Schmidt projection coefficient / norm)
            sp = self.space
            ctx = sp.ctx
            if ctx.uploads:
                ctx.wait_uploads()
            pend = ctx.pending
            if pend is not None:
                timpl, tcol, alpha, ximpl, xcol = pend
                same = other is self and ind.start == oind.start
                if other is timpl and oind.start == tcol and (same or not (self is timpl and ind.start == tcol)):
                    # the pending `v += alpha x` and this `<q, v>` (or ||v||^2) run as ONE pass over v
                    ctx.pending = None
                    out, flags = sp.reduction_target(1)
                    ctx.scalar[0] = alpha
                    ctx.call('pmc_axpy_dot', timpl._ptr(tcol), ximpl._ptr(xcol), ctx.scalar_ptr,
                             0 if same else self._ptr(ind.start), self.n, weight_ptr, out.data_ptr(), flags)
                    return sp.reduction_result(out, 1)
                ctx.flush()
            out, flags = sp.reduction_target(1)
            if other is self and ind.start == oind.start:
                sp.ctx.call('pmc_norm2', self._ptr(ind.start), self.ld, 1, self.n, weight_ptr, out.data_ptr(), flags)
            else:
                sp.ctx.call('pmc_pairwise_inner', self._ptr(ind.start), self.ld, other._ptr(oind.start), other.ld,
                            1, self.n, weight_ptr, out.data_ptr(), flags)
            return sp.reduction_result(out, 1)
        pa, csa, ka, keep_a = self._panel(ind)
        if ka == 0:
            return np.zeros(0)
        sp = self.space
        out, flags = sp.reduction_target(ka)
        if other is self and _same_index(ind, oind):
            sp.ctx.call('pmc_norm2', pa, csa, ka, self.n, weight_ptr, out.data_ptr(), flags)
        else:
            pb, csb, kb, keep_b = other._panel(oind)
            assert ka == kb
            sp.ctx.call('pmc_pairwise_inner', pa, csa, pb, csb, ka, self.n, weight_ptr, out.data_ptr(), flags)
            del keep_b
        del keep_a
        return sp.reduction_result(out, ka)

    def mgs_sweep(self, nq, col, weight_ptr=0, start=0, weighted=None):
        """Orthogonalise column ``col`` against columns ``start .. start+nq-1`` (orthonormal) in ONE launch:
        the inner loop of gram_schmidt (gram_schmidt.py:84-93) as a persistent kernel, cross-GPU sums
        included (pmc_mgs_sweep).  Returns the ``nq`` projection coefficients and the squared norm of
        the updated vector -- on every rank the same bits.  ``weighted``: optional impl holding the
        pre-weighted columns ``w * q_j`` at the same positions (takes the weight stream out of every step)."""
        sp, ctx = self.space, self.space.ctx
        assert 0 <= start and start + nq <= col < self._len
        ctx.flush()
        if sp.comm.size > 1:
            ctx.ensure_peers(sp.comm)
        out = ctx.host_f64(nq + 2)
        zp, zcs = (weighted._ptr(start), weighted.ld) if weighted is not None and weight_ptr else (0, 0)
        ctx.call('pmc_mgs_sweep', self._ptr(start), self.ld, zp, zcs, nq, self._ptr(col), self.n, weight_ptr,
                 out.data_ptr(), PMC_SYNC | (PMC_LOCAL_ONLY if sp.comm.size == 1 else 0))
        res = out[:nq + 2].numpy().copy()
        if res[nq + 1] != 0.0:
            raise _lib.PmcError('pmc_mgs_sweep: a wait inside the kernel timed out (ranks out of step?)')
        return res[:nq], float(res[nq])

    def norm2(self, ind, weight_ptr=0):
        pa, csa, ka, keep_a = self._panel(ind)
        if ka == 0:
            return np.zeros(0)
        sp = self.space
        out, flags = sp.reduction_target(ka)
        sp.ctx.call('pmc_norm2', pa, csa, ka, self.n, weight_ptr, out.data_ptr(), flags)
        del keep_a
        return sp.reduction_result(out, ka)

    def norm(self, ind):
        return np.sqrt(self.norm2(ind))

    def lincomb(self, coefficients, ind):
        pa, csa, k, keep = self._panel(ind)
        coefficients = np.asarray(coefficients)
        if np.iscomplexobj(coefficients):
            from pymor_b200.complexarray import CudaComplexVectorArrayImpl
            del keep
            return CudaComplexVectorArrayImpl.wrap(self).lincomb(coefficients, ind)
        assert coefficients.ndim == 2 and coefficients.shape[0] == k
        m = coefficients.shape[1]
        new = CudaVectorArrayImpl.allocate(self.space, m)
        if m == 0 or self.n == 0:
            return new
        ldc = max(_round_up(k, 2), 2)
        ct = self.space.ctx.upload_coefficients(coefficients, k, m, ldc)   # (m, ldc) C-order = F-ordered (ldc, m)
        self.space.ctx.call('pmc_lincomb', pa, csa, k, self.n, ct.data_ptr(), ldc, m, new._ptr(0), new.ld,
                            self.space.kernel_flags)
        # `ct` is released in stream order by torch's caching allocator (same stream as the kernel)
        del keep
        return new

    def dofs(self, dof_indices, ind):
        pa, csa, k, keep = self._panel(ind)
        sp = self.space
        idx = np.ascontiguousarray(dof_indices, dtype=np.int64)
        nd = idx.size
        if nd == 0 or k == 0:
            return np.zeros((nd, k), order='F')
        if sp.comm.size == 1:
            out = sp.ctx.host_f64(nd * k)
            sp.ctx.call('pmc_dofs', pa, csa, k, self.n, idx.ctypes.data, nd, out.data_ptr(), PMC_SYNC)
            return np.array(out[:nd * k].numpy()).reshape((nd, k), order='F')
        # sharded: the owner contributes the value, everybody else zero, then sum (mpi.py:421-432)
        mine = np.nonzero((idx >= sp.offset) & (idx < sp.offset + self.n))[0]
        local = torch.zeros((k, nd), dtype=torch.float64, device=sp.device)
        if mine.size:
            loc_idx = np.ascontiguousarray(idx[mine] - sp.offset)
            tmp = torch.empty((k, mine.size), dtype=torch.float64, device=sp.device)
            sp.ctx.call('pmc_dofs', pa, csa, k, self.n, loc_idx.ctypes.data, mine.size, tmp.data_ptr(), 0)
            local[:, torch.from_numpy(mine).to(sp.device)] = tmp
        sp.comm.allreduce_sum_(local)
        del keep
        return np.asfortranarray(local.to('cpu').numpy().T)

    def amax(self, ind):
        pa, csa, k, keep = self._panel(ind)
        sp = self.space
        if k == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        if sp.comm.size == 1:
            oi, ov = sp.ctx.host_i64(k), sp.ctx.host_f64(k)
            sp.ctx.call('pmc_amax', pa, csa, k, self.n, oi.data_ptr(), ov.data_ptr(), PMC_SYNC)
            return np.array(oi[:k].numpy()), np.array(ov[:k].numpy())
        oi = torch.zeros(k, dtype=torch.int64, device=sp.device)
        ov = torch.full((k,), -1.0, dtype=torch.float64, device=sp.device)
        if self.n:
            sp.ctx.call('pmc_amax', pa, csa, k, self.n, oi.data_ptr(), ov.data_ptr(), 0)
            oi += sp.offset
        inds = torch.stack(sp.comm.allgather(oi)).to('cpu').numpy()
        vals = torch.stack(sp.comm.allgather(ov)).to('cpu').numpy()
        del keep
        return merge_amax(inds, vals)


def _same_index(a, b):
    if a is b:
        return True
    if type(a) is not type(b):
        return False
    return a == b


from pymor_b200.complexarray import CudaComplexVectorArrayImpl  # noqa: E402


class CudaVectorArray(VectorArray):
    """|VectorArray| whose vectors live in B200 HBM."""

    impl_type = (CudaVectorArrayImpl, CudaComplexVectorArrayImpl)

    def to_torch(self):
        """Zero-copy ``(len, local_dim)`` torch view of the (non-view) array's shard."""
        impl = self.impl
        if type(impl) is not CudaVectorArrayImpl:
            raise NotImplementedError('DLPack / torch export is defined for real arrays only')
        self.space.ctx.flush()
        prog = _progression(self.ind, impl._len)
        if prog is None or prog[1] < 1:
            impl, prog = impl._copied(self.ind), (0, 1, self._len)
        start, step, count = prog
        return impl._buf[start:start + step * count:step, :impl.n]

    def __dlpack__(self, *args, **kwargs):
        return self.to_torch().__dlpack__(*args, **kwargs)

    def __dlpack_device__(self):
        return self.to_torch().__dlpack_device__()

    def __str__(self):
        return f'CudaVectorArray(len={len(self)}, dim={self.dim}, device={self.space.device})'


class CudaVectorSpace(VectorSpace):
    """|VectorSpace| of :class:`CudaVectorArray` (reference counterpart: numpy.py:242-337).

    Parameters
    ----------
    dim
        GLOBAL number of DOFs.
    device
        torch device of this rank (default: current CUDA device).
    comm
        ``None`` (single GPU) or a :class:`pymor_b200.dist.ProcessGroupComm`: rows are split evenly
        over the ranks.
    id
        See pyMOR's ``VectorSpace.id``.
    """

    def __init__(self, dim, device=None, comm=None, id=None):
        self.dim = int(dim)
        self.id = id
        self._comm = comm if comm is not None else SingleRank()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else 'cuda'
        self._ctx = _lib.get_context(device)           # raises without library / device
        self.device = self._ctx.device
        self._comm.attach(self._ctx)                   # collective: NCCL communicator behind the C ABI (dist.py)
        dims, offsets = partition(self.dim, self._comm.size)
        self._local_dims = dims
        self._local_dim = int(dims[self._comm.rank])
        self._offset = int(offsets[self._comm.rank])
        self._ld = max(_round_up(self._local_dim, _ALIGN), _ALIGN)
        self._red_dev = None
        self._host_reducer = None
        self._host_reduce_enabled = os.environ.get('PMC_NO_HOST_REDUCE', '0') != '1'
        self._kernel_flags = 0
        self._host_np = None

    # read-only views of the private state (the pyMOR base class freezes public attributes)
    ctx = property(lambda self: self._ctx)
    local_dim = property(lambda self: self._local_dim)
    local_dims = property(lambda self: self._local_dims)
    offset = property(lambda self: self._offset)
    ld = property(lambda self: self._ld)
    kernel_flags = property(lambda self: self._kernel_flags)

    comm = property(lambda self: self._comm)

    def force_generic_kernels(self, enable=True):
        """Route gramian/lincomb through the generic kernels (test hook for kernel cross-checks)."""
        self._kernel_flags = PMC_FORCE_SIMPLE if enable else 0

    # -- reductions: single GPU -> kernel writes pinned host memory; sharded -> device + all-reduce
    def reduction_target(self, count):
        if self._comm.size == 1:
            return self._ctx.host_f64(count), PMC_SYNC
        if count <= _HOST_REDUCE_MAX and self._host_reduce_enabled:
            # a handful of doubles the host needs anyway: per-rank partials meet in pinned shared memory
            if self._host_reducer is None:
                from pymor_b200.dist import get_host_reducer
                self._host_reducer = get_host_reducer(self._comm, self._ctx, _HOST_REDUCE_MAX)
            return _RawTarget(self._host_reducer.target()), PMC_SYNC
        if self._red_dev is None or self._red_dev.numel() < count:
            self._red_dev = torch.empty(max(int(count), 4096), dtype=torch.float64, device=self.device)
        return self._red_dev, 0

    def reduction_result(self, buf, count, symmetric_k=0):
        """Host copy of ``count`` reduced doubles (summed over the ranks).  ``symmetric_k``: the result is a symmetric
        ``k x k`` matrix -- after a cross-GPU all-reduce its lower triangle is mirrored (``pmc_mirror_lower``), because
        the collective adds different chunks in different rank orders."""
        if type(buf) is _RawTarget:
            return self._host_reducer.finish(count)
        if self._comm.size == 1:
            host = self._host_np
            if host is None or host[0] is not buf:
                host = self._host_np = (buf, buf.numpy())
            return host[1][:count].copy()
        part = buf[:count]
        self._comm.allreduce_sum_(part)
        if symmetric_k > 1 and self._comm.size > 2:
            self._ctx.call('pmc_mirror_lower', part.data_ptr(), symmetric_k, symmetric_k, 0)
        if part.is_cuda:                                # D2H through the context's pinned buffer: one DMA, no staging
            host = self._ctx.host_f64(count)[:count]
            host.copy_(part, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            return host.numpy().copy()
        return part.to('cpu').numpy().copy()

    # -- factory ------------------------------------------------------------------------------
    def __eq__(self, other):
        if other is self:
            return True
        return (type(other) is type(self) and self.dim == other.dim and self.id == other.id
                and self.device == other.device and self._comm == other._comm)

    def __hash__(self):
        return hash((self.dim, self.id, str(self.device)))

    def zeros(self, count=1, reserve=0):
        assert count >= 0 and reserve >= 0
        return CudaVectorArray(self, CudaVectorArrayImpl.allocate(self, count, reserve, zero=True))

    def full(self, value, count=1, reserve=0):
        assert count >= 0 and reserve >= 0
        if isinstance(value, complex) or np.iscomplexobj(value):
            re, im = self.full(complex(value).real, count, reserve), self.full(complex(value).imag, count, reserve)
            return CudaVectorArray(self, CudaComplexVectorArrayImpl(self, re.impl, im.impl))
        impl = CudaVectorArrayImpl.allocate(self, count, reserve)
        impl._buf[:count].fill_(float(value))
        return CudaVectorArray(self, impl)

    def random(self, count=1, distribution='uniform', reserve=0, **kwargs):
        """Host-generated through pyMOR's global RNG, so that inputs are bit-identical with
        ``NumpyVectorSpace.random`` (interface.py:986-1007, numpy.py:272-277)."""
        assert count >= 0 and reserve >= 0
        values = create_random_values((self.dim, count), distribution, **kwargs)
        return self._upload(values, reserve)

    def random_device(self, count=1, seed=0, distribution='normal', scale=1.0, reserve=0):
        """Device-side generation (torch Philox) for full-size benchmark inputs; per-rank seed."""
        impl = CudaVectorArrayImpl.allocate(self, count, reserve)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(seed) * 1000003 + self._comm.rank)
        view = impl._buf[:count, :self._local_dim]
        if distribution == 'normal':
            view.normal_(0.0, scale, generator=gen)
        else:
            view.uniform_(0.0, scale, generator=gen)
        return CudaVectorArray(self, impl)

    def _upload(self, data, reserve=0):
        data = np.asarray(data)
        if np.iscomplexobj(data):
            re, im = self._upload(data.real, reserve), self._upload(data.imag, reserve)
            return CudaVectorArray(self, CudaComplexVectorArrayImpl(self, re.impl, im.impl))
        if data.ndim == 1:
            data = data.reshape((-1, 1))
        assert data.ndim == 2 and data.shape[0] == self.dim
        count = data.shape[1]
        impl = CudaVectorArrayImpl.allocate(self, count, reserve)
        if count and self._local_dim:
            rows = np.ascontiguousarray(data[self._offset:self._offset + self._local_dim].T, dtype=np.float64)
            impl._buf[:count, :self._local_dim].copy_(torch.from_numpy(rows))
        return CudaVectorArray(self, impl)

    def wait_uploads(self):
        """Block the host until every host -> device upload started by :meth:`from_local_numpy` has left its
        source buffer.  Call this before refilling a (pinned) staging buffer that was handed to
        ``from_local_numpy`` -- the copies are asynchronous."""
        ctx = self._ctx
        if ctx.uploads:
            ctx.wait_uploads()
        if ctx._copy_stream is not None:
            ctx._copy_stream.synchronize()
        if self.device.type == 'cuda':
            torch.cuda.current_stream(self.device).synchronize()

    def from_local_numpy(self, shard, reserve=0):
        """Upload THIS RANK's rows: ``shard`` has shape ``(local_dim, len)``.  A Fortran-ordered
        shard (one vector per contiguous column) in pinned memory is copied by a single DMA.

        The copy is ASYNCHRONOUS (device work is ordered after it; large pinned shards travel on a side stream
        in column blocks and the first Gramian starts on the blocks as they land): the caller must not modify
        ``shard`` until :meth:`wait_uploads` has returned or a result computed from the array has reached the
        host."""
        shard = np.asarray(shard)
        if np.iscomplexobj(shard):
            re, im = self.from_local_numpy(shard.real, reserve), self.from_local_numpy(shard.imag, reserve)
            return CudaVectorArray(self, CudaComplexVectorArrayImpl(self, re.impl, im.impl))
        assert shard.ndim == 2 and shard.shape[0] == self._local_dim
        count = shard.shape[1]
        impl = CudaVectorArrayImpl.allocate(self, count, reserve)
        if count and self._local_dim:
            rows = shard.T if shard.T.flags.c_contiguous and shard.dtype == np.float64 \
                else np.ascontiguousarray(shard.T, dtype=np.float64)
            src = torch.from_numpy(rows)
            n = self._local_dim
            if (self.device.type == 'cuda' and count >= 64 and src.numel() * 8 >= _STREAMED_UPLOAD_BYTES
                    and src.is_pinned()):
                # large pinned shard: copy column blocks on a side stream; consumers wait per block
                # (the Gramian works on the blocks as they land, see _gramian_while_uploading)
                blk = _round_up(-(-count // _UPLOAD_BLOCKS), 8)
                cs = self._ctx.copy_stream()
                cs.wait_stream(torch.cuda.current_stream(self.device))
                arrivals = []
                with torch.cuda.stream(cs):
                    for c0 in range(0, count, blk):
                        c1 = min(count, c0 + blk)
                        impl._buf[c0:c1, :n].copy_(src[c0:c1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(cs)
                        arrivals.append((c0, c1, ev))
                impl._buf.record_stream(cs)
                impl._arrivals, impl._arrival_src = arrivals, src
                self._ctx.uploads.append(impl)
            else:
                impl._buf[:count, :n].copy_(src, non_blocking=True)
        return CudaVectorArray(self, impl)

    def from_numpy(self, data, ensure_copy=False):
        """``data`` is the GLOBAL ``(dim, len)`` array; always copies (host -> device)."""
        return self._upload(data)

    def make_array(self, obj):
        """Adopt raw backend data: a ``(len, local_dim)`` fp64 CUDA tensor or any ``__dlpack__``
        producer of one (zero-copy when 16-byte aligned with an even row stride), or a NumPy array
        of shape ``(dim, len)``."""
        if isinstance(obj, np.ndarray):
            return self._upload(obj)
        t = obj if isinstance(obj, torch.Tensor) else torch.from_dlpack(obj)
        assert t.dtype == torch.float64 and t.dim() == 2 and t.shape[1] == self._local_dim
        assert t.device == self.device
        ok = (t.shape[0] > 0 and t.shape[1] > 0 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]
              and t.stride(0) % 2 == 0 and t.data_ptr() % 16 == 0)
        if ok:
            return CudaVectorArray(self, CudaVectorArrayImpl(self, t, t.shape[0]))
        impl = CudaVectorArrayImpl.allocate(self, t.shape[0])
        if t.shape[0]:
            impl._buf[:t.shape[0], :self._local_dim].copy_(t)
        return CudaVectorArray(self, impl)

    def __repr__(self):
        return f'CudaVectorSpace(dim={self.dim}, device={self.device!s}, ranks={self._comm.size})'
