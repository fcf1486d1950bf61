"""NumPy emulation of the C ABI of include/pymor_b200.h -- TEST INFRASTRUCTURE ONLY.

The build container has no GPU.  To exercise the host-side Python layer (view / copy-on-write /
aliasing logic, pyMOR integration, rank-sharded reductions) on CPU, tests inject this object in
place of the ctypes library (``pymor_b200._lib.set_library_for_testing``).  Every entry point has
the exact signature of the C function and operates on raw host addresses, using the CPU oracle's
arithmetic.  The product never imports this module; without an injected library (and without a
B200) creating a ``CudaVectorSpace`` raises.
"""
import ctypes

import numpy as np

PMC_OK, PMC_ERR_INVALID = 0, -1
PMC_SYMMETRIC = 2


def _vec(ptr, count, dtype=np.float64):
    if count == 0:
        return np.zeros(0, dtype=dtype)
    ctype = {np.float64: ctypes.c_double, np.int64: ctypes.c_int64, np.int32: ctypes.c_int32}[dtype]
    return np.ctypeslib.as_array((ctype * int(count)).from_address(int(ptr)))


def _col(ptr, cs, c, n):
    return _vec(int(ptr) + 8 * c * cs, n)


def _panel(ptr, cs, k, n):
    """Read a strided panel into a fresh (n, k) array."""
    out = np.empty((n, k), order='F')
    for c in range(k):
        out[:, c] = _col(ptr, cs, c, n)
    return out


class EmulatedLib:
    def __init__(self):
        self.launches = 0
        self.calls = []
        self._err = b''

    # -- context -------------------------------------------------------------------------------
    def pmc_version(self):
        return 100

    def pmc_last_error(self):
        return self._err

    def pmc_ctx_create(self, device, stream, out):
        out._obj.value = 1
        return PMC_OK

    def pmc_ctx_destroy(self, ctx):
        return PMC_OK

    def pmc_ctx_set_workspace(self, ctx, ptr, nbytes):
        return PMC_OK

    def pmc_ctx_workspace_needed(self, ctx):
        return 0

    def pmc_ctx_sync(self, ctx):
        return PMC_OK

    def pmc_ctx_sm_count(self, ctx):
        return 148

    def pmc_ctx_launch_count(self, ctx):
        return self.launches

    def pmc_ctx_set_option(self, ctx, name, value):
        if name not in (b'gram_opts', b'gram_waves', b'gram_l2_mb', b'csr_variant', b'l2_keep', b'kernel_timing'):
            self._err = b'unknown option'
            return PMC_ERR_INVALID
        self.options = getattr(self, 'options', {})
        self.options[name] = int(value)
        return PMC_OK

    def pmc_ctx_get_option(self, ctx, name, out):
        out._obj.value = getattr(self, 'options', {}).get(name, 0)
        return PMC_OK

    def pmc_ctx_kernel_times(self, ctx, kinds, ms, cap, count):
        count._obj.value = 0                       # the emulation launches no kernels
        return PMC_OK

    def pmc_host_register(self, ctx, host_ptr, nbytes, out):
        out._obj.value = int(host_ptr)
        return PMC_OK

    def pmc_host_unregister(self, ctx, host_ptr):
        return PMC_OK

    def _note(self, name):
        self.launches += 1
        self.calls.append(name)

    # -- kernels -------------------------------------------------------------------------------
    def pmc_gram(self, ctx, A, csA, kA, B, csB, kB, n, w, out, ldo, flags):
        self._note('pmc_gram')
        if flags & PMC_SYMMETRIC and not (A == B and csA == csB and kA == kB):
            self._err = b'PMC_SYMMETRIC needs identical panels'
            return PMC_ERR_INVALID
        a, b = _panel(A, csA, kA, n), _panel(B, csB, kB, n)
        if w:
            b = b * _vec(w, n)[:, None]
        g = a.T @ b
        if flags & PMC_SYMMETRIC:
            g = np.tril(g) + np.tril(g, -1).T
        o = _vec(out, ldo * kB).reshape((ldo, kB), order='F')
        o[:kA, :] = g
        return PMC_OK

    def pmc_lincomb(self, ctx, A, csA, k, n, CT, ldc, m, R, csR, flags):
        self._note('pmc_lincomb')
        a = _panel(A, csA, k, n)
        c = _vec(CT, ldc * m).reshape((ldc, m), order='F')[:k]
        r = a @ c
        for j in range(m):
            _col(R, csR, j, n)[:] = r[:, j]
        return PMC_OK

    def pmc_pairwise_inner(self, ctx, A, csA, B, csB, k, n, w, out, flags):
        self._note('pmc_pairwise_inner')
        a, b = _panel(A, csA, k, n), _panel(B, csB, k, n)
        if w:
            b = b * _vec(w, n)[:, None]
        _vec(out, k)[:] = np.sum(a * b, axis=0)
        return PMC_OK

    def pmc_norm2(self, ctx, A, csA, k, n, w, out, flags):
        self._note('pmc_norm2')
        a = _panel(A, csA, k, n)
        b = a * _vec(w, n)[:, None] if w else a
        _vec(out, k)[:] = np.sum(a * b, axis=0)
        return PMC_OK

    def pmc_axpy(self, ctx, Y, csY, X, csX, k, n, alpha, n_alpha, flags):
        self._note('pmc_axpy')
        al = _vec(alpha, n_alpha)
        x = _panel(X, csX, k, n)            # evaluated first; the host layer guarantees no aliasing
        for c in range(k):
            _col(Y, csY, c, n)[:] += al[c if n_alpha > 1 else 0] * x[:, c]
        return PMC_OK

    def pmc_axpy_dot(self, ctx, Y, X, alpha, Q, n, w, out, flags):
        self._note('pmc_axpy_dot')
        y = _vec(Y, n)
        y += _vec(alpha, 1)[0] * _vec(X, n).copy()
        b = y * _vec(w, n) if w else y
        _vec(out, 1)[0] = np.sum((_vec(Q, n) if Q else y) * b)
        return PMC_OK

    def pmc_mgs_sweep(self, ctx, Q, csQ, Z, csZ, nq, v, n, w, out, flags):
        """One launch on the device; across ranks the kernel sums over NVLink peer memory -- emulated by the
        ``sweep_allreduce`` hook the multi-process tests install (sum of a double over the ranks)."""
        self._note('pmc_mgs_sweep')
        red = getattr(self, 'sweep_allreduce', None) or (lambda x: x)
        y = _vec(v, n)
        ww = _vec(w, n) if w else None
        o = _vec(out, nq + 2)
        for j in range(nq):
            q = _col(Q, csQ, j, n)
            if Z:
                assert w, 'pre-weighted panel without weight vector'
                o[j] = red(float(np.sum(_col(Z, csZ, j, n) * y)))
            else:
                o[j] = red(float(np.sum(q * (y * ww if w else y))))
            y -= o[j] * q
        o[nq] = red(float(np.sum(y * (y * ww if w else y))))
        o[nq + 1] = 0.0
        return PMC_OK

    def pmc_peer_create(self, ctx, rank, size, handle_out):
        return PMC_OK

    def pmc_peer_connect(self, ctx, handles):
        return PMC_OK

    def pmc_peer_destroy(self, ctx):
        return PMC_OK

    def pmc_scal(self, ctx, Y, csY, k, n, alpha, n_alpha, flags):
        self._note('pmc_scal')
        al = _vec(alpha, n_alpha)
        for c in range(k):
            _col(Y, csY, c, n)[:] *= al[c if n_alpha > 1 else 0]
        return PMC_OK

    def pmc_copy(self, ctx, D, csD, S, csS, k, n, flags):
        self._note('pmc_copy')
        # column-parallel on the device: emulate by reading everything before writing, and verify the
        # host layer never asks for overlapping source/destination columns
        src = _panel(S, csS, k, n)
        if n:
            s_addr = {int(S) + 8 * c * csS for c in range(k)}
            d_addr = {int(D) + 8 * c * csD for c in range(k)}
            assert s_addr.isdisjoint(d_addr), 'pmc_copy called with overlapping columns'
        for c in range(k):
            _col(D, csD, c, n)[:] = src[:, c]
        return PMC_OK

    def pmc_gather_cols(self, ctx, D, csD, base, ld, idx, k, n, flags):
        self._note('pmc_gather_cols')
        ix = _vec(idx, k, np.int64)
        for c in range(k):
            _col(D, csD, c, n)[:] = _col(base, ld, int(ix[c]), n)
        return PMC_OK

    def pmc_dofs(self, ctx, A, csA, k, n, dof, ndofs, out, flags):
        self._note('pmc_dofs')
        ix = _vec(dof, ndofs, np.int64)
        if ndofs and (ix.min() < 0 or ix.max() >= n):
            self._err = b'dof index out of range'
            return PMC_ERR_INVALID
        o = _vec(out, ndofs * k).reshape((ndofs, k), order='F')
        for c in range(k):
            o[:, c] = _col(A, csA, c, n)[ix]
        return PMC_OK

    def pmc_amax(self, ctx, A, csA, k, n, out_idx, out_val, flags):
        self._note('pmc_amax')
        a = np.abs(_panel(A, csA, k, n))
        idx = np.argmax(a, axis=0)
        _vec(out_idx, k, np.int64)[:] = idx
        _vec(out_val, k)[:] = a[idx, np.arange(k)]
        return PMC_OK

    def pmc_diag_apply(self, ctx, w, X, csX, Y, csY, k, n, flags):
        self._note('pmc_diag_apply')
        x = _panel(X, csX, k, n) * _vec(w, n)[:, None]
        for c in range(k):
            _col(Y, csY, c, n)[:] = x[:, c]
        return PMC_OK

    def _csr_times(self, nrows, rowptr, colidx, vals, X, csX, k, ghost):
        """M [X; ghost] with column indices >= nrows addressing the row-major ghost buffer."""
        import scipy.sparse as sps
        rp = _vec(rowptr, nrows + 1, np.int64)
        nnz = int(rp[-1])
        ci = _vec(colidx, nnz, np.int32).copy()
        nghost = int(ci.max()) + 1 - nrows if nnz and ci.max() >= nrows else 0
        m = sps.csr_matrix((_vec(vals, nnz).copy(), ci, rp.copy()), shape=(nrows, nrows + nghost))
        x = _panel(X, csX, k, nrows)
        if nghost:
            assert ghost, 'halo columns present but no ghost buffer passed'
            x = np.vstack([x, _vec(ghost, nghost * k).reshape((nghost, k))])
        return m @ x

    def pmc_csr_apply(self, ctx, nrows, rowptr, colidx, vals, X, csX, ghost, Y, csY, k, flags):
        self._note('pmc_csr_apply')
        y = self._csr_times(nrows, rowptr, colidx, vals, X, csX, k, ghost)
        for c in range(k):
            _col(Y, csY, c, nrows)[:] = y[:, c]
        return PMC_OK

    def pmc_mirror_lower(self, ctx, G, k, ld, flags):
        self._note('pmc_mirror_lower')
        g = _vec(G, ld * k).reshape((ld, k), order='F')
        low = np.tril_indices(k, -1)
        g[low[1], low[0]] = g[low]
        return PMC_OK

    def pmc_shift_cols(self, ctx, A, cs, first, count, shift, n, flags):
        self._note('pmc_shift_cols')
        for c in range(first, first + count):                 # increasing order: overlapping ranges are fine
            _col(A, cs, c - shift, n)[:] = _col(A, cs, c, n)
        return PMC_OK

    def pmc_gather_rows(self, ctx, A, csA, k, n, idx, nidx, out, flags):
        self._note('pmc_gather_rows')
        ix = _vec(idx, nidx, np.int64)
        _vec(out, nidx * k).reshape((nidx, k))[:] = _panel(A, csA, k, n)[ix, :]
        return PMC_OK

    def pmc_csr_gram(self, ctx, nrows, rowptr, colidx, vals, V, csV, kV, U, csU, kU, ghost, out, ldo, flags):
        self._note('pmc_csr_gram')
        g = _panel(V, csV, kV, nrows).T @ self._csr_times(nrows, rowptr, colidx, vals, U, csU, kU, ghost)
        if flags & 8:                              # PMC_LOWER_ONLY: lower triangle computed, the rest mirrored
            g = np.tril(g) + np.tril(g, -1).T
        o = _vec(out, ldo * kU).reshape((ldo, kU), order='F')
        o[:kV, :] = g
        return PMC_OK

    def pmc_csr_pairwise(self, ctx, nrows, rowptr, colidx, vals, V, csV, U, csU, ghost, k, out, flags):
        self._note('pmc_csr_pairwise')
        MU = self._csr_times(nrows, rowptr, colidx, vals, U, csU, k, ghost)
        _vec(out, k)[:] = np.einsum('ij,ij->j', _panel(V, csV, k, nrows), MU) if nrows else 0.0
        return PMC_OK

    def pmc_measure_fp64_peak(self, ctx, use_dmma, ms, out):
        _vec(out, 1)[0] = 0.0
        return PMC_OK
