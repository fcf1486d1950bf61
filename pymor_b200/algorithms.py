"""Stand-alone drivers of the hot-path algorithms.

These are the CALLERS of the backend.  With pyMOR installed use pyMOR's own functions -- they run
unchanged on :class:`~pymor_b200.vectorarray.CudaVectorArray`.  On a box without pyMOR (the GPU box of
this project) the functions below provide the same algorithms with the same names, arguments,
defaults and error behaviour, written against the VectorArray/Operator interface only (they never
touch raw data, like their reference counterparts):

* :func:`gram_schmidt`          -- src/pymor/algorithms/gram_schmidt.py:13-122
* :func:`method_of_snapshots`   -- src/pymor/algorithms/svd_va.py:20-99
* :func:`qr_svd`                -- src/pymor/algorithms/svd_va.py:103-168
* :func:`pod`                   -- src/pymor/algorithms/pod.py:16-90
* :func:`project`               -- src/pymor/algorithms/projection.py:28-85 (rule ``action_apply_basis``)
* :func:`shifted_chol_qr`       -- src/pymor/algorithms/chol_qr.py:16-252 ("next" row N1: the GEMM-bound
  orthonormalisation -- three Gramians + three block lincombs instead of k^2 dot/axpy pairs with a
  host sync each, i.e. the variant that scales over DOF-sharded GPUs)
* :func:`inc_vectorarray_hapod`, :func:`dist_vectorarray_hapod`, :func:`inc_hapod_chunks` (snapshots streamed
  from host memory chunk by chunk) -- src/pymor/algorithms/hapod.py:333, :372 ("next" row N3)
* :func:`randomized_range_finder` -- src/pymor/algorithms/rand_la.py:20 ("next" row N2, fixed-size mode)
* :func:`block_gram_schmidt` (= ``gram_schmidt(variant='bcgs2')``) and ``gram_schmidt(variant='cgs2')`` -- not in
  the reference: the classical / block-classical Gram-Schmidt variants with re-orthogonalisation that turn the
  O(k^2) host-synchronised pair-steps of gram_schmidt.py:84-91 into O(k) streaming resp. tensor-core calls; same
  contract and (the QR factorisation being unique) the same results to rounding x conditioning
* :func:`ei_greedy`, :func:`deim` -- src/pymor/algorithms/ei.py:30, :182 ("next" row N5: empirical
  interpolation data; the backend sees ``amax`` / ``dofs`` / per-column ``axpy`` / ``norm`` / ``lincomb``)
"""
from __future__ import annotations

import logging
import os

import numpy as np
import scipy.linalg as spla

logger = logging.getLogger('pymor_b200.algorithms')


class AccuracyError(Exception):
    """Raised when the result of :func:`gram_schmidt` fails the orthonormality check."""


def _sweep_weight(A, product):
    """Device address of the diagonal weight (0: Euclidean) if the inner loop of :func:`gram_schmidt` can
    run as one fused device sweep per pass (``CudaVectorArrayImpl.mgs_sweep``), else ``None``: a real,
    non-view CudaVectorArray and no product or a diagonal one."""
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorArray, CudaVectorArrayImpl
    if not isinstance(A, CudaVectorArray) or type(A.impl) is not CudaVectorArrayImpl or A.ind is not None:
        return None
    if A.space.comm.size > 8:
        return None
    if product is None:
        return 0
    if isinstance(product, CudaDiagonalOperator) and product.source == A.space:
        return product._w.data_ptr()
    return None


def gram_schmidt(A, product=None, return_R=False, atol=1e-13, rtol=1e-13, offset=0,
                 reiterate=True, reiteration_threshold=9e-1, check=True, check_tol=1e-3, copy=True, sweep=None,
                 variant='mgs'):
    """Modified Gram-Schmidt with re-iteration; vectors that are (numerically) zero or linearly
    dependent are removed.  Returns ``Q`` or ``(Q, R)``.

    ``sweep`` (not in the reference; default: environment ``PMC_MGS_SWEEP=1``, else off): run each pass of
    the inner loop -- all projections of vector i onto the previous vectors and the norm that follows -- as
    ONE persistent device kernel (``pmc_mgs_sweep``) instead of one launch and one host synchronisation per
    pair; on row-sharded arrays the per-pair sums over the GPUs happen inside that kernel over NVLink.
    Same recurrences in the same order; sums are grouped differently, so results agree to rounding.

    ``variant`` (not in the reference): ``'mgs'`` is the reference's modified Gram-Schmidt.  ``'cgs2'`` projects
    against ALL previous vectors at once per pass -- ``r = Q^T W v`` (one streaming launch, one host sync),
    ``v -= Q r`` (one block lincomb + axpy) -- with the same norm-drop re-iteration rule ("twice is enough"):
    the QR factorisation it converges to is the same (it is unique), the HBM traffic per pass is the same as
    MGS's (Q is read twice), but a pass costs two host synchronisations instead of one per previous vector,
    which is what makes the orthonormalisation scale over row-sharded GPUs.  Uses only the VectorArray
    interface (``inner``, ``lincomb``, ``axpy``), so it also runs on pyMOR's own array classes.
    ``'bcgs2'``: :func:`block_gram_schmidt` (panels of vectors against all previous ones through the tensor-core
    Gramian / lincomb kernels)."""
    assert variant in ('mgs', 'cgs2', 'bcgs2')
    if variant == 'bcgs2':
        return block_gram_schmidt(A, product=product, return_R=return_R, atol=atol, rtol=rtol, offset=offset,
                                  reiteration_threshold=reiteration_threshold, check=check, check_tol=check_tol,
                                  copy=copy)
    if copy:
        A = A.copy()
    k = len(A)
    R = np.eye(k)
    dropped = []
    if sweep is None:
        sweep = os.environ.get('PMC_MGS_SWEEP', '0') == '1'
    weight_ptr = _sweep_weight(A, product) if sweep else None
    Z = None                # weighted sweeps: panel of the pre-weighted final columns w * q_j (same positions as in A)
    if weight_ptr:
        from pymor_b200.vectorarray import CudaVectorArrayImpl
        try:
            Z = CudaVectorArrayImpl.allocate(A.space, k)
        except RuntimeError:                       # no room for a second panel: the sweeps stream w instead
            Z = None

    def weigh(first, count):
        # Z[:, first:first+count] = w * A[:, first:first+count] for columns that have become final
        if Z is not None and count and A.impl.n:
            A.space.ctx.flush()
            A.space.ctx.call('pmc_diag_apply', weight_ptr, A.impl._ptr(first), A.impl.ld, Z._ptr(first), Z.ld, count,
                             A.impl.n, 0)
    weigh(0, offset)
    for i in range(offset, k):
        v = A[i]
        norm0 = v.norm(product)[0]
        if norm0 <= atol:
            logger.info('Removing vector %d of norm %g', i, norm0)
            dropped.append(i)
            continue
        if i == 0:
            v.scal(1 / norm0)
            R[0, 0] = norm0
            weigh(0, 1)
            continue
        cur = norm0
        while True:
            if variant == 'cgs2':
                # classical pass against all kept previous vectors at once (runs of consecutive columns are views)
                kept = [j for j in range(i) if j not in dropped]
                runs = []
                for j in kept:
                    if runs and runs[-1][0] + runs[-1][1] == j:
                        runs[-1][1] += 1
                    else:
                        runs.append([j, 1])
                # r = Q^T (W v): weigh the single vector once instead of streaming the weight past every column of Q
                # (this is the reference's own evaluation order of apply2, operators/interface.py:108)
                wv = v if product is None else product.apply(v)
                coeffs = [A[start:start + count].inner(wv)[:, 0] for start, count in runs]
                for (start, count), r in zip(runs, coeffs):
                    if np.iscomplexobj(r) and not np.iscomplexobj(R):
                        R = R.astype(np.complex128)
                    v.axpy(-1., A[start:start + count].lincomb(r))
                    R[start:start + count, i] += r
                prev, cur = cur, v.norm(product)[0]
            elif weight_ptr is not None:
                # fused pass: the kept columns before i are final and orthonormal, column i is swept against
                # them -- one launch per run of consecutive kept columns (a single one unless vectors were removed)
                kept = [j for j in range(i) if j not in dropped]
                runs = []
                for j in kept:
                    if runs and runs[-1][0] + runs[-1][1] == j:
                        runs[-1][1] += 1
                    else:
                        runs.append([j, 1])
                nrm2 = None
                for start, count in runs or [[0, 0]]:
                    coeffs, nrm2 = A.impl.mgs_sweep(count, i, weight_ptr, start=start, weighted=Z)
                    R[start:start + count, i] += coeffs
                prev, cur = cur, float(np.sqrt(nrm2))
            else:
                for j in range(i):
                    if j in dropped:
                        continue
                    q = A[j]
                    p = q.pairwise_inner(v, product)[0]
                    v.axpy(-p, q)
                    if np.iscomplexobj(p) and not np.iscomplexobj(R):
                        R = R.astype(np.complex128)
                    R[j, i] += p
                prev, cur = cur, v.norm(product)[0]
            if cur <= rtol * norm0:
                logger.info('Removing linearly dependent vector %d', i)
                dropped.append(i)
                break
            if reiterate and cur < reiteration_threshold * prev:
                logger.info('Orthonormalizing vector %d again', i)
                continue
            v.scal(1 / cur)
            R[i, i] = cur
            weigh(i, 1)
            break
    if dropped:
        del A[dropped]
        R = np.delete(R, dropped, axis=0)
    if check:
        m = len(A) - offset
        err = A[offset:len(A)].inner(A, product)
        err[:m, offset:len(A)] -= np.eye(m)
        if err.size > 0:
            worst = np.max(np.abs(err))
            if worst >= check_tol:
                raise AccuracyError(f'result not orthogonal (max err={worst})')
    return (A, R) if return_R else A


def _column_runs(cols):
    """Maximal runs of consecutive integers in the sorted list ``cols`` as ``(start, count)``: each run is a
    zero-copy slice view of the array."""
    runs = []
    for c in cols:
        if runs and runs[-1][0] + runs[-1][1] == c:
            runs[-1][1] += 1
        else:
            runs.append([c, 1])
    return runs


def block_gram_schmidt(A, product=None, return_R=False, atol=1e-13, rtol=1e-13, offset=0, panel=32,
                       reiteration_threshold=9e-1, check=True, check_tol=1e-3, copy=True):
    """Block classical Gram-Schmidt with re-orthogonalisation (BCGS2) -- the Gram-Schmidt a tensor-core machine
    wants (not in the reference; same contract as :func:`gram_schmidt`: orthonormal ``Q`` spanning the same
    nested spaces, upper-triangular ``R`` with ``A = Q R``, vectors of norm <= ``atol`` or numerically dependent
    on their predecessors removed together with their rows of ``R``).

    The vectors are taken in panels of ``panel`` columns.  A panel ``X`` is projected against all previous
    vectors ``Q`` at once, ``S = Q^T W X`` (tall-skinny GEMM on the FP64 tensor pipe, ONE host sync, one k x b
    all-reduce on row-sharded arrays) and ``X -= Q S`` (block lincomb), orthonormalised in itself, and the two
    steps are repeated once ("twice is enough", Barlow & Smoktunowicz 2013: orthogonality O(eps) as long as the
    input is numerically full rank).  With ``A_P = Q (S1 + S2 R1) + Q_P (R2 R1)`` the factor ``R`` is accumulated
    on the host.  ``panel`` may be a tuple, e.g. ``(128, 16)``: a panel is then orthonormalised in itself by the
    same scheme with the next size (128-wide panels keep the 128 x 128 tensor-core tiles full, 16-wide
    sub-panels keep the innermost work small); the innermost panels use classical Gram-Schmidt with
    re-iteration, vector by vector, on the streaming kernels.  Work: ~4 n k^2 flop on the tensor cores +
    O(n k b) streamed bytes, O(k) host syncs instead of O(k^2) -- the QR factorisation is unique, so ``Q`` and
    ``R`` agree with modified Gram-Schmidt to rounding x conditioning.  A vector counts as dependent when its
    norm after the projections is <= ``rtol`` times its norm when it entered the step (the reference compares
    with the norm of the input vector)."""
    panels = tuple(int(b) for b in (panel if isinstance(panel, (tuple, list)) else (panel,)))
    assert 0 <= offset <= len(A) and all(b >= 1 for b in panels)
    if copy:
        A = A.copy()
    k = len(A)
    dropped = []

    def weigh(V):
        return V if product is None else product.apply(V)

    def view(cols):
        return A[cols[0]:cols[-1] + 1] if len(cols) == cols[-1] - cols[0] + 1 else A[cols]

    def project(cols, against):
        """cols -= Q (Q^T W cols) for Q = A[against]; returns the coefficients (len(against), len(cols))."""
        X = view(cols)
        # products that weigh inside the Gramian kernel (diagonal weights on the device) skip the temporary W X
        fused = product is not None and getattr(product, 'fused_apply2', False)
        WX = X if fused else weigh(X)
        S = np.zeros((len(against), len(cols)))
        row = 0
        for start, count in _column_runs(against):
            Q = A[start:start + count]
            C = product.apply2(Q, X) if fused else Q.inner(WX)
            X.axpy(-1., Q.lincomb(C))
            S[row:row + count] = C
            row += count
        return S

    def vector_loop(cols):
        """Classical Gram-Schmidt with re-iteration among the columns ``cols`` themselves (in place).  Returns the
        kept columns and the (len(kept), len(cols)) triangular factor."""
        ok, F = [], np.zeros((len(cols), len(cols)))
        pos = {c: i for i, c in enumerate(cols)}
        for a, c in enumerate(cols):
            v = A[c]
            norm0 = cur = v.norm(product)[0]
            if norm0 <= atol:
                dropped.append(c)
                continue
            while True:
                if ok:
                    wv = weigh(v)
                    for start, count in _column_runs(ok):
                        Q = A[start:start + count]
                        r = Q.inner(wv)[:, 0]
                        v.axpy(-1., Q.lincomb(r))
                        F[[pos[cc] for cc in range(start, start + count)], a] += r
                prev, cur = cur, v.norm(product)[0]
                if cur <= rtol * norm0:
                    logger.info('Removing linearly dependent vector %d', c)
                    dropped.append(c)
                    break
                if ok and cur < reiteration_threshold * prev:
                    continue
                v.scal(1 / cur)
                F[a, a] = cur
                ok.append(c)
                break
        return ok, F[[pos[c] for c in ok]]

    def orthonormalize(cols, level, before=()):
        """Orthonormalise the columns ``cols`` (in place) against the final columns ``before`` and among
        themselves with panels of ``panels[level]`` columns.  Returns the kept columns ``ok`` and the factor ``F`` of
        shape (len(before) + len(ok), len(cols)): ``A_cols(input) = A[before + ok] F``."""
        if level >= len(panels) or (not before and len(cols) <= panels[-1] and level > 0):
            assert not before
            return vector_loop(cols)
        b = panels[level]
        done = list(before)
        pos = {c: i for i, c in enumerate(cols)}
        F = np.zeros((len(before) + len(cols), len(cols)))       # rows: `before`, then one per column of `cols`
        row_of = {c: i for i, c in enumerate(before)}
        row_of.update({c: len(before) + i for i, c in enumerate(cols)})
        for p0 in range(0, len(cols), b):
            P = cols[p0:p0 + b]
            norms0 = view(P).norm(product)
            zero = [c for c, nv in zip(P, norms0) if nv <= atol]
            dropped.extend(zero)
            n0 = {c: nv for c, nv in zip(P, norms0)}
            P = [c for c in P if c not in zero]
            if not P:
                continue
            prev = list(done)
            if prev:
                S1 = project(P, prev)
                F[np.ix_([row_of[c] for c in prev], [pos[c] for c in P])] += S1
                gone = [c for c, nv in zip(P, view(P).norm(product)) if nv <= rtol * n0[c]]
                for c in gone:
                    logger.info('Removing linearly dependent vector %d', c)
                dropped.extend(gone)
                P = [c for c in P if c not in gone]
                if not P:
                    continue
            ok1, R1 = orthonormalize(P, level + 1)                       # X = Q1 R1
            if not ok1:
                continue
            if prev:
                S2 = project(ok1, prev)                                   # Q1 = Q S2 + Z
                ok2, R2 = orthonormalize(ok1, level + 1)                 # Z = Q2 R2
                F[np.ix_([row_of[c] for c in prev], [pos[c] for c in P])] += S2 @ R1
                F[np.ix_([row_of[c] for c in ok2], [pos[c] for c in P])] = R2 @ R1
            else:
                ok2 = ok1
                F[np.ix_([row_of[c] for c in ok2], [pos[c] for c in P])] = R1
            done.extend(ok2)
        ok = done[len(before):]
        return ok, F[[row_of[c] for c in list(before) + ok]]

    cols = list(range(offset, k))
    R = np.eye(k)
    if cols:
        before = list(range(offset))
        ok, F = orthonormalize(cols, 0, before)
        R[:, offset:] = 0.
        R[np.ix_(before + ok, cols)] = F
    dropped = sorted(set(dropped))
    if dropped:
        del A[dropped]
        R = np.delete(R, dropped, axis=0)
    if check:
        m = len(A) - offset
        err = A[offset:len(A)].inner(A, product)
        err[:m, offset:len(A)] -= np.eye(m)
        if err.size > 0:
            worst = np.max(np.abs(err))
            if worst >= check_tol:
                raise AccuracyError(f'result not orthogonal (max err={worst})')
    return (A, R) if return_R else A


def _select_modes(s, modes, rtol, atol, l2_err):
    above = np.nonzero(s >= max(rtol * s[0], atol))[0]
    if above.size == 0:
        return 0
    tail_energy = np.concatenate((np.cumsum(s[::-1] ** 2)[::-1], [0.]))
    n_err = int(np.nonzero(tail_energy <= l2_err ** 2)[0][0])
    count = min(n_err, int(above[-1]) + 1)
    return count if modes is None else min(count, modes)


def method_of_snapshots(A, product=None, modes=None, rtol=1e-7, atol=0., l2_err=0.):
    """SVD of ``A`` through the eigendecomposition of its (weighted) Gramian.  Returns
    ``U`` (VectorArray of left singular vectors), ``s``, ``Vh``.  The small eigenproblem stays on
    the host (LAPACK ``dsyevr`` via SciPy)."""
    if A.dim == 0 or len(A) == 0:
        return A.space.empty(), np.array([]), np.zeros((0, len(A)))
    k = len(A)
    G = A.gramian(product)
    window = None if (modes is None or l2_err > 0.) else (max(k - modes, 0), k - 1)
    evals, V = spla.eigh(G, overwrite_a=True, subset_by_index=window)
    evals, V = evals[::-1], V[:, ::-1]
    s = np.sqrt(np.clip(evals, a_min=0., a_max=None))
    m = _select_modes(s, modes, rtol, atol, l2_err)
    if m > A.dim:
        logger.warning('Number of computed singular vectors larger than array dimension! Truncating ...')
        m = A.dim
    s, V = s[:m], V[:, :m]
    return A.lincomb(V / s), s, V.conj().T


def qr_svd(A, product=None, modes=None, rtol=4e-8, atol=0., l2_err=0., gs_variant='mgs'):
    """SVD of ``A`` through Gram-Schmidt QR and the SVD of the small ``R`` factor.

    ``gs_variant`` (not in the reference): the Gram-Schmidt variant of :func:`gram_schmidt` used for the QR step.
    ``'bcgs2'`` runs it on the tensor cores (~4 n k^2 flop, O(k) host syncs) instead of k(k-1) host-synchronised
    pair-steps -- the POD that does NOT square the condition number then costs about as much as the method of
    snapshots."""
    if A.dim == 0 or len(A) == 0:
        return A.space.empty(), np.array([]), np.zeros((0, len(A)))
    Q, R = gram_schmidt(A, product=product, return_R=True, check=False, variant=gs_variant)
    U2, s, Vh = spla.svd(R, lapack_driver='gesvd')
    m = _select_modes(s, modes, rtol, atol, l2_err)
    return Q.lincomb(U2[:, :m]), s[:m], Vh[:m]


SVD_VA_METHODS = {'method_of_snapshots': method_of_snapshots, 'qr_svd': qr_svd}


def pod(A, product=None, modes=None, rtol=1e-7, atol=0., l2_err=0., method='qr_svd', orth_tol=1e-10,
        return_reduced_coefficients=False, **method_options):
    """Proper orthogonal decomposition of ``A``; re-orthonormalises the modes with
    :func:`gram_schmidt` when their orthogonality error is not below ``orth_tol``.  ``method_options`` (not in
    the reference) are handed to the SVD method, e.g. ``method='qr_svd', gs_variant='bcgs2'``."""
    assert method in SVD_VA_METHODS
    modes_va, svals, coeffs = SVD_VA_METHODS[method](A, product=product, modes=modes, rtol=rtol, atol=atol,
                                                     l2_err=l2_err, **method_options)
    if modes_va.dim > 0 and len(modes_va) > 0 and np.isfinite(orth_tol):
        err = np.max(np.abs(modes_va.inner(modes_va, product) - np.eye(len(modes_va))))
        if err >= orth_tol:
            logger.info('Reorthogonalizing POD modes ...')
            gram_schmidt(modes_va, product=product, atol=0., rtol=0., copy=False)
    if return_reduced_coefficients:
        return modes_va, svals, coeffs
    return modes_va, svals


def project(op, range_basis, source_basis, product=None):
    """Petrov-Galerkin projection of a linear, non-parametric operator.

    Returns the projected matrix ``(len(range_basis), len(source_basis))`` as a NumPy array when both
    bases are given (pyMOR wraps the same array in a ``NumpyMatrixOperator``); with only one basis the
    VectorArray ``op.apply_adjoint(range_basis)`` resp. ``op.apply(source_basis)`` is returned."""
    assert source_basis is None or source_basis in op.source
    assert range_basis is None or range_basis in op.range
    assert product is None or product.source == product.range == op.range
    if range_basis is None and source_basis is None:
        return op
    rb = product.apply(range_basis) if (product is not None and range_basis is not None) else range_basis
    if source_basis is None:
        return op.apply_adjoint(rb)
    if rb is None:
        return op.apply(source_basis)
    return op.apply2(rb, source_basis)


def _product_norm_estimate(product, iters=20):
    """Spectral norm of an SPD product operator by power iteration (the reference calls
    ``pymor.algorithms.eigs.eigs(product, k=1)``, chol_qr.py:185-187; any estimate of the same
    quantity serves: it only scales the diagonal shift)."""
    v = product.source.random(1, distribution='normal')
    lam = 1.0
    for _ in range(iters):
        v.scal(1.0 / v.norm()[0])
        v = product.apply(v)
        lam = v.norm()[0]
    return float(lam)


class _ShiftedCholesky:
    """Cholesky factor of a Gramian, shifted on breakdown (chol_qr.py:171-252)."""
    INNER_ITERS = 10

    def __init__(self, dim, product, product_norm, check_finite, recompute):
        self.dim, self.product, self.product_norm = dim, product, product_norm
        self.check_finite, self.recompute, self.shift = check_finite, recompute, None

    def _compute_shift(self, X):
        m, n = self.dim, len(X)
        eps = np.finfo(X.dtype).eps
        shift = 11 * eps
        if self.product is None:
            shift *= m * n + n * (n + 1)
        else:
            if self.product_norm is None:
                self.product_norm = np.sqrt(_product_norm_estimate(self.product))
            shift *= (2 * m * np.sqrt(m * n) + n * (n + 1)) * self.product_norm
        top = spla.eigh(X, eigvals_only=True, subset_by_index=[n - 1, n - 1], driver='evr')[0]
        return max(shift * top, eps)

    def apply(self, X):
        shift = 0.0
        for it in range(self.INNER_ITERS):
            try:
                M = X + np.eye(len(X)) * shift if self.recompute else X
                return spla.cholesky(M, overwrite_a=False, check_finite=self.check_finite)
            except spla.LinAlgError:
                logger.warning('Cholesky factorization broke down! Matrix is ill-conditioned.')
                if self.recompute:
                    shift = self._compute_shift(X) if it == 0 else shift * 10
                else:
                    if not self.shift:
                        self.shift = self._compute_shift(X)
                    X[np.diag_indices_from(X)] += self.shift
        raise AccuracyError('Failed to compute a Cholesky factorization of X.')


def shifted_chol_qr(A, product=None, return_R=False, maxiter=3, offset=0, orth_tol=None,
                    recompute_shift=False, check_finite=True, copy=True, product_norm=None):
    """Shifted CholeskyQR(3): Gramian -> Cholesky (diagonally shifted on breakdown) -> block
    lincomb with the inverse factor, repeated ``maxiter`` times.  Same arguments and behaviour as the
    reference; all n-sized work is GEMM-class (``inner``/``gramian``/``lincomb``/``append``)."""
    assert 0 <= offset <= len(A)
    assert A.dim >= len(A)
    assert 0 < maxiter
    assert orth_tol is None or 0 < orth_tol
    if copy:
        A = A.copy()
    k = len(A)
    if offset == k:
        return (A, np.eye(k)) if return_R else A
    chol = _ShiftedCholesky(A.dim, product, product_norm, check_finite, recompute_shift)

    def gram_blocks():
        G = A.inner(A[offset:], product=product)
        return G[:offset].astype(np.float64, copy=False), G[offset:].astype(np.float64, copy=False)

    B, X = gram_blocks()
    Bi = Ri = None
    it = 1
    while it <= maxiter:
        X = X - B.conj().T @ B
        Rx = chol.apply(X)
        Rinv = spla.solve_triangular(Rx, np.eye(len(Rx)), lower=False, check_finite=check_finite)
        todo = A[:offset].lincomb(-B @ Rinv) + A[offset:].lincomb(Rinv)
        del A[offset:]
        A.append(todo)
        if it == 1:
            Bi, Ri = B, Rx
        else:
            Bi = Bi + B @ Ri
            Ri = Rx @ Ri
        if it < maxiter:
            B, X = gram_blocks()
        elif orth_tol is not None:
            X = A[offset:].gramian(product=product)
        if orth_tol is not None:
            res = spla.norm(X - np.eye(k - offset), ord='fro', check_finite=check_finite)
            if res <= orth_tol * np.sqrt(k):
                break
            if it == maxiter:
                raise AccuracyError('Orthonormality could not be achieved within the given tolerance. '
                                    'Consider increasing maxiter or enabling recompute_shift.')
        it += 1
    if not return_R:
        return A
    R = np.eye(k)
    R[:offset, offset:] = Bi
    R[offset:, offset:] = Ri
    return A, R


def inc_hapod_chunks(chunks, steps, eps, omega, product=None):
    """Incremental HAPOD over snapshot chunks that arrive one at a time (``chunks``: an iterable of ``steps``
    |VectorArrays|, e.g. ``space.from_local_numpy(host_block)`` per block of a snapshot set that does not fit into
    HBM next to its modes; each chunk is consumed -- emptied -- by the node it feeds).  Same chain tree and local
    tolerances as :func:`inc_vectorarray_hapod` (src/pymor/algorithms/hapod.py:333-370, :414-424); only one chunk
    and the carried modes are resident at any time.  Returns ``modes, svals, snap_count``."""
    levels = int(steps)
    carried, snap_count = None, 0
    modes = svals = None
    idx = -1
    for idx, node in enumerate(chunks):
        assert idx < levels, 'more chunks than announced steps'
        snap_count += len(node)
        if carried is not None:
            node.append(carried, remove_from_other=True)
        is_root = idx == levels - 1
        if is_root:
            tol = np.sqrt(snap_count) * omega * eps
        else:
            tol = np.sqrt(snap_count) / np.sqrt(levels - 1) * np.sqrt(1 - omega ** 2) * eps
        modes, svals = pod(node, product=product, atol=0., rtol=0., l2_err=tol,
                           orth_tol=1e-10 if is_root else np.inf)
        del node[:]
        if not is_root:
            modes.scal(svals)
            carried = modes
    assert idx == levels - 1, 'fewer chunks than announced steps'
    return modes, svals, snap_count


def inc_vectorarray_hapod(steps, U, eps, omega, product=None):
    """Incremental hierarchical approximate POD ("next" row N3; reference:
    src/pymor/algorithms/hapod.py:333-370 over the chain tree of :95-104 with the local tolerances of
    :414-424): the snapshots are consumed in ``steps`` chunks, each intermediate node is one
    :func:`pod` of [new chunk, previous modes scaled by their singular values].  Returns
    ``modes, svals, snap_count``.  Lets POD run on snapshot sets that do not fit HBM at once."""
    chunk = -(-len(U) // steps)
    starts = list(range(0, len(U), chunk))
    levels = len(starts)                       # tree depth - 1 (no POD on the leaves)
    carried, snap_count = None, 0
    modes = svals = None
    for idx, s in enumerate(starts):
        node = U[s:s + chunk].copy()
        snap_count += len(node)
        if carried is not None:
            node.append(carried, remove_from_other=True)
        is_root = idx == levels - 1
        if is_root:
            tol = np.sqrt(snap_count) * omega * eps
        else:
            tol = np.sqrt(snap_count) / np.sqrt(levels - 1) * np.sqrt(1 - omega ** 2) * eps
        modes, svals = pod(node, product=product, atol=0., rtol=0., l2_err=tol,
                           orth_tol=1e-10 if is_root else np.inf)
        if not is_root:
            modes.scal(svals)
            carried = modes
    return modes, svals, snap_count


def dist_vectorarray_hapod(num_slices, U, eps, omega, arity=None, product=None):
    """Distributed hierarchical approximate POD ("next" row N3; reference:
    src/pymor/algorithms/hapod.py:372-410 over the tree of :106-126 with the local tolerances of :414-424):
    the snapshots are cut into ``num_slices`` slices, every leaf is the :func:`pod` of one slice, every inner
    node the :func:`pod` of its children's modes scaled by their singular values; ``arity=None`` gives the
    two-level tree (one POD per slice, one of the collected modes).  Returns ``modes, svals, snap_count``.
    The nodes run one after the other on the device (the reference's ``executor`` only adds host threads)."""
    chunk = -(-len(U) // num_slices)
    starts = list(range(0, len(U), chunk))
    fan = len(starts) if arity is None else arity

    def build(slices):                         # nested lists; an int is a leaf (slice number)
        if len(slices) > fan:
            return [build(list(part)) if len(part) > 1 else int(part[0]) for part in np.array_split(slices, fan)]
        return [int(i) for i in slices]

    def depth(node):
        return 1 if isinstance(node, int) else 1 + max(depth(c) for c in node)

    tree = build(list(range(len(starts))))
    levels = depth(tree)

    def local_pod(V, snap_count, is_root):
        if is_root:
            tol = np.sqrt(snap_count) * omega * eps
        else:
            tol = np.sqrt(snap_count) / np.sqrt(levels - 1) * np.sqrt(1 - omega ** 2) * eps
        return pod(V, product=product, atol=0., rtol=0., l2_err=tol, orth_tol=1e-10 if is_root else np.inf)

    def visit(node, is_root):
        if isinstance(node, int):
            V = U[starts[node]:starts[node] + chunk].copy()
            count = len(V)
        else:
            V, count = None, 0
            for child in node:
                modes, svals, c = visit(child, False)
                modes.scal(svals)
                count += c
                if V is None:
                    V = modes
                else:
                    V.append(modes, remove_from_other=True)
        modes, svals = local_pod(V, count, is_root)
        return modes, svals, count

    return visit(tree, True)


def randomized_range_finder(A, basis_size, power_iterations=0, range_product=None, qr_method='gram_schmidt'):
    """Orthonormal basis of the (approximate) range of the operator ``A`` from ``basis_size`` random
    samples ("next" row N2; reference: src/pymor/algorithms/rand_la.py:20-252, fixed-size mode):
    ``Q = orth(A Omega)``, then ``power_iterations`` times ``Q = orth(A A^T Q)``.  Uses only
    ``Operator.apply`` / ``apply_adjoint``, ``space.random`` and the orthonormalisation."""
    def orth(V):
        if qr_method == 'shifted_chol_qr':
            return shifted_chol_qr(V, product=range_product, copy=False)
        return gram_schmidt(V, product=range_product, atol=0., rtol=0., copy=False)

    Q = orth(A.apply(A.source.random(basis_size, distribution='normal')))
    for _ in range(power_iterations):
        Q = orth(A.apply(A.apply_adjoint(Q)))
    return Q


def ei_greedy(U, error_norm=None, atol=None, rtol=None, max_interpolation_dofs=None, nodal_basis=False,
              copy=True):
    """Interpolation DOFs and collateral basis by the EI-greedy search ("next" row N5; reference:
    src/pymor/algorithms/ei.py:30-179, sequential branch).  In every round the worst approximated
    vector becomes a basis vector, normalised to 1 at its largest entry (the new interpolation DOF),
    and is eliminated from all vectors at that DOF.  Per round the backend runs one ``amax``, two
    ``dofs``, one ``axpy`` with per-column coefficients over the whole array (24 n k bytes) and one
    ``norm`` / ``sup_norm`` (8 n k bytes): HBM-bound, one host sync per reduction.

    ``error_norm``: ``None`` (Euclidean), ``'sup'`` or a callable ``VectorArray -> ndarray``.
    Returns ``interpolation_dofs, collateral_basis, data`` with the reference's ``data`` keys."""
    assert not isinstance(error_norm, str) or error_norm == 'sup'
    assert len(U) > 0
    if error_norm is None:
        def measure(V):
            return V.norm()
    elif error_norm == 'sup':
        def measure(V):
            return V.sup_norm()
    else:
        measure = error_norm
    k = len(U)
    dofs = np.zeros((0,), dtype=np.int32)
    basis = U.space.empty()
    K = np.eye(k)                                  # U_current = U_initial.lincomb(K)
    coefficients = np.zeros((k, 0))
    max_errs = []
    if copy:
        U = U.copy()
    errs = measure(U)
    worst = int(np.argmax(errs))
    first_err = err = errs[worst]
    while True:
        if max_interpolation_dofs is not None and len(dofs) >= max_interpolation_dofs:
            logger.info('Maximum number of interpolation DOFs reached (error %g)', err)
            break
        if atol is not None and err <= atol:
            break
        if rtol is not None and err / first_err <= rtol:
            break
        vec = U[worst].copy()
        dof = vec.amax()[0][0]
        if dof in dofs:
            logger.info('DOF %d selected twice for interpolation, stopping', dof)
            break
        pivot = vec.dofs([dof])[0, 0]
        if pivot == 0.:
            logger.info('DOF %d selected for interpolation has zero maximum error, stopping', dof)
            break
        vec.scal(1 / pivot)
        dofs = np.hstack((dofs, dof))
        basis.append(vec)
        coefficients = np.hstack([coefficients, K[:, worst:worst + 1] / pivot])
        max_errs.append(err)
        row = U.dofs([dof])
        U.axpy(-row[0, :], vec)
        K -= K[:, worst:worst + 1] @ (row / pivot)
        errs = measure(U)
        worst = int(np.argmax(errs))
        err = errs[worst]
    matrix = basis.dofs(dofs)
    tri = np.abs(matrix - np.tril(matrix))
    tri_errs = [np.max(tri[:d, :d]) for d in range(1, len(matrix) + 1)]
    if nodal_basis:
        inv = spla.inv(matrix)
        basis = basis.lincomb(inv)
        coefficients = coefficients @ inv
        matrix = np.eye(len(basis))
    return dofs, basis, {'errors': max_errs, 'triangularity_errors': tri_errs, 'coefficients': coefficients,
                         'interpolation_matrix': matrix}


def deim(U, modes=None, pod=True, atol=None, rtol=None, product=None, pod_options={}):
    """Interpolation DOFs and collateral basis by DEIM ("next" row N5; reference:
    src/pymor/algorithms/ei.py:182-264): POD of ``U`` (unless ``pod=False``), then for every mode the
    DOF where its interpolation residual w.r.t. the previous modes is largest.  Returns
    ``interpolation_dofs, collateral_basis, data`` (``data['svals']`` when the POD ran)."""
    data = {}
    if pod:
        kw = dict(pod_options)
        if atol is not None:
            kw['atol'] = atol
        if rtol is not None:
            kw['rtol'] = rtol
        basis, svals = globals()['pod'](U, modes=modes, product=product, **kw)
        data['svals'] = svals
    else:
        basis = U
    dofs = np.zeros((0,), dtype=np.int32)
    matrix = np.zeros((0, 0))
    for i in range(len(basis)):
        if len(dofs) > 0:
            coeff = spla.solve(matrix, basis[i].dofs(dofs))
            resid = basis[i] - basis[:len(dofs)].lincomb(coeff)
        else:
            resid = basis[i]
        dof = resid.amax()[0][0]
        if dof in dofs:
            logger.info('DOF %d selected twice for interpolation, stopping', dof)
            break
        dofs = np.hstack((dofs, dof))
        matrix = basis[:len(dofs)].dofs(dofs)
    if len(dofs) < len(basis):
        del basis[len(dofs):len(basis)]
    return dofs, basis, data
