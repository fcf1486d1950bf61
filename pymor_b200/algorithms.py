"""Extensions of pyMOR's hot-path algorithms for device-resident arrays.

pyMOR's own functions are the callers of this backend and run on :class:`~pymor_b200.vectorarray.CudaVectorArray`
unchanged; this module re-exports them under their reference names so that `pymor_b200.algorithms.<name>` IS the
reference function (no restatement):

* :func:`method_of_snapshots`   -- pymor.algorithms.svd_va (src/pymor/algorithms/svd_va.py:20-99)
* :func:`project`               -- pymor.algorithms.projection (src/pymor/algorithms/projection.py:28-85)
* :func:`shifted_chol_qr`       -- pymor.algorithms.chol_qr (src/pymor/algorithms/chol_qr.py:16-252; "next" row N1)
* :func:`inc_vectorarray_hapod`, :func:`dist_vectorarray_hapod` -- pymor.algorithms.hapod (:333, :372; N3)
* :func:`ei_greedy`, :func:`deim` -- pymor.algorithms.ei (:30, :182; N5)
* ``RandomizedRangeFinder`` / ``randomized_svd`` -- pymor.algorithms.rand_la (:20, :378; N2)

What is defined HERE is what the reference does not have:

* :func:`gram_schmidt` -- pyMOR's ``gram_schmidt`` (src/pymor/algorithms/gram_schmidt.py:13-122; called directly for
  ``variant='mgs'`` without ``sweep``) plus ``sweep=True`` (one fused device kernel per pass instead of one launch
  and one host synchronisation per pair) and ``variant='cgs2'`` / ``'bcgs2'`` (:func:`block_gram_schmidt`): the
  classical / block-classical Gram-Schmidt variants with re-orthogonalisation that turn the O(k^2)
  host-synchronised pair-steps of gram_schmidt.py:84-91 into O(k) streaming resp. tensor-core calls; same contract
  and (the QR factorisation being unique) the same results to rounding x conditioning
* :func:`qr_svd` / :func:`pod` -- pyMOR's functions, with the option to run the QR step with one of those variants
* :func:`inc_hapod_chunks` -- incremental HAPOD over snapshot chunks streamed from host memory
"""
from __future__ import annotations

import logging
import os

import numpy as np
import scipy.linalg as spla

from pymor_b200 import interface as _interface  # noqa: F401  (puts pyMOR on sys.path)
from pymor.algorithms import gram_schmidt as _ref_gs  # noqa: E402
from pymor.algorithms import pod as _ref_pod  # noqa: E402
from pymor.algorithms import svd_va as _ref_svd_va  # noqa: E402
from pymor.algorithms.chol_qr import shifted_chol_qr  # noqa: E402,F401
from pymor.algorithms.ei import deim, ei_greedy  # noqa: E402,F401
from pymor.algorithms.hapod import dist_vectorarray_hapod, inc_vectorarray_hapod  # noqa: E402,F401
from pymor.algorithms.projection import project  # noqa: E402,F401
from pymor.algorithms.rand_la import RandomizedRangeFinder, randomized_svd  # noqa: E402,F401
from pymor.algorithms.svd_va import method_of_snapshots  # noqa: E402,F401
from pymor.core.exceptions import AccuracyError  # noqa: E402,F401

# pyMOR's own function, also while accelerate_pymor() has the module attribute point to its dispatcher
_pymor_gram_schmidt = getattr(_ref_gs.gram_schmidt, '__wrapped__', _ref_gs.gram_schmidt)

logger = logging.getLogger('pymor_b200.algorithms')

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
    if sweep is None:
        sweep = os.environ.get('PMC_MGS_SWEEP', '0') == '1'
    if variant == 'mgs' and not sweep:
        return _pymor_gram_schmidt(A, product=product, return_R=return_R, atol=atol, rtol=rtol, offset=offset,
                                    reiterate=reiterate, reiteration_threshold=reiteration_threshold, check=check,
                                    check_tol=check_tol, copy=copy)
    if variant == 'bcgs2':
        return block_gram_schmidt(A, product=product, return_R=return_R, atol=atol, rtol=rtol, offset=offset,
                                  reiteration_threshold=reiteration_threshold, check=check, check_tol=check_tol,
                                  copy=copy)
    if copy:
        A = A.copy()
    k = len(A)
    R = np.eye(k)
    dropped = []
    weight_ptr = _sweep_weight(A, product) if sweep else None
    if weight_ptr is not None:
        # the sweeps write through the raw impl: un-share it first (copy() above is lazy, and with copy=False other
        # arrays may hold lazy copies of A) -- what VectorArray-level writes do via _copy_impl_if_multiple_refs
        A._copy_impl_if_multiple_refs()
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


def _holds_complex(A):
    """Whether the array stores complex data (device arrays: a pair of real panels; NumPy arrays: the dtype)."""
    impl = getattr(A, 'impl', None)
    if type(impl).__name__ == 'CudaComplexVectorArrayImpl':
        return True
    arr = getattr(impl, '_array', None)
    return arr is not None and np.iscomplexobj(arr)


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
    if _holds_complex(A):
        # the host factors S / F / R below are real; complex data (a contract-completeness path, not a performance
        # path: DESIGN.md §8) takes the classical variant, which promotes R on the first complex coefficient
        return gram_schmidt(A, product=product, return_R=return_R, atol=atol, rtol=rtol, offset=offset,
                            reiteration_threshold=reiteration_threshold, check=check, check_tol=check_tol, copy=copy,
                            variant='cgs2', sweep=False)
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


def qr_svd(A, product=None, modes=None, rtol=4e-8, atol=0., l2_err=0., gs_variant='mgs'):
    """SVD of ``A`` through Gram-Schmidt QR and the SVD of the small ``R`` factor: pyMOR's ``qr_svd``
    (src/pymor/algorithms/svd_va.py:103-168), called directly for ``gs_variant='mgs'``.

    ``gs_variant`` (not in the reference): the Gram-Schmidt variant of :func:`gram_schmidt` used for the QR step.
    ``'bcgs2'`` runs it on the tensor cores (~4 n k^2 flop, O(k) host syncs) instead of k(k-1) host-synchronised
    pair-steps -- the POD that does NOT square the condition number then costs about as much as the method of
    snapshots."""
    if gs_variant == 'mgs':
        return _ref_svd_va.qr_svd(A, product=product, modes=modes, rtol=rtol, atol=atol, l2_err=l2_err)
    if A.dim == 0 or len(A) == 0:
        return A.space.empty(), np.array([]), np.zeros((0, len(A)))
    Q, R = gram_schmidt(A, product=product, return_R=True, check=False, variant=gs_variant)
    U2, s, Vh = spla.svd(R, lapack_driver='gesvd')
    m = _ref_svd_va._select_modes(s, modes, rtol, atol, l2_err)
    return Q.lincomb(U2[:, :m]), s[:m], Vh[:m]


def pod(A, product=None, modes=None, rtol=1e-7, atol=0., l2_err=0., method='qr_svd', orth_tol=1e-10,
        return_reduced_coefficients=False, **method_options):
    """pyMOR's ``pod`` (src/pymor/algorithms/pod.py:16-90), called directly unless ``method_options`` (not in the
    reference) are given: those are handed to the SVD method, e.g. ``method='qr_svd', gs_variant='bcgs2'``, and the
    orthonormality check / re-orthonormalisation of pod.py:78-86 follows as in the reference."""
    if not method_options:
        return _ref_pod.pod(A, product=product, modes=modes, rtol=rtol, atol=atol, l2_err=l2_err, method=method,
                            orth_tol=orth_tol, return_reduced_coefficients=return_reduced_coefficients)
    svd = {'method_of_snapshots': method_of_snapshots, 'qr_svd': qr_svd}[method]
    modes_va, svals, coeffs = svd(A, product=product, modes=modes, rtol=rtol, atol=atol, l2_err=l2_err,
                                  **method_options)
    if modes_va.dim > 0 and len(modes_va) > 0 and np.isfinite(orth_tol):
        err = np.max(np.abs(modes_va.inner(modes_va, product) - np.eye(len(modes_va))))
        if err >= orth_tol:
            logger.info('Reorthogonalizing POD modes ...')
            gram_schmidt(modes_va, product=product, atol=0., rtol=0., copy=False)
    if return_reduced_coefficients:
        return modes_va, svals, coeffs
    return modes_va, svals


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
