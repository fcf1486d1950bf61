"""CPU oracle for the snapshot linear-algebra hot path -- TEST INFRASTRUCTURE ONLY.

A plain-NumPy restatement of what pyMOR's CPU backend computes on this path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product (``pymor_b200/``) never does.

Parity status: PINNED.  The arithmetic of the reference lives in NumPy/OpenBLAS and SciPy/LAPACK
(numpy 2.3.5 / scipy 1.18.1 in this image; not vendored under /root/reference).  This restatement is
checked (tests/test_oracle_pinned.py) against golden vectors produced by importing the reference
itself in the build container (tests/golden/generate.py, fixtures tests/golden/*.npz) and, when
/root/reference is present, against the live reference on fresh random inputs.

Data convention (reference: src/pymor/vectorarrays/numpy.py:16-18, :264): a vector array is a
``(dim, len)`` ndarray, one vector per column.  A product ``W`` is ``None`` (Euclidean), a 1-D
ndarray (diagonal matrix) or anything supporting ``W @ X`` (dense ndarray / scipy.sparse matrix), which
is how ``NumpyMatrixOperator.apply`` evaluates it (src/pymor/operators/numpy.py:232-237).

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as spla


# ------------------------------------------------------------------------------------------------
# VectorArray primitives (src/pymor/vectorarrays/numpy.py)
# ------------------------------------------------------------------------------------------------
def apply_product(W, X):
    """``product.apply(X)``: W·X (src/pymor/operators/numpy.py:232-237)."""
    if W is None:
        return X
    if isinstance(W, np.ndarray) and W.ndim == 1:
        return W[:, None] * X
    return np.asarray(W @ X)


def inner(A, B, W=None):
    """``A.inner(B, product)``: G[i, j] = conj(a_i)·(W b_j).

    numpy.py:155-163 (``matmul(A.conj().T, B)``); with a product interface.py:432-436 defers to
    ``Operator.apply2`` = ``V.inner(op.apply(U))`` (operators/interface.py:79-109)."""
    return np.matmul(A.conj().T, apply_product(W, B))


def gramian(A, W=None):
    """``A.gramian(product)`` (interface.py:661-666, :1068-1069)."""
    return inner(A, A, W)


def pairwise_inner(A, B, W=None):
    """``A.pairwise_inner(B, product)``: r[i] = conj(a_i)·(W b_i).

    numpy.py:165-173; product case operators/interface.py:111-143."""
    return np.sum(A.conj() * apply_product(W, B), axis=0)


def lincomb(A, C):
    """``A.lincomb(C)``: columns of A·C, C of shape (len(A), m) (numpy.py:175-180)."""
    C = np.asarray(C)
    if C.ndim == 1:
        C = C[:, None]          # interface.py:516-518
    return np.matmul(A, C)


def norm2(A, W=None):
    """Squared norms (numpy.py:189-194; product: interface.py:597-601)."""
    if W is None:
        return np.sum((A * A.conj()).real, axis=0)
    return pairwise_inner(A, A, W).real


def norm(A, W=None):
    """Norms (numpy.py:182-187; product: interface.py:553-555)."""
    if W is None:
        return np.linalg.norm(A, axis=0)
    return np.sqrt(norm2(A, W))


def axpy(Y, alpha, X):
    """In-place ``Y += alpha * X`` with scalar or per-column alpha, X of len(Y) or 1 column
    (numpy.py:112-138).  Returns Y (possibly a promoted copy when dtypes differ)."""
    alpha_arr = np.asarray(alpha)
    dt = np.result_type(Y.dtype, alpha_arr.dtype, X.dtype)
    if dt != Y.dtype:
        Y = Y.astype(dt, order='F')
    if alpha_arr.ndim == 1:
        Y += X * alpha_arr[None, :]
    else:
        Y += X * alpha
    return Y


def scal(Y, alpha):
    """In-place ``Y *= alpha`` (numpy.py:85-100)."""
    alpha_arr = np.asarray(alpha)
    dt = np.result_type(Y.dtype, alpha_arr.dtype)
    if dt != Y.dtype:
        Y = Y.astype(dt, order='F')
    if alpha_arr.ndim == 1:
        Y *= alpha_arr[None, :]
    else:
        Y *= alpha
    return Y


def dofs(A, dof_indices):
    """``A.dofs(idx)`` -> (len(idx), len(A)) (numpy.py:196-201)."""
    return np.asfortranarray(A[np.asarray(dof_indices, dtype=np.int64), :])


def amax(A):
    """First index of the maximal absolute value per column and that value (numpy.py:203-211)."""
    absA = np.abs(A)
    idx = np.argmax(absA, axis=0)
    return idx, absA[idx, np.arange(A.shape[1])]


def apply2(M, V, U):
    """Two-form ``op.apply2(V, U) = V^H (M U)`` (operators/interface.py:79-109)."""
    return inner(V, U, M)


def sharded_reduce(partials):
    """Sum of per-shard partial results: what ``MPIVectorArrayAutoComm`` does with the gathered
    blocks on rank 0 (src/pymor/vectorarrays/mpi.py:394-416)."""
    return np.sum(np.stack(partials), axis=0)


def sharded_amax(inds, vals, offsets):
    """Merge per-shard amax results; the lowest shard wins ties (mpi.py:337-347)."""
    inds = np.stack(inds) + np.asarray(offsets)[:, None]
    vals = np.stack(vals)
    win = np.argmax(vals, axis=0)
    cols = np.arange(vals.shape[1])
    return inds[win, cols], vals[win, cols]


# ------------------------------------------------------------------------------------------------
# gram_schmidt (src/pymor/algorithms/gram_schmidt.py:13-122)
# ------------------------------------------------------------------------------------------------
class AccuracyError(Exception):
    pass


def gram_schmidt(A, W=None, return_R=False, atol=1e-13, rtol=1e-13, offset=0, reiterate=True,
                 reiteration_threshold=9e-1, check=True, check_tol=1e-3, stats=None):
    """Modified Gram-Schmidt with re-iteration.  ``A`` (dim, len) is not modified; returns Q
    (and R).  ``stats`` (dict) receives the number of pair-steps and passes, used by bench.py's byte
    model."""
    A = np.array(A, order='F', copy=True)                 # gram_schmidt.py:57-58 (copy=True)
    k = A.shape[1]
    R = np.eye(k)                                         # :61
    removed = []
    pair_steps = 0
    reiterations = 0
    for i in range(offset, k):                            # :63
        first_norm = norm(A[:, i:i + 1], W)[0]            # :65
        if first_norm <= atol:                            # :67-70
            removed.append(i)
            continue
        if i == 0:                                        # :72-74
            A[:, 0] *= 1 / first_norm
            R[0, 0] = first_norm
            continue
        cur = first_norm
        while True:                                       # :79
            for j in range(i):                            # :81
                if j in removed:
                    continue
                p = pairwise_inner(A[:, j:j + 1], A[:, i:i + 1], W)[0]   # :84
                A[:, i] -= p * A[:, j]                    # :85
                R = R.astype(np.promote_types(R.dtype, type(p)), copy=False)
                R[j, i] += p                              # :86-88
                pair_steps += 1
            prev, cur = cur, norm(A[:, i:i + 1], W)[0]    # :91
            if cur <= rtol * first_norm:                  # :94-97
                removed.append(i)
                break
            if reiterate and cur < reiteration_threshold * prev:   # :100-101
                reiterations += 1
                continue
            A[:, i] *= 1 / cur                            # :103-105
            R[i, i] = cur
            break
    if removed:                                           # :107-109
        keep = [c for c in range(k) if c not in removed]
        A = np.asfortranarray(A[:, keep])
        R = np.delete(R, removed, axis=0)
    if check:                                             # :111-117
        err = inner(A[:, offset:], A, W)
        m = A.shape[1] - offset
        err[:m, offset:A.shape[1]] -= np.eye(m)
        if err.size > 0 and np.max(np.abs(err)) >= check_tol:
            raise AccuracyError(f'result not orthogonal (max err={np.max(np.abs(err))})')
    if stats is not None:
        stats.update(pair_steps=pair_steps, reiterations=reiterations, removed=list(removed))
    return (A, R) if return_R else A


# ------------------------------------------------------------------------------------------------
# SVDs of a vector array (src/pymor/algorithms/svd_va.py)
# ------------------------------------------------------------------------------------------------
def select_modes(s, modes, rtol, atol, l2_err):
    """Number of singular values to keep (svd_va.py:245-259)."""
    tol = max(rtol * s[0], atol)
    keep = np.nonzero(s >= tol)[0]
    if keep.size == 0:
        return 0
    tail = np.concatenate((np.cumsum(s[::-1] ** 2)[::-1], [0.]))
    first_ok = np.nonzero(tail <= l2_err ** 2)[0][0]
    count = min(first_ok, keep[-1] + 1)
    return count if modes is None else min(count, modes)


def method_of_snapshots(A, W=None, modes=None, rtol=1e-7, atol=0., l2_err=0.):
    """Gram matrix -> symmetric eigenproblem -> modes (svd_va.py:20-99).  Returns U, s, Vh."""
    dim, k = A.shape
    if dim == 0 or k == 0:                                # :69-70
        return np.zeros((dim, 0)), np.array([]), np.zeros((0, k))
    G = gramian(A, W)                                     # :75
    subset = None if (modes is None or l2_err > 0.) else (max(k - modes, 0), k - 1)   # :78-80
    evals, V = spla.eigh(G, subset_by_index=subset)       # :82
    evals, V = evals[::-1], V[:, ::-1]                    # :83-84
    s = np.sqrt(np.clip(evals, 0., None))                 # :85
    m = min(select_modes(s, modes, rtol, atol, l2_err), dim)   # :86-89
    s, V = s[:m], V[:, :m]
    return lincomb(A, V / s), s, V.conj().T               # :97


def qr_svd(A, W=None, modes=None, rtol=4e-8, atol=0., l2_err=0.):
    """Gram-Schmidt QR -> SVD of R -> modes (svd_va.py:103-168)."""
    dim, k = A.shape
    if dim == 0 or k == 0:
        return np.zeros((dim, 0)), np.array([]), np.zeros((0, k))
    Q, R = gram_schmidt(A, W, return_R=True, check=False)       # :154
    U2, s, Vh = spla.svd(R, lapack_driver='gesvd')        # :157 (bindings/scipy.py:40-54)
    m = select_modes(s, modes, rtol, atol, l2_err)        # :160
    return lincomb(Q, U2[:, :m]), s[:m], Vh[:m]           # :161-166


def pod(A, W=None, modes=None, rtol=1e-7, atol=0., l2_err=0., method='method_of_snapshots',
        orth_tol=1e-10, stats=None):
    """POD driver (src/pymor/algorithms/pod.py:16-90).  Returns modes, singular values.

    Note the reference passes ITS rtol default (1e-7) to whichever SVD method is chosen (pod.py:15-17)."""
    svd = {'method_of_snapshots': method_of_snapshots, 'qr_svd': qr_svd}[method]
    U, s, _ = svd(A, W, modes=modes, rtol=rtol, atol=atol, l2_err=l2_err)       # :79
    reorth = False
    if U.shape[0] > 0 and U.shape[1] > 0 and np.isfinite(orth_tol):             # :80
        err = np.max(np.abs(inner(U, U, W) - np.eye(U.shape[1])))               # :82
        if err >= orth_tol:                                                     # :83-86
            U = gram_schmidt(U, W, atol=0., rtol=0.)
            reorth = True
        if stats is not None:
            stats.update(orth_err=err, reorthogonalized=reorth)
    return U, s


def project(M, range_basis, source_basis, W=None):
    """Petrov-Galerkin projection of a linear operator given as matrix ``M``
    (src/pymor/algorithms/projection.py:28-85, rule ``action_apply_basis`` :118-143):
    both bases -> ``apply2(rb, sb)`` with ``rb = product.apply(range_basis)``;
    only range basis -> ``apply_adjoint(rb)^H``; only source basis -> ``apply(sb)``."""
    rb = apply_product(W, range_basis) if (W is not None and range_basis is not None) else range_basis
    if source_basis is None and rb is None:
        return M
    if source_basis is None:
        return np.asarray((M.T.conj() @ rb)).conj().T      # :123-127
    if rb is None:
        return np.asarray(M @ source_basis)                # :134-136
    return apply2(M, rb, source_basis)                     # :143


def shifted_chol_qr(A, W=None, return_R=False, maxiter=3, offset=0, product_norm=None, recompute_shift=False):
    """Shifted CholeskyQR3 (src/pymor/algorithms/chol_qr.py:16-168; kernels :205-252; "next" row N1 of
    SURVEY.md §8f).  ``product_norm`` must be given when ``W`` is.  Note: in the reference the default
    (non-recomputing) kernel cannot actually apply a shift -- it raises AttributeError at :227 -- so the
    shifted branch is only pinned for ``recompute_shift=True``."""
    A = np.array(A, order='F', copy=True)
    dim, k = A.shape
    if offset == k:
        return (A, np.eye(k)) if return_R else A

    def gram_blocks():                                     # chol_qr.py:159-168
        G = inner(A, A[:, offset:], W)
        return G[:offset], G[offset:]

    def shift_for(X):                                      # chol_qr.py:178-203
        n = len(X)
        eps = np.finfo(np.float64).eps
        shift = 11 * eps
        shift *= (dim * n + n * (n + 1)) if W is None else \
            (2 * dim * np.sqrt(dim * n) + n * (n + 1)) * product_norm
        top = spla.eigh(X, eigvals_only=True, subset_by_index=[n - 1, n - 1])[0]
        return max(shift * top, eps)

    B, X = gram_blocks()
    shift = None
    Bi = Ri = None
    for it in range(1, maxiter + 1):                       # chol_qr.py:113-146
        X = X - B.T @ B
        cur = 0.0
        for attempt in range(10):                          # kernels :208-224 / :232-252
            try:
                Rx = spla.cholesky(X + np.eye(len(X)) * cur if recompute_shift else X)
                break
            except spla.LinAlgError:
                if recompute_shift:
                    cur = shift_for(X) if attempt == 0 else cur * 10
                else:
                    if not shift:
                        shift = shift_for(X)
                    X[np.diag_indices_from(X)] += shift
        else:
            raise AccuracyError('Failed to compute a Cholesky factorization of X.')
        Rinv = spla.solve_triangular(Rx, np.eye(len(Rx)))
        A[:, offset:] = A[:, :offset] @ (-B @ Rinv) + A[:, offset:] @ Rinv
        if it == 1:
            Bi, Ri = B, Rx
        else:
            Bi = Bi + B @ Ri
            Ri = Rx @ Ri
        if it < maxiter:
            B, X = gram_blocks()
    if not return_R:
        return A
    R = np.eye(k)
    R[:offset, offset:] = Bi
    R[offset:, offset:] = Ri
    return A, R


# ------------------------------------------------------------------------------------------------
# parity metrics named by the north star
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# empirical interpolation ("next" row N5; src/pymor/algorithms/ei.py)
# ------------------------------------------------------------------------------------------------
def ei_greedy(U, error_norm=None, atol=None, rtol=None, max_interpolation_dofs=None, nodal_basis=False):
    """EI-greedy (ei.py:30-179) on a (dim, len) matrix; ``error_norm`` None (Euclidean), 'sup' or a
    callable ``(dim, len) matrix -> (len,) norms`` (e.g. ``lambda E: norm(E, W)`` for a product norm).
    Returns interpolation DOFs, collateral basis (dim, m) and the data dict of the reference."""
    U = np.array(U, order='F', dtype=float)
    k = U.shape[1]
    if callable(error_norm):
        measure = error_norm
    else:
        measure = (lambda E: np.max(np.abs(E), axis=0)) if error_norm == 'sup' else (lambda E: norm(E))
    picked, basis = [], np.zeros((U.shape[0], 0), order='F')
    K, coefficients = np.eye(k), np.zeros((k, 0))
    max_errs = []
    errs = measure(U)                                                       # :114
    worst = int(np.argmax(errs))
    first_err = err = errs[worst]
    while True:
        if max_interpolation_dofs is not None and len(picked) >= max_interpolation_dofs:   # :120
            break
        if atol is not None and err <= atol:                                # :129
            break
        if rtol is not None and err / first_err <= rtol:                    # :133
            break
        vec = U[:, worst].copy()                                            # :138
        dof = int(amax(vec[:, None])[0][0])
        if dof in picked:                                                   # :140
            break
        pivot = vec[dof]
        if pivot == 0.:                                                     # :144
            break
        vec *= 1 / pivot                                                    # :147
        picked.append(dof)
        basis = np.asfortranarray(np.hstack([basis, vec[:, None]]))
        coefficients = np.hstack([coefficients, K[:, worst:worst + 1] / pivot])
        max_errs.append(err)
        row = U[dof:dof + 1, :].copy()                                      # :154  U.dofs([dof])
        U = axpy(U, -row[0], vec[:, None])                                  # :155
        K -= K[:, worst:worst + 1] @ (row / pivot)
        errs = measure(U)
        worst = int(np.argmax(errs))
        err = errs[worst]
    dofs_ = np.array(picked, dtype=np.int64)
    matrix = dofs(basis, dofs_)                                             # :161
    tri = np.abs(matrix - np.tril(matrix))
    tri_errs = [np.max(tri[:d, :d]) for d in range(1, len(matrix) + 1)]
    if nodal_basis:                                                         # :169-174
        inv = spla.inv(matrix)
        basis = lincomb(basis, inv)
        coefficients = coefficients @ inv
        matrix = np.eye(basis.shape[1])
    return dofs_, basis, {'errors': max_errs, 'triangularity_errors': tri_errs, 'coefficients': coefficients,
                          'interpolation_matrix': matrix}


def deim(U, W=None, modes=None, use_pod=True, atol=None, rtol=None):
    """DEIM (ei.py:182-264): POD of the snapshots (reference defaults: ``qr_svd``), then one interpolation DOF per
    mode at the largest entry of its interpolation residual."""
    data = {}
    if use_pod:
        kw = {}
        if atol is not None:
            kw['atol'] = atol
        if rtol is not None:
            kw['rtol'] = rtol
        basis, svals = pod(np.asarray(U), W, modes=modes, method='qr_svd', **kw)   # :229
        data['svals'] = svals
    else:
        basis = np.array(U, order='F', dtype=float)
    picked, matrix = [], np.zeros((0, 0))
    for i in range(basis.shape[1]):
        if picked:                                                          # :241-245
            coeff = spla.solve(matrix, basis[picked, i])
            resid = basis[:, i] - basis[:, :len(picked)] @ coeff
        else:
            resid = basis[:, i]
        dof = int(amax(resid[:, None])[0][0])                               # :250
        if dof in picked:
            break
        picked.append(dof)
        matrix = dofs(basis[:, :len(picked)], picked)                       # :257
    return np.array(picked, dtype=np.int64), np.asfortranarray(basis[:, :len(picked)]), data


def subspace_angle(U, V, W=None):
    """Largest principal angle between span(U) and span(V) (columns W-orthonormal)."""
    if U.shape[1] == 0 and V.shape[1] == 0:
        return 0.0
    # sin of the largest angle = || (I - U U^T W) V ||_W,2  (accurate for small angles)
    Rm = V - U @ inner(U, V, W)
    G = inner(Rm, Rm, W)
    smax = np.sqrt(max(np.linalg.eigvalsh((G + G.T) / 2)[-1], 0.0))
    return float(np.arcsin(min(smax, 1.0)))


def orthogonality_error(U, W=None):
    return float(np.max(np.abs(inner(U, U, W) - np.eye(U.shape[1])))) if U.shape[1] else 0.0


def gram_rel_error(G, G_ref, A, B, W=None):
    """max |dG_ij| / (||a_i||_W ||b_j||_W): the entrywise-relative measure of SURVEY.md §8c."""
    na, nb = norm(A, W), norm(B, W)
    scale = np.outer(na, nb)
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(G - G_ref) / scale)) if G.size else 0.0


def lincomb_rel_error(R, R_ref, A, C):
    """max |dR_dj| / (||A[d, :]|| ||c_j||)."""
    if R.size == 0:
        return 0.0
    scale = np.outer(np.linalg.norm(A, axis=1), np.linalg.norm(np.atleast_2d(C), axis=0))
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(R - R_ref) / scale))
