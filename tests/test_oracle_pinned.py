"""Pin the CPU oracle (oracle/numpy_path.py) to the reference.

1. against the golden fixtures produced by running pyMOR itself (tests/golden/generate.py);
2. when /root/reference is importable (build container), against the live reference on fresh inputs.
"""
import sys

import numpy as np
import pytest
import scipy.sparse as sps

from conftest import HAVE_REFERENCE, REFERENCE_SRC, csr_from, load_golden
from oracle import numpy_path as O

TIGHT = dict(rtol=1e-13, atol=1e-13)


def test_vecops_golden():
    g = load_golden('vecops')
    A, B, w, C, alpha = g['A'], g['B'], g['w'], g['C'], g['alpha']
    M = csr_from(g, 'M')
    lb = B.shape[1]
    np.testing.assert_allclose(O.inner(A, B), g['inner'], **TIGHT)
    np.testing.assert_allclose(O.inner(A, B, w), g['inner_diag'], **TIGHT)
    np.testing.assert_allclose(O.inner(A, B, M), g['inner_csr'], **TIGHT)
    np.testing.assert_allclose(O.gramian(A), g['gramian'], **TIGHT)
    np.testing.assert_allclose(O.gramian(A, w), g['gramian_diag'], **TIGHT)
    np.testing.assert_allclose(O.gramian(A, M), g['gramian_csr'], **TIGHT)
    np.testing.assert_allclose(O.pairwise_inner(A[:, :lb], B), g['pairwise'], **TIGHT)
    np.testing.assert_allclose(O.pairwise_inner(A[:, :lb], B, w), g['pairwise_diag'], **TIGHT)
    np.testing.assert_allclose(O.pairwise_inner(A[:, :lb], B, M), g['pairwise_csr'], **TIGHT)
    np.testing.assert_allclose(O.lincomb(A, C), g['lincomb'], **TIGHT)
    np.testing.assert_allclose(O.norm(A), g['norm'], **TIGHT)
    np.testing.assert_allclose(O.norm(A, w), g['norm_diag'], **TIGHT)
    np.testing.assert_allclose(O.norm2(A), g['norm2'], **TIGHT)
    np.testing.assert_allclose(O.norm2(A, M), g['norm2_csr'], **TIGHT)
    np.testing.assert_allclose(O.axpy(A.copy(), 0.37, A[:, ::-1]), g['axpy_scalar_rev'], **TIGHT)
    np.testing.assert_allclose(O.axpy(A.copy(), alpha, B[:, :1]), g['axpy_vec_bcast'], **TIGHT)
    Y = A.copy()
    Y[:, [5, 1, 3]] = O.axpy(Y[:, [5, 1, 3]], alpha[:3], B[:, [0, 4, 2]])
    np.testing.assert_allclose(Y, g['axpy_list'], **TIGHT)
    np.testing.assert_allclose(O.scal(A.copy(), alpha), g['scal_vec'], **TIGHT)
    Y = A.copy()
    Y[:, 1:6:2] = O.scal(Y[:, 1:6:2], -2.5)
    np.testing.assert_allclose(Y, g['scal_slice'], **TIGHT)
    np.testing.assert_array_equal(O.dofs(A[:, [6, 0, 2]], g['dof_idx']), g['dofs'])
    idx, val = O.amax(A)
    np.testing.assert_array_equal(idx, g['amax_idx'])
    np.testing.assert_array_equal(val, g['amax_val'])
    np.testing.assert_allclose(O.inner(A[:, [6, 0, 0, 2]], B[:, 3:0:-1]), g['view_inner'], **TIGHT)
    np.testing.assert_allclose(O.apply_product(M, A), g['apply_csr'], **TIGHT)


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_gram_schmidt_golden(name):
    g = load_golden('gram_schmidt')
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    stats = {}
    Q, R = O.gram_schmidt(g['A'], W, return_R=True, stats=stats)
    assert Q.shape == g[f'Q_{name}'].shape           # same vectors removed
    assert sorted(stats['removed']) == [7, 9]
    assert stats['reiterations'] >= 1
    np.testing.assert_allclose(Q, g[f'Q_{name}'], rtol=0, atol=1e-11)
    np.testing.assert_allclose(R, g[f'R_{name}'], rtol=1e-11, atol=1e-11)


def test_gram_schmidt_offset_golden():
    g = load_golden('gram_schmidt')
    A = g['A']
    Q0 = O.gram_schmidt(A[:, :4])
    Q1 = O.gram_schmidt(np.hstack([Q0, A[:, 10:]]), offset=4)
    np.testing.assert_allclose(Q1, g['Q_offset'], rtol=0, atol=1e-11)


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_pod_golden(name):
    g = load_golden('pod')
    A = g['A']
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    U, s, Vh = O.method_of_snapshots(A, W)
    np.testing.assert_allclose(s, g[f'mos_s_{name}'], rtol=1e-10)
    assert U.shape == g[f'mos_U_{name}'].shape
    assert O.subspace_angle(g[f'mos_U_{name}'], U, W) < 1e-8
    # columns agree up to sign (singular values are well separated)
    signs = np.sign(np.sum(O.apply_product(W, U) * g[f'mos_U_{name}'], axis=0))
    np.testing.assert_allclose(U * signs, g[f'mos_U_{name}'], atol=1e-9)
    U, s, _ = O.qr_svd(A, W)
    np.testing.assert_allclose(s, g[f'qr_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'qr_U_{name}'], U, W) < 1e-8
    U, s = O.pod(A, W, modes=10)
    np.testing.assert_allclose(s, g[f'pod_s_{name}'], rtol=1e-10)
    assert U.shape[1] == 10 and O.subspace_angle(g[f'pod_U_{name}'], U, W) < 1e-8


def test_pod_illconditioned_takes_reorthogonalisation_branch():
    g = load_golden('pod')
    stats = {}
    U, s = O.pod(g['A_illcond'], rtol=1e-7, stats=stats)
    assert stats['reorthogonalized']
    assert U.shape == g['pod_U_illcond'].shape
    np.testing.assert_allclose(s, g['pod_s_illcond'], rtol=1e-6)
    assert O.orthogonality_error(U) < 1e-10


def test_projection_golden():
    g = load_golden('projection')
    M, P = csr_from(g, 'Mop'), csr_from(g, 'P')
    V, U = g['V'], g['U']
    np.testing.assert_allclose(O.project(M, V, U), g['proj_both'], **TIGHT)
    np.testing.assert_allclose(O.project(M, V, U, P), g['proj_both_product'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.project(M, U, U), g['proj_galerkin'], **TIGHT)
    np.testing.assert_allclose(O.project(M, V, None) @ U, g['proj_range_only'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.project(M, None, U), g['proj_source_only'], **TIGHT)
    np.testing.assert_allclose(O.apply2(M, V, U), g['apply2'], **TIGHT)
    np.testing.assert_allclose(O.pairwise_inner(V, U[:, :V.shape[1]], M), g['pairwise_apply2'], **TIGHT)


def test_thermalblock_golden():
    """Config 1 in miniature: POD of 81 thermal-block snapshots w.r.t. the H1_0-semi product."""
    g = load_golden('thermalblock')
    S, P = g['snapshots'], csr_from(g, 'P')
    U, s = O.pod(S, P, modes=5, method='qr_svd')
    np.testing.assert_allclose(s, g['pod_qr_s'], rtol=1e-10)
    assert O.subspace_angle(g['pod_qr_U'], U, P) < 1e-8
    U2, s2 = O.pod(S, P, modes=5, method='method_of_snapshots')
    np.testing.assert_allclose(s2, g['pod_mos_s'], rtol=1e-10)
    assert O.subspace_angle(g['pod_mos_U'], U2, P) < 1e-8
    B = g['pod_qr_U']
    for i in range(int(g['n_ops'])):
        n = len(g[f'op{i}_indptr']) - 1
        Mi = sps.csr_matrix((g[f'op{i}_data'], g[f'op{i}_indices'], g[f'op{i}_indptr']), shape=(n, n))
        np.testing.assert_allclose(O.project(Mi, B, B), g[f'op{i}_proj'], rtol=1e-12, atol=1e-12)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REFERENCE, reason='reference not present')
def test_oracle_against_live_reference(rng):
    sys.path.insert(0, str(REFERENCE_SRC))
    from pymor.algorithms.gram_schmidt import gram_schmidt
    from pymor.algorithms.pod import pod
    from pymor.core.logger import set_log_levels
    from pymor.operators.numpy import NumpyMatrixOperator
    from pymor.vectorarrays.numpy import NumpyVectorSpace
    set_log_levels({'pymor': 'ERROR'})
    for dim, k in [(1, 1), (50, 3), (400, 17)]:
        A = np.asfortranarray(rng.standard_normal((dim, k)))
        B = np.asfortranarray(rng.standard_normal((dim, k)))
        w = rng.uniform(0.5, 1.5, dim)
        sp = NumpyVectorSpace(dim)
        VA, VB = sp.from_numpy(A.copy()), sp.from_numpy(B.copy())
        Wd = NumpyMatrixOperator(sps.diags(w).tocsr())
        np.testing.assert_allclose(O.inner(A, B, w), VA.inner(VB, Wd), **TIGHT)
        np.testing.assert_allclose(O.pairwise_inner(A, B, w), VA.pairwise_inner(VB, Wd), **TIGHT)
        np.testing.assert_allclose(O.norm(A), VA.norm(), **TIGHT)
        Q, R = gram_schmidt(VA, product=Wd, return_R=True)
        Qo, Ro = O.gram_schmidt(A, w, return_R=True)
        np.testing.assert_allclose(Qo, Q.to_numpy(), atol=1e-11)
        np.testing.assert_allclose(Ro, R, atol=1e-11)
        U, s = pod(VA, product=Wd, method='method_of_snapshots')
        Uo, so = O.pod(A, w)
        np.testing.assert_allclose(so, s, rtol=1e-10)
        assert O.subspace_angle(U.to_numpy(), Uo, w) < 1e-8


def test_shifted_chol_qr_golden():
    """"Next" row N1.  R is pinned to 1e-13; Q of the ill-conditioned input (cond 1e9) only to
    cond * eps, its orthogonality and A = Q R to rounding."""
    g = load_golden('chol_qr')
    for tag, tolQ in (('mild', 1e-12), ('ill', 1e-6)):
        A, w = g[f'A_{tag}'], g[f'w_{tag}']
        rs = tag == 'ill'
        Q, R = O.shifted_chol_qr(A, return_R=True, recompute_shift=rs)
        np.testing.assert_allclose(R, g[f'R_euclid_{tag}'], atol=1e-13)
        np.testing.assert_allclose(Q, g[f'Q_euclid_{tag}'], atol=tolQ)
        assert O.orthogonality_error(Q) < 1e-14
        Q, R = O.shifted_chol_qr(A, w, return_R=True, product_norm=float(np.sqrt(w.max())), recompute_shift=rs)
        np.testing.assert_allclose(R, g[f'R_diag_{tag}'], atol=1e-13)
        np.testing.assert_allclose(Q, g[f'Q_diag_{tag}'], atol=tolQ)
        np.testing.assert_allclose(Q @ R, A, atol=1e-13)
    Q0 = O.shifted_chol_qr(g['A_mild'][:, :5])
    Q1, R1 = O.shifted_chol_qr(np.hstack([Q0, g['A_mild'][:, 5:]]), offset=5, return_R=True)
    np.testing.assert_allclose(Q1, g['Q_offset'], atol=1e-12)
    np.testing.assert_allclose(R1, g['R_offset'], atol=1e-13)


def test_empirical_interpolation_golden():
    """"Next" row N5: EI-greedy and DEIM restatements against the reference's results."""
    g = load_golden('ei')
    A = g['A']
    for tag, kw in (('l2', dict(rtol=1e-6)), ('sup', dict(error_norm='sup', rtol=1e-6)),
                    ('max8', dict(max_interpolation_dofs=8)), ('nodal', dict(rtol=1e-4, nodal_basis=True))):
        dofs, basis, data = O.ei_greedy(A, **kw)
        np.testing.assert_array_equal(dofs, g[f'ei_{tag}_dofs'])
        np.testing.assert_allclose(basis, g[f'ei_{tag}_basis'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(data['errors'], g[f'ei_{tag}_errors'], rtol=1e-9)
        np.testing.assert_allclose(data['coefficients'], g[f'ei_{tag}_coefficients'], rtol=0, atol=1e-7)
        np.testing.assert_allclose(data['interpolation_matrix'], g[f'ei_{tag}_matrix'], rtol=0, atol=1e-9)
    for tag, w in (('euclid', None), ('diag', g['w'])):
        dofs, basis, data = O.deim(A, w, modes=9)
        np.testing.assert_array_equal(dofs, g[f'deim_{tag}_dofs'])
        np.testing.assert_allclose(data['svals'], g[f'deim_{tag}_svals'], rtol=1e-10)
        signs = np.sign(np.sum(basis * g[f'deim_{tag}_basis'], axis=0))
        np.testing.assert_allclose(basis * signs, g[f'deim_{tag}_basis'], rtol=0, atol=1e-8)
    dofs, basis, _ = O.deim(A[:, :6], use_pod=False)
    np.testing.assert_array_equal(dofs, g['deim_nopod_dofs'])
    np.testing.assert_array_equal(basis, g['deim_nopod_basis'])
