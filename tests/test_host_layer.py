"""Host-side logic of the CUDA backend on CPU (no GPU needed).

The C ABI is served by tests/emulated_lib.py (NumPy) -- see that module's docstring.  What is under
test here is everything ABOVE the boundary: index/view normalisation, copy-on-write, aliasing,
capacity growth, in-place compaction on delete, operator fusion dispatch, the stand-alone algorithm
drivers, and -- when /root/reference is present -- the reference's own contract suite and pyMOR's own
algorithms running on CudaVectorArray.
"""
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sps

from conftest import HAVE_REFERENCE, ROOT, assert_gram_schmidt_golden, csr_from, load_golden
from emulated_lib import EmulatedLib
from oracle import numpy_path as O

from pymor_b200 import _lib, algorithms as alg


@pytest.fixture
def emu():
    lib = EmulatedLib()
    _lib.set_library_for_testing(lib)
    yield lib
    _lib.set_library_for_testing(None)


def space_of(dim):
    from pymor_b200.vectorarray import CudaVectorSpace
    return CudaVectorSpace(dim, device='cpu')


def test_no_cpu_fallback_without_library_or_gpu():
    """Without an injected test library the product must fail loudly on a box without a B200."""
    _lib.set_library_for_testing(None)
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pymor_b200.vectorarray import CudaVectorSpace
    with pytest.raises(Exception):
        CudaVectorSpace(10)


INDS = [None, slice(1, 6, 2), slice(5, 0, -2), [4, 0, 2], [3, 3, 1], [2], [], slice(0, 0)]


def np_ind(A, ind):
    return A if ind is None else A[:, ind]


def va_ind(V, ind):
    return V if ind is None else V[ind]


@pytest.mark.parametrize('dim', [0, 1, 7, 37])
def test_reductions_and_views(emu, rng, dim):
    sp = space_of(dim)
    A, B = rng.standard_normal((dim, 7)), rng.standard_normal((dim, 7))
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    np.testing.assert_array_equal(VA.to_numpy(), A)
    for i1 in INDS:
        a, va = np_ind(A, i1), va_ind(VA, i1)
        np.testing.assert_allclose(va.norm(), O.norm(a), atol=1e-14)
        np.testing.assert_allclose(va.norm2(), O.norm2(a), atol=1e-14)
        g = va.gramian()
        np.testing.assert_allclose(g, O.gramian(a), atol=1e-13)
        assert g.flags.f_contiguous or g.size <= 1
        np.testing.assert_array_equal(g, g.T)
        for i2 in INDS:
            b, vb = np_ind(B, i2), va_ind(VB, i2)
            np.testing.assert_allclose(va.inner(vb), O.inner(a, b), atol=1e-13)
            np.testing.assert_allclose(va.inner(va_ind(VA, i2)), O.inner(a, np_ind(A, i2)), atol=1e-13)
            if a.shape[1] == b.shape[1]:
                np.testing.assert_allclose(va.pairwise_inner(vb), O.pairwise_inner(a, b), atol=1e-13)
        if dim:
            idx, val = va.amax()
            oi, ov = O.amax(a)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(val, ov)
            dofs = [0, dim - 1, dim // 2, 0]
            np.testing.assert_array_equal(va.dofs(dofs), O.dofs(a, dofs))


def test_lincomb_and_coefficient_layouts(emu, rng):
    sp = space_of(23)
    A = rng.standard_normal((23, 6))
    VA = sp.from_numpy(A)
    C = rng.standard_normal((6, 4))
    for coeffs in (C, np.asfortranarray(C), C[:, ::-1], C.astype(np.float32).astype(np.float64), C[:, :0],
                   np.arange(6)):
        np.testing.assert_allclose(VA.lincomb(coeffs).to_numpy(), O.lincomb(A, coeffs), atol=1e-13)
    np.testing.assert_allclose(VA[[5, 1, 1]].lincomb(C[:3]).to_numpy(), O.lincomb(A[:, [5, 1, 1]], C[:3]), atol=1e-13)
    np.testing.assert_allclose(VA.lincomb(C * (1 + 2j)).to_numpy(), A @ (C * (1 + 2j)), atol=1e-13)


def test_axpy_scal_aliasing_and_index_forms(emu, rng):
    sp = space_of(19)
    A, B = rng.standard_normal((19, 7)), rng.standard_normal((19, 7))
    al = rng.standard_normal(7)
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    # scalar / per-column alpha, broadcast x
    Y = VA.copy(); Y.axpy(0.5, VB); np.testing.assert_allclose(Y.to_numpy(), A + 0.5 * B, atol=1e-14)
    Y = VA.copy(); Y.axpy(al, VB[2]); np.testing.assert_allclose(Y.to_numpy(), A + B[:, [2]] * al, atol=1e-14)
    Y = VA.copy(); Y.axpy(np.arange(7), VB); np.testing.assert_allclose(Y.to_numpy(), A + B * np.arange(7), atol=1e-14)
    # x is a permuted view of self: NumPy semantics = right-hand side evaluated first
    Y = VA.copy(); Y.axpy(2.0, Y[::-1]); np.testing.assert_allclose(Y.to_numpy(), A + 2 * A[:, ::-1], atol=1e-14)
    Y = VA.copy(); Y[[0, 3, 5]].axpy(al[:3], Y[[3, 5, 6]])
    ref = A.copy(); ref[:, [0, 3, 5]] += A[:, [3, 5, 6]] * al[:3]
    np.testing.assert_allclose(Y.to_numpy(), ref, atol=1e-14)
    # disjoint views of the same array do not snapshot (the Gram-Schmidt inner step)
    Y = VA.copy(); Y.scal(1.0)
    before = emu.calls.count('pmc_copy')
    Y[4].axpy(-0.3, Y[1])
    assert emu.calls.count('pmc_copy') == before
    ref = A.copy(); ref[:, 4] -= 0.3 * A[:, 1]
    np.testing.assert_allclose(Y.to_numpy(), ref, atol=1e-14)
    # scal: slices with negative step, irregular lists, int dtype
    Y = VA.copy(); Y[5:0:-2].scal(np.array([2, 3, 4])); ref = A.copy(); ref[:, 5:0:-2] *= [2, 3, 4]
    np.testing.assert_allclose(Y.to_numpy(), ref, atol=1e-14)
    Y = VA.copy(); Y[[6, 0, 1, 4]].scal(al[:4]); ref = A.copy(); ref[:, [6, 0, 1, 4]] *= al[:4]
    np.testing.assert_allclose(Y.to_numpy(), ref, atol=1e-14)
    Y = VA.copy(); Y.scal(1j)                       # in-place promotion to complex (numpy.py:90-95)
    np.testing.assert_allclose(Y.to_numpy(), 1j * A, atol=1e-15)
    np.testing.assert_array_equal(VA.to_numpy(), A)
    with pytest.raises(Exception):
        VA[[1, 1]].scal(2.0)          # repeated indices are not writable


def test_cow_append_delete_growth(emu, rng):
    sp = space_of(11)
    A, B = rng.standard_normal((11, 5)), rng.standard_normal((11, 4))
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    C = VA.copy()
    assert C.impl is VA.impl                       # lazy copy
    C.scal(2.0)
    assert C.impl is not VA.impl                   # copy on first write
    np.testing.assert_array_equal(VA.to_numpy(), A)
    np.testing.assert_allclose(C.to_numpy(), 2 * A)
    V = sp.empty(reserve=2)
    for _ in range(3):                             # growth beyond the reserve
        V.append(VB)
    V.append(V[[1, 0]])                            # append a view of itself
    ref = np.hstack([B, B, B, B[:, [1, 0]]])
    np.testing.assert_array_equal(V.to_numpy(), ref)
    del V[[0, 2, 3, 9, -1]]
    keep = [i for i in range(ref.shape[1]) if i not in (0, 2, 3, 9, ref.shape[1] - 1)]
    np.testing.assert_array_equal(V.to_numpy(), ref[:, keep])
    del V[1:4]
    np.testing.assert_array_equal(V.to_numpy(), np.delete(ref[:, keep], slice(1, 4), axis=1))
    W = sp.from_numpy(B)
    X = sp.empty()
    X.append(W[1:3], remove_from_other=True)
    assert len(W) == 2 and len(X) == 2
    np.testing.assert_array_equal(W.to_numpy(), B[:, [0, 3]])
    np.testing.assert_array_equal(X.to_numpy(), B[:, 1:3])
    Z = sp.from_numpy(A * (1 + 1j))
    np.testing.assert_allclose(Z.to_numpy(), A * (1 + 1j), atol=1e-15)


@pytest.mark.parametrize('kind', ['diag', 'csr'])
def test_product_operators_fused_paths(emu, rng, kind):
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    dim = 41
    sp = space_of(dim)
    A, B = rng.standard_normal((dim, 6)), rng.standard_normal((dim, 4))
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    if kind == 'diag':
        w = rng.uniform(0.5, 1.5, dim)
        W, op = w, CudaDiagonalOperator.from_scipy(sps.diags(w).tocsr(), sp)
    else:
        M = (sps.random(dim, dim, density=0.1, random_state=1) + sps.eye(dim)).tocsr()
        W, op = M, CudaCsrOperator(M, sp)
    emu.calls.clear()
    np.testing.assert_allclose(op.apply2(VA, VB), O.inner(A, B, W), atol=1e-13)
    np.testing.assert_allclose(VA.gramian(op), O.gramian(A, W), atol=1e-13)
    np.testing.assert_allclose(VA[1:5].inner(VA, op), O.inner(A[:, 1:5], A, W), atol=1e-13)
    np.testing.assert_allclose(VA[:4].pairwise_inner(VB, op), O.pairwise_inner(A[:, :4], B, W), atol=1e-13)
    np.testing.assert_allclose(VA.norm(op), O.norm(A, W), atol=1e-13)
    np.testing.assert_allclose(op.apply(VB).to_numpy(), O.apply_product(W, B), atol=1e-14)
    WT = W if kind == 'diag' else W.T
    np.testing.assert_allclose(op.apply_adjoint(VB).to_numpy(), O.apply_product(WT, B), atol=1e-14)
    if kind == 'diag':       # fused: only the two explicit apply() calls ever produce W·U
        assert emu.calls.count('pmc_diag_apply') == 2
    else:
        assert emu.calls.count('pmc_csr_gram') == 3
        assert emu.calls.count('pmc_csr_pairwise') == 2 and emu.calls.count('pmc_csr_apply') == 2   # no M U temporary


def test_csr_operator_symmetric_lower_only_and_local_rows(emu, rng):
    """V^T (M V) with symmetric M is contracted on the lower tile triangle only (PMC_LOWER_ONLY) and mirrored;
    a non-symmetric M never is.  `from_local_rows` (this rank's row block, global column numbers) equals the
    operator built from the global matrix; symmetric operators serve apply_adjoint with apply."""
    from pymor_b200.operators import CudaCsrOperator
    dim, k = 300, 140                               # k > 128: more than one tile row
    h = 1.0 / (dim + 1)
    P = sps.diags([np.full(dim - 1, h / 6), np.full(dim, 4 * h / 6), np.full(dim - 1, h / 6)], [-1, 0, 1]).tocsr()
    N = (P + sps.random(dim, dim, density=0.02, random_state=5)).tocsr()
    sp = space_of(dim)
    A = rng.standard_normal((dim, k))
    VA = sp.from_numpy(A)
    seen = []
    orig = emu.pmc_csr_gram
    emu.pmc_csr_gram = lambda *a: (seen.append(a[-1]), orig(*a))[1]
    try:
        for M, want in ((P, 8), (N, 0)):
            op = CudaCsrOperator(M, sp)
            G = op.apply2(VA, VA)
            np.testing.assert_allclose(G, A.T @ (M @ A), atol=1e-13)
            assert seen[-1] & 8 == want
            op.apply2(VA[:100], VA)                 # different panels: never lower-only
            assert seen[-1] & 8 == 0
    finally:
        emu.pmc_csr_gram = orig
    loc = CudaCsrOperator.from_local_rows(P[sp.offset:sp.offset + sp.local_dim], sp, symmetric=True)
    np.testing.assert_allclose(loc.apply(VA[:5]).to_numpy(), P @ A[:, :5], atol=1e-14)
    np.testing.assert_allclose(loc.apply_adjoint(VA[:5]).to_numpy(), P.T @ A[:, :5], atol=1e-14)
    np.testing.assert_allclose(loc.apply2(VA, VA), A.T @ (P @ A), atol=1e-13)
    np.testing.assert_allclose(VA[:9].norm(loc), O.norm(A[:, :9], P), atol=1e-13)
    X = loc.apply_inverse(loc.apply(VA[:3]))        # symmetric => device CG with the Jacobi preconditioner
    np.testing.assert_allclose(X.to_numpy(), A[:, :3], atol=1e-7)
    from pymor_b200.solvers import CgSolver
    loc2 = loc.with_(solver=CgSolver(tol=1e-12), name='again')      # ImmutableObject.with_ re-runs __init__
    assert loc2.name == 'again' and loc2.is_symmetric()
    np.testing.assert_allclose(loc2.apply(VA[:2]).to_numpy(), P @ A[:, :2], atol=1e-14)
    nonsym = CudaCsrOperator.from_local_rows(N[sp.offset:sp.offset + sp.local_dim], sp)
    with pytest.raises(NotImplementedError):
        nonsym.apply_adjoint(VA[:2])


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_standalone_gram_schmidt_golden(emu, name):
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp),
            'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    VA = sp.from_numpy(g['A'])
    Q, R = alg.gram_schmidt(VA, product=prod, return_R=True)
    np.testing.assert_array_equal(VA.to_numpy(), g['A'])        # copy=True leaves the input alone
    if name == 'csr':
        # the fused pairwise two-form (pmc_csr_pairwise) sums v_r (M u)_r in one pass: rounding differs from the
        # reference's apply-then-dot, which the near-dependent column of this input amplifies by 1e9
        from conftest import assert_gram_schmidt_golden
        assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, csr_from(g, 'M'))
        assert 'pmc_csr_pairwise' in emu.calls
    else:
        np.testing.assert_allclose(Q.to_numpy(), g[f'Q_{name}'], atol=1e-11)
        np.testing.assert_allclose(R, g[f'R_{name}'], atol=1e-11)
    Q0 = alg.gram_schmidt(sp.from_numpy(g['A'][:, :4]))
    Q0.append(sp.from_numpy(g['A'][:, 10:]))
    if name == 'euclid':
        Q1 = alg.gram_schmidt(Q0, offset=4, copy=False)
        assert Q1 is Q0
        np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-11)


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_standalone_pod_golden(emu, name):
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('pod')
    sp = space_of(g['A'].shape[0])
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp),
            'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    VA = sp.from_numpy(g['A'])
    U, s, Vh = alg.method_of_snapshots(VA, product=prod)
    np.testing.assert_allclose(s, g[f'mos_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'mos_U_{name}'], U.to_numpy(), W) < 1e-8
    U, s, _ = alg.qr_svd(VA, product=prod)
    np.testing.assert_allclose(s, g[f'qr_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'qr_U_{name}'], U.to_numpy(), W) < 1e-8
    U, s = alg.pod(VA, product=prod, modes=10, method='method_of_snapshots')
    np.testing.assert_allclose(s, g[f'pod_s_{name}'], rtol=1e-10)
    assert len(U) == 10 and O.subspace_angle(g[f'pod_U_{name}'], U.to_numpy(), W) < 1e-8
    if name == 'euclid':
        U, s = alg.pod(sp.from_numpy(g['A_illcond']), rtol=1e-7, method='method_of_snapshots')
        # method_of_snapshots squares the condition number: eigenvalues carry an absolute error of
        # eps * lambda_max, so the small singular values are only comparable through s**2
        np.testing.assert_allclose(s ** 2, g['pod_s_illcond'] ** 2, rtol=1e-9, atol=1e-13)
        assert O.orthogonality_error(U.to_numpy()) < 1e-10


def test_standalone_project_golden(emu):
    from pymor_b200.operators import CudaCsrOperator
    g = load_golden('projection')
    sp = space_of(g['V'].shape[0])
    op, prod = CudaCsrOperator(csr_from(g, 'Mop'), sp), CudaCsrOperator(csr_from(g, 'P'), sp)
    V, U = sp.from_numpy(g['V']), sp.from_numpy(g['U'])
    np.testing.assert_allclose(alg.project(op, V, U).matrix, g['proj_both'], atol=1e-12)
    np.testing.assert_allclose(alg.project(op, V, U, product=prod).matrix, g['proj_both_product'], atol=1e-12)
    np.testing.assert_allclose(alg.project(op, U, U).matrix, g['proj_galerkin'], atol=1e-12)
    np.testing.assert_allclose(alg.project(op, None, U).array.to_numpy(), g['proj_source_only'], atol=1e-12)
    np.testing.assert_allclose(alg.project(op, V, None).array.inner(U), g['proj_range_only'], atol=1e-12)
    np.testing.assert_allclose(op.pairwise_apply2(V, U[:len(V)]), g['pairwise_apply2'], atol=1e-12)


def test_thermalblock_golden(emu):
    """Config 1 in miniature through the stand-alone drivers."""
    from pymor_b200.operators import CudaCsrOperator
    g = load_golden('thermalblock')
    S, P = g['snapshots'], csr_from(g, 'P')
    sp = space_of(S.shape[0])
    prod = CudaCsrOperator(P, sp)
    snaps = sp.empty(reserve=S.shape[1])
    for i in range(S.shape[1]):                      # appended one solve at a time like reduce_pod
        snaps.append(sp.from_numpy(S[:, [i]]))
    basis, svals = alg.pod(snaps, modes=5, product=prod)
    np.testing.assert_allclose(svals, g['pod_qr_s'], rtol=1e-10)
    assert O.subspace_angle(g['pod_qr_U'], basis.to_numpy(), P) < 1e-8
    basis2, svals2 = alg.pod(snaps, modes=5, product=prod, method='method_of_snapshots')
    np.testing.assert_allclose(svals2, g['pod_mos_s'], rtol=1e-10)
    B = sp.from_numpy(g['pod_qr_U'])
    for i in range(int(g['n_ops'])):
        n = len(g[f'op{i}_indptr']) - 1
        Mi = sps.csr_matrix((g[f'op{i}_data'], g[f'op{i}_indices'], g[f'op{i}_indptr']), shape=(n, n))
        np.testing.assert_allclose(alg.project(CudaCsrOperator(Mi, sp), B, B).matrix, g[f'op{i}_proj'], atol=1e-12)


def test_dlpack_roundtrip(emu, rng):
    import torch
    sp = space_of(32)
    t = torch.from_numpy(rng.standard_normal((5, 32)))
    V = sp.make_array(t)
    assert V.impl._buf.data_ptr() == t.data_ptr()          # zero-copy adoption
    np.testing.assert_array_equal(V.to_numpy(), t.numpy().T)
    back = torch.from_dlpack(V)
    assert back.data_ptr() == t.data_ptr() and back.shape == (5, 32)
    odd = torch.from_numpy(rng.standard_normal((3, 33)))[:, :32]   # odd row stride: copied
    W = sp.make_array(odd)
    assert W.impl._buf.data_ptr() != odd.data_ptr()
    np.testing.assert_array_equal(W.to_numpy(), odd.numpy().T)


# --------------------------------------------------------------------------------------------
# with the reference present: its own test-suite and its own algorithms on CudaVectorArray
# --------------------------------------------------------------------------------------------
def _run(script, *args):
    return subprocess.run([sys.executable, str(ROOT / 'tests' / 'reference_contract' / script), *args],
                          capture_output=True, text=True, timeout=3000)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REFERENCE, reason='reference not present')
def test_reference_contract_suite_on_cuda_backend():
    """pymortests/vectorarray.py + algorithms/{gram_schmidt,svd_va,pod,chol_qr,basic}.py with backend
    'cuda', unrestricted: float64 and complex128 inputs, nothing deselected."""
    r = _run('run_contract.py', 'src/pymortests/vectorarray.py', 'src/pymortests/algorithms/gram_schmidt.py',
             'src/pymortests/algorithms/svd_va.py', 'src/pymortests/algorithms/pod.py',
             'src/pymortests/algorithms/chol_qr.py', 'src/pymortests/algorithms/basic.py')
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    import re
    assert ' passed' in tail and not re.search(r'\d+ (failed|error)', tail)     # ('xfailed' is fine)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REFERENCE, reason='reference not present')
def test_reference_operator_contract_on_cuda_operators():
    """pymortests/operators/operators.py (apply, apply2, pairwise_apply2, apply_adjoint, H, assemble, jacobian,
    as_range_array, restricted, inverse wrappers, arithmetic) and pymortests/algorithms/projection.py (project with
    and without products, to sub-bases) with CudaDiagonalOperator / CudaCsrOperator / a LincombOperator of them
    registered as the fixture generators -- the Operator half of the plug-in boundary."""
    import re
    r = _run('run_operator_contract.py')
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    assert ' passed' in tail and not re.search(r'\d+ (failed|error)', tail)


@pytest.mark.reference
@pytest.mark.skipif(not HAVE_REFERENCE, reason='reference not present')
def test_pymor_algorithms_run_unchanged():
    r = _run('run_contract.py', str(ROOT / 'tests' / 'reference_contract' / 'pymor_mode_checks.py'))
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail


def test_standalone_shifted_chol_qr_golden(emu):
    """"Next" row N1 through the stand-alone driver (Gramian / lincomb / append / delete only)."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('chol_qr')
    for tag, tolQ in (('mild', 1e-12), ('ill', 1e-6)):
        A, w = g[f'A_{tag}'], g[f'w_{tag}']
        sp = space_of(A.shape[0])
        rs = tag == 'ill'
        Q, R = alg.shifted_chol_qr(sp.from_numpy(A), return_R=True, recompute_shift=rs)
        np.testing.assert_allclose(R, g[f'R_euclid_{tag}'], atol=1e-12)
        np.testing.assert_allclose(Q.to_numpy(), g[f'Q_euclid_{tag}'], atol=tolQ)
        prod = CudaDiagonalOperator(w, sp)
        Q, R = alg.shifted_chol_qr(sp.from_numpy(A), product=prod, return_R=True,
                                   product_norm=float(np.sqrt(w.max())), recompute_shift=rs)
        np.testing.assert_allclose(R, g[f'R_diag_{tag}'], atol=1e-12)
        np.testing.assert_allclose(Q.to_numpy(), g[f'Q_diag_{tag}'], atol=tolQ)
        assert O.orthogonality_error(Q.to_numpy(), w) < 1e-13
    sp = space_of(g['A_mild'].shape[0])
    Q0 = alg.shifted_chol_qr(sp.from_numpy(g['A_mild'][:, :5]))
    Q0.append(sp.from_numpy(g['A_mild'][:, 5:]))
    Q1, R1 = alg.shifted_chol_qr(Q0, offset=5, return_R=True)
    np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-12)
    np.testing.assert_allclose(R1, g['R_offset'], atol=1e-12)


def test_deferred_axpy_fusion_semantics(emu, rng):
    """`v.axpy(a, q)` on single vectors is deferred and merged with the next dot/norm on v
    (pmc_axpy_dot); any other access must see the updated v, and q must be read before it changes."""
    from pymor_b200.operators import CudaDiagonalOperator
    from pymor_b200.vectorarray import CudaVectorSpace
    dim = 33
    sp = space_of(dim)
    A = rng.standard_normal((dim, 5))
    w = rng.uniform(0.5, 1.5, dim)
    prod = CudaDiagonalOperator(w, sp)

    def fresh():
        return sp.from_numpy(A)

    # fused with the next dot (Gram-Schmidt order: q.pairwise_inner(v)) and with the norm
    V = fresh(); emu.calls.clear()
    V[2].axpy(-0.5, V[0])
    p = V[1].pairwise_inner(V[2], prod)[0]
    assert emu.calls == ['pmc_axpy_dot']
    ref = A[:, 2] - 0.5 * A[:, 0]
    np.testing.assert_allclose(p, np.sum(A[:, 1] * w * ref), rtol=1e-14)
    V[2].axpy(0.25, V[1])
    nrm = V[2].norm(prod)[0]
    assert emu.calls == ['pmc_axpy_dot', 'pmc_axpy_dot']
    ref = ref + 0.25 * A[:, 1]
    np.testing.assert_allclose(nrm, np.sqrt(np.sum(ref * w * ref)), rtol=1e-14)
    np.testing.assert_allclose(V.to_numpy()[:, 2], ref, atol=1e-15)
    # not fusable orders fall back to axpy + reduction, same numbers
    V = fresh(); emu.calls.clear()
    V[2].axpy(-0.5, V[0])
    p2 = V[2].pairwise_inner(V[1], prod)[0]                 # target is the FIRST operand
    assert emu.calls == ['pmc_axpy', 'pmc_pairwise_inner']
    np.testing.assert_allclose(p2, np.sum((A[:, 2] - 0.5 * A[:, 0]) * w * A[:, 1]), rtol=1e-14)
    # the source vector is modified right after: the deferred axpy must have used the OLD source
    V = fresh()
    V[2].axpy(2.0, V[0])
    V[0].scal(100.0)
    np.testing.assert_allclose(V.to_numpy()[:, 2], A[:, 2] + 2.0 * A[:, 0], atol=1e-14)
    # ... also when the source lives in another array of an equal-but-distinct space object
    sp2 = CudaVectorSpace(dim, device='cpu')
    assert sp2 == sp and sp2 is not sp
    Q = sp2.from_numpy(A[:, [0]])
    V = fresh()
    V[3].axpy(-1.5, Q[0])
    Q.scal(0.0)
    np.testing.assert_allclose(V.to_numpy()[:, 3], A[:, 3] - 1.5 * A[:, 0], atol=1e-14)
    # every other consumer sees the update: copy, lincomb, inner, gramian, dofs, append, two deferrals in a row
    V = fresh(); V[1].axpy(1.0, V[4]); C = V.copy(deep=True)
    np.testing.assert_allclose(C.to_numpy()[:, 1], A[:, 1] + A[:, 4], atol=1e-15)
    V = fresh(); V[1].axpy(1.0, V[4])
    np.testing.assert_allclose(V.lincomb(np.eye(5)[:, [1]]).to_numpy()[:, 0], A[:, 1] + A[:, 4], atol=1e-15)
    V = fresh(); V[1].axpy(1.0, V[4]); V[2].axpy(-1.0, V[1])
    ref = A.copy(); ref[:, 1] += ref[:, 4]; ref[:, 2] -= ref[:, 1]
    np.testing.assert_allclose(V.gramian(prod), O.gramian(ref, w), atol=1e-13)
    np.testing.assert_allclose(V.dofs([0, 7]), O.dofs(ref, [0, 7]), atol=1e-15)
    V = fresh(); V[0].axpy(3.0, V[1]); W2 = sp.empty(); W2.append(V[0])
    np.testing.assert_allclose(W2.to_numpy()[:, 0], A[:, 0] + 3.0 * A[:, 1], atol=1e-15)


def test_complex_arrays(emu, rng):
    """Complex arrays = pairs of real device arrays; every op against NumPy (conjugate-linear inner)."""
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    dim = 17
    sp = space_of(dim)
    A = rng.standard_normal((dim, 5)) + 1j * rng.standard_normal((dim, 5))
    B = rng.standard_normal((dim, 5)) + 1j * rng.standard_normal((dim, 5))
    Rl = rng.standard_normal((dim, 5))
    w = rng.uniform(0.5, 1.5, dim)
    VA, VB, VR = sp.from_numpy(A), sp.from_numpy(B), sp.from_numpy(Rl)
    np.testing.assert_allclose(VA.to_numpy(), A)
    np.testing.assert_allclose(VA.inner(VB), A.conj().T @ B, atol=1e-13)
    np.testing.assert_allclose(VA[1:4].inner(VR[[4, 0]]), A[:, 1:4].conj().T @ Rl[:, [4, 0]], atol=1e-13)
    np.testing.assert_allclose(VR.inner(VA), Rl.T @ A, atol=1e-13)
    np.testing.assert_allclose(VA.pairwise_inner(VB), np.sum(A.conj() * B, axis=0), atol=1e-13)
    np.testing.assert_allclose(VA.norm(), np.linalg.norm(A, axis=0), atol=1e-13)
    g = VA.gramian()
    np.testing.assert_allclose(g, A.conj().T @ A, atol=1e-13)
    np.testing.assert_array_equal(g, g.conj().T)
    C = rng.standard_normal((5, 3)) + 1j * rng.standard_normal((5, 3))
    np.testing.assert_allclose(VA.lincomb(C).to_numpy(), A @ C, atol=1e-13)
    np.testing.assert_allclose(VA.lincomb(C.real).to_numpy(), A @ C.real, atol=1e-13)
    al = rng.standard_normal(5) + 1j * rng.standard_normal(5)
    Y = VA.copy(); Y.scal(al); np.testing.assert_allclose(Y.to_numpy(), A * al, atol=1e-13)
    Y = VA.copy(); Y.axpy(al, VB); np.testing.assert_allclose(Y.to_numpy(), A + B * al, atol=1e-13)
    Y = VR.copy(); Y.axpy(2 - 1j, VA[2]); np.testing.assert_allclose(Y.to_numpy(), Rl + (2 - 1j) * A[:, [2]], atol=1e-13)
    Y = VA.copy(); Y.axpy(0.5j, Y[::-1]); np.testing.assert_allclose(Y.to_numpy(), A + 0.5j * A[:, ::-1], atol=1e-13)
    Y = VR.copy(); Y.append(VA[[3, 3]]); del Y[1]
    np.testing.assert_allclose(Y.to_numpy(), np.delete(np.hstack([Rl, A[:, [3, 3]]]), 1, axis=1), atol=1e-15)
    np.testing.assert_allclose(VA.real.to_numpy(), A.real); np.testing.assert_allclose(VA.imag.to_numpy(), A.imag)
    np.testing.assert_allclose(VA.conj().to_numpy(), A.conj())
    idx, val = VA.amax()
    np.testing.assert_array_equal(idx, np.argmax(np.abs(A), axis=0))
    np.testing.assert_allclose(val, np.max(np.abs(A), axis=0), rtol=1e-14)
    np.testing.assert_allclose(VA.dofs([0, 5]), A[[0, 5]])
    prod = CudaDiagonalOperator(w, sp)
    np.testing.assert_allclose(VA.inner(VB, prod), A.conj().T @ (w[:, None] * B), atol=1e-13)
    np.testing.assert_allclose(VA.norm(prod), np.sqrt(np.sum((A.conj() * (w[:, None] * A)).real, axis=0)), atol=1e-13)
    M = (sps.random(dim, dim, density=0.3, random_state=2) + sps.eye(dim)).tocsr()
    op = CudaCsrOperator(M, sp)
    np.testing.assert_allclose(op.apply(VA).to_numpy(), M @ A, atol=1e-13)
    np.testing.assert_allclose(op.apply2(VA, VB), A.conj().T @ (M @ B), atol=1e-12)
    Q, R = alg.gram_schmidt(VA, return_R=True)
    np.testing.assert_allclose(Q.lincomb(R).to_numpy(), A, atol=1e-12)
    np.testing.assert_allclose(Q.gramian(), np.eye(5), atol=1e-12)


@pytest.mark.parametrize('name', ['euclid', 'diag'])
def test_standalone_inc_hapod_golden(emu, name):
    """"Next" row N3: the stand-alone incremental HAPOD reproduces pyMOR's (mode count, singular
    values, subspace) for 1, 4 and 7 incremental steps."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('hapod')
    A, w = g['A'], g['w']
    sp = space_of(A.shape[0])
    W = None if name == 'euclid' else w
    prod = None if name == 'euclid' else CudaDiagonalOperator(w, sp)
    for steps in (1, 4, 7):
        modes, svals, count = alg.inc_vectorarray_hapod(steps, sp.from_numpy(A), 1e-3, 0.75, product=prod)
        ref_modes, ref_svals = g[f'modes_{name}_{steps}'], g[f'svals_{name}_{steps}']
        assert count == A.shape[1] and len(modes) == ref_modes.shape[1]
        np.testing.assert_allclose(svals, ref_svals, rtol=1e-8)
        assert O.subspace_angle(ref_modes, modes.to_numpy(), W) < 1e-6
        assert O.orthogonality_error(modes.to_numpy(), W) < 1e-10


def test_inc_hapod_over_host_chunks(emu):
    """inc_hapod_chunks: the snapshots arrive chunk by chunk from host memory (a set too large for HBM); identical
    to inc_vectorarray_hapod on the resident array."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('hapod')
    A, w = g['A'], g['w']
    sp = space_of(A.shape[0])
    prod = CudaDiagonalOperator(w, sp)
    steps = 4
    chunk = -(-A.shape[1] // steps)
    blocks = (sp.from_local_numpy(np.asfortranarray(A[:, s:s + chunk])) for s in range(0, A.shape[1], chunk))
    modes, svals, count = alg.inc_hapod_chunks(blocks, steps, 1e-3, 0.75, product=prod)
    ref_modes, ref_svals, ref_count = alg.inc_vectorarray_hapod(steps, sp.from_numpy(A), 1e-3, 0.75, product=prod)
    assert count == ref_count == A.shape[1] and len(modes) == len(ref_modes)
    np.testing.assert_array_equal(svals, ref_svals)
    np.testing.assert_array_equal(modes.to_numpy(), ref_modes.to_numpy())
    np.testing.assert_allclose(svals, g[f'svals_diag_{steps}'], rtol=1e-8)
    with pytest.raises(AssertionError):
        alg.inc_hapod_chunks((sp.from_numpy(A[:, :5]) for _ in range(2)), 3, 1e-3, 0.75)


@pytest.mark.parametrize('name', ['euclid', 'diag'])
def test_standalone_dist_hapod_golden(emu, name):
    """N3, distributed tree: two-level (one POD per slice + root) and a binary tree over 7 slices."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('hapod')
    A, w = g['A'], g['w']
    sp = space_of(A.shape[0])
    W = None if name == 'euclid' else w
    prod = None if name == 'euclid' else CudaDiagonalOperator(w, sp)
    for slices, arity in ((1, None), (5, None), (7, 2)):
        V = sp.from_numpy(A)
        modes, svals, count = alg.dist_vectorarray_hapod(slices, V, 1e-3, 0.75, arity=arity, product=prod)
        ref_modes, ref_svals = g[f'dist_modes_{name}_{slices}_{arity}'], g[f'dist_svals_{name}_{slices}_{arity}']
        assert count == A.shape[1] and len(modes) == ref_modes.shape[1]
        np.testing.assert_allclose(svals, ref_svals, rtol=1e-8)
        assert O.subspace_angle(ref_modes[:, :7], modes.to_numpy()[:, :7], W) < 1e-6      # the 7 signal modes
        assert O.orthogonality_error(modes.to_numpy(), W) < 1e-10
        np.testing.assert_array_equal(V.to_numpy(), A)                                     # input untouched


def test_standalone_randomized_range_finder(emu, rng):
    """"Next" row N2: the range of a rank-5 operator is captured by 8 random samples."""
    from pymor_b200.operators import CudaCsrOperator
    dim = 90
    L, Rr = rng.standard_normal((dim, 5)), rng.standard_normal((5, dim))
    M = sps.csr_matrix(L @ Rr)
    sp = space_of(dim)
    op = CudaCsrOperator(M, sp)
    for qr in ('gram_schmidt', 'shifted_chol_qr'):
        for power in (0, 2):
            Q = alg.RandomizedRangeFinder(op, power_iterations=power, qr_method=qr, error_estimator='loo').find_range(
                basis_size=8 if qr == 'gram_schmidt' else 5)
            Qh = Q.to_numpy()
            assert O.orthogonality_error(Qh) < 1e-10
            Md = M.toarray()
            assert np.linalg.norm(Md - Qh @ (Qh.T @ Md)) < 1e-8 * np.linalg.norm(Md)


def test_standalone_empirical_interpolation_golden(emu):
    """"Next" row N5 (EI-greedy, DEIM) through the stand-alone drivers vs the reference's results."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('ei')
    A = g['A']
    sp = space_of(A.shape[0])
    for tag, kw in (('l2', dict(rtol=1e-6)), ('sup', dict(error_norm='sup', rtol=1e-6)),
                    ('max8', dict(max_interpolation_dofs=8)), ('nodal', dict(rtol=1e-4, nodal_basis=True))):
        V = sp.from_numpy(A)
        dofs, basis, data = alg.ei_greedy(V, **kw)
        np.testing.assert_array_equal(dofs, g[f'ei_{tag}_dofs'])                      # index work: exact
        np.testing.assert_allclose(basis.to_numpy(), g[f'ei_{tag}_basis'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(data['errors'], g[f'ei_{tag}_errors'], rtol=1e-9)
        np.testing.assert_allclose(data['interpolation_matrix'], g[f'ei_{tag}_matrix'], rtol=0, atol=1e-9)
        np.testing.assert_allclose(V.lincomb(data['coefficients']).to_numpy(), basis.to_numpy(), atol=1e-7)
        np.testing.assert_array_equal(V.to_numpy(), A)                                 # copy=True: input untouched
    for tag, prod in (('euclid', None), ('diag', CudaDiagonalOperator(g['w'], sp))):
        dofs, basis, data = alg.deim(sp.from_numpy(A), modes=9, product=prod)
        np.testing.assert_array_equal(dofs, g[f'deim_{tag}_dofs'])
        np.testing.assert_allclose(data['svals'], g[f'deim_{tag}_svals'], rtol=1e-10)
        B = basis.to_numpy()
        signs = np.sign(np.sum(B * g[f'deim_{tag}_basis'], axis=0))
        np.testing.assert_allclose(B * signs, g[f'deim_{tag}_basis'], rtol=0, atol=1e-8)
    dofs, basis, _ = alg.deim(sp.from_numpy(A[:, :6]), pod=False)
    np.testing.assert_array_equal(dofs, g['deim_nopod_dofs'])
    np.testing.assert_array_equal(basis.to_numpy(), g['deim_nopod_basis'])


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_standalone_gram_schmidt_fused_sweep(emu, name):
    """gram_schmidt(sweep=True): every pass of the inner loop is ONE pmc_mgs_sweep call (plus norm / scal per
    vector); products without a device weight vector fall back to the pair-by-pair loop.  Same Q, R."""
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp),
            'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    VA = sp.from_numpy(g['A'])
    emu.calls.clear()
    Q, R = alg.gram_schmidt(VA, product=prod, return_R=True, sweep=True)
    np.testing.assert_array_equal(VA.to_numpy(), g['A'])
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, W)     # (the pre-weighted panel rounds (w q) v, not q (w v))
    k = g['A'].shape[1]
    if name == 'csr':
        assert 'pmc_mgs_sweep' not in emu.calls
    else:
        assert k - 1 <= emu.calls.count('pmc_mgs_sweep') <= 4 * (k - 1)      # about one per pass, not one per pair
        assert 'pmc_axpy_dot' not in emu.calls and 'pmc_axpy' not in emu.calls
        assert emu.calls.count('pmc_pairwise_inner') == 0
        # weighted: every column that became final was pre-weighted once (w * q_j), the sweeps read that panel
        assert emu.calls.count('pmc_diag_apply') == (len(Q) if name == 'diag' else 0)
    # views, removal of dependent vectors and the offset form keep working (fallbacks inside)
    A = g['A'].copy()
    A[:, 3] = A[:, 1] - 2 * A[:, 2]
    Qd, Rd = alg.gram_schmidt(sp.from_numpy(A), product=prod, return_R=True, sweep=True)
    Qo, Ro = O.gram_schmidt(A, None if name == 'euclid' else (g['w'] if name == 'diag' else csr_from(g, 'M')), return_R=True)
    assert len(Qd) == Qo.shape[1] < A.shape[1]
    np.testing.assert_allclose(Qd.to_numpy(), Qo, atol=1e-6)            # near-dependent column 4: cond * eps
    np.testing.assert_allclose(Rd, Ro, atol=1e-5)
    np.testing.assert_allclose(Qd.to_numpy() @ Rd, A, atol=1e-12)
    if name == 'euclid':
        Q0 = alg.gram_schmidt(sp.from_numpy(g['A'][:, :4]), sweep=True)
        Q0.append(sp.from_numpy(g['A'][:, 10:]))
        Q1 = alg.gram_schmidt(Q0, offset=4, copy=False, sweep=True)
        np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-11)
        V = sp.from_numpy(g['A'])
        Qv = alg.gram_schmidt(V[2:9], sweep=True)                       # a view: pair-by-pair path
        np.testing.assert_allclose(Qv.to_numpy(), O.gram_schmidt(g['A'][:, 2:9]), atol=1e-11)


def test_fused_sweep_randomised(emu, rng):
    """Swept and pair-by-pair Gram-Schmidt take the same decisions (removed vectors, re-iterations) on random
    inputs with zero, dependent and strongly overlapping columns; A = Q R and orthonormality hold for both."""
    from pymor_b200.operators import CudaDiagonalOperator
    for _ in range(150):
        dim = int(rng.integers(1, 40))
        k = int(rng.integers(1, min(dim, 12) + 1))
        A = rng.standard_normal((dim, k))
        for _ in range(int(rng.integers(0, 3))):
            j, kind = int(rng.integers(0, k)), int(rng.integers(0, 3))
            if kind == 0:
                A[:, j] = 0
            elif kind == 1 and j > 0:
                A[:, j] = A[:, :j] @ rng.standard_normal(j)
            else:
                A[:, j] += 5 * A[:, 0]
        sp = space_of(dim)
        prod = CudaDiagonalOperator(rng.uniform(0.5, 1.5, dim), sp) if rng.integers(0, 2) else None
        Q0, R0 = alg.gram_schmidt(sp.from_numpy(A), product=prod, return_R=True, sweep=False)
        Q1, R1 = alg.gram_schmidt(sp.from_numpy(A), product=prod, return_R=True, sweep=True)
        assert len(Q0) == len(Q1)
        np.testing.assert_allclose(Q1.to_numpy() @ R1, A, atol=1e-11)
        np.testing.assert_allclose(Q0.to_numpy() @ R0, A, atol=1e-11)
        if len(Q1):
            assert np.abs(Q1.gramian(prod) - np.eye(len(Q1))).max() < 1e-13


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_standalone_gram_schmidt_cgs2(emu, rng, name):
    """variant='cgs2': one block projection (streaming inner with a broadcast column) + one block lincomb per pass
    instead of one dot + axpy per previous vector; converges to the same (unique) QR factorisation."""
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp),
            'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    VA = sp.from_numpy(g['A'])
    emu.calls.clear()
    Q, R = alg.gram_schmidt(VA, product=prod, return_R=True, variant='cgs2')
    np.testing.assert_array_equal(VA.to_numpy(), g['A'])
    assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, W, tight=1e-11)
    k = g['A'].shape[1]
    assert emu.calls.count('pmc_lincomb') <= 4 * k and 'pmc_axpy_dot' not in emu.calls     # per pass, not per pair
    if name == 'euclid':
        Q0 = alg.gram_schmidt(sp.from_numpy(g['A'][:, :4]), variant='cgs2')
        Q0.append(sp.from_numpy(g['A'][:, 10:]))
        Q1 = alg.gram_schmidt(Q0, offset=4, copy=False, variant='cgs2')
        np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-11)
    for _ in range(60):                                  # random inputs with degeneracies: same decisions as MGS
        dim = int(rng.integers(1, 40))
        kk = int(rng.integers(1, min(dim, 12) + 1))
        A = rng.standard_normal((dim, kk))
        for _ in range(int(rng.integers(0, 3))):
            j, kind = int(rng.integers(0, kk)), int(rng.integers(0, 3))
            if kind == 0:
                A[:, j] = 0
            elif kind == 1 and j > 0:
                A[:, j] = A[:, :j] @ rng.standard_normal(j)
            else:
                A[:, j] += 5 * A[:, 0]
        sp2 = space_of(dim)
        p2 = None if name == 'euclid' else CudaDiagonalOperator(rng.uniform(0.5, 1.5, dim), sp2)
        Q0, R0 = alg.gram_schmidt(sp2.from_numpy(A), product=p2, return_R=True)
        Q1, R1 = alg.gram_schmidt(sp2.from_numpy(A), product=p2, return_R=True, variant='cgs2')
        assert len(Q0) == len(Q1)
        np.testing.assert_allclose(Q1.to_numpy() @ R1, A, atol=1e-11)
        if len(Q1):
            assert np.abs(Q1.gramian(p2) - np.eye(len(Q1))).max() < 1e-13


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_standalone_block_gram_schmidt(emu, rng, name):
    """BCGS2: panels against all previous vectors through the Gramian / lincomb entry points, twice; same
    (unique) QR factorisation, same removed vectors, a handful of launches per panel instead of two per pair."""
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp),
            'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    for panel in (1, 4, 5, 32, (8, 2), (6, 3, 2)):
        VA = sp.from_numpy(g['A'])
        Q, R = alg.gram_schmidt(VA, product=prod, return_R=True, variant='bcgs2') if panel == 32 else \
            alg.block_gram_schmidt(VA, product=prod, return_R=True, panel=panel)
        np.testing.assert_array_equal(VA.to_numpy(), g['A'])
        assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, W, tight=1e-11)
        assert np.allclose(np.tril(R[:, :R.shape[0]], -1)[:4, :4], 0)
    if name == 'euclid':
        Q0 = alg.gram_schmidt(sp.from_numpy(g['A'][:, :4]), variant='bcgs2')
        Q0.append(sp.from_numpy(g['A'][:, 10:]))
        Q1 = alg.block_gram_schmidt(Q0, offset=4, copy=False, panel=3)
        assert Q1 is Q0
        np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-11)
    for _ in range(80):
        dim = int(rng.integers(1, 60))
        kk = int(rng.integers(1, min(dim, 20) + 1))
        A = rng.standard_normal((dim, kk))
        for _ in range(int(rng.integers(0, 3))):
            j, kind = int(rng.integers(0, kk)), int(rng.integers(0, 3))
            if kind == 0:
                A[:, j] = 0
            elif kind == 1 and j > 0:
                A[:, j] = A[:, :j] @ rng.standard_normal(j)
            else:
                A[:, j] += 5 * A[:, 0]
        sp2 = space_of(dim)
        w2 = rng.uniform(0.5, 1.5, dim)
        p2 = None if name == 'euclid' else CudaDiagonalOperator(w2, sp2)
        Q0, R0 = alg.gram_schmidt(sp2.from_numpy(A), product=p2, return_R=True)
        panel = [int(rng.integers(1, 9)), (8, 3), (16, 4, 2), (5, 5)][int(rng.integers(0, 4))]
        Q1, R1 = alg.block_gram_schmidt(sp2.from_numpy(A), product=p2, return_R=True, panel=panel)
        assert len(Q0) == len(Q1) and R1.shape == R0.shape
        np.testing.assert_allclose(Q1.to_numpy() @ R1, A, atol=1e-10)
        if len(Q1):
            assert np.abs(Q1.gramian(p2) - np.eye(len(Q1))).max() < 1e-12
            overlap = np.sum(Q1.to_numpy() * ((w2[:, None] if p2 is not None else 1.0) * Q0.to_numpy()), axis=0)
            np.testing.assert_allclose(overlap, 1.0, atol=1e-6)            # same vectors, same signs


@pytest.mark.parametrize('name', ['euclid', 'diag'])
def test_pod_qr_svd_with_block_gram_schmidt(emu, name):
    """pod(method='qr_svd', gs_variant='bcgs2'): the reference's default POD method with the QR step on the
    tensor-core kernels -- same singular values and subspace as the reference's result."""
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('pod')
    A = g['A']
    sp = space_of(A.shape[0])
    W = None if name == 'euclid' else g['w']
    prod = None if name == 'euclid' else CudaDiagonalOperator(g['w'], sp)
    U_ref, s_ref = O.pod(A, W, method='qr_svd')
    for variant in ('mgs', 'cgs2', 'bcgs2'):
        U, s = alg.pod(sp.from_numpy(A), product=prod, method='qr_svd', gs_variant=variant)
        assert len(U) == U_ref.shape[1]
        np.testing.assert_allclose(s, s_ref, rtol=1e-10)
        assert O.subspace_angle(U_ref, U.to_numpy(), W) < 1e-8
        assert O.orthogonality_error(U.to_numpy(), W) < 1e-10


def test_device_solvers(emu, rng):
    """apply_inverse on the CUDA operators: conjugate gradients on the whole block of right-hand sides (SpMM +
    pairwise_inner + per-column axpy / scal), Jacobi-preconditioned, converged columns frozen; exact inverse for
    diagonal operators; non-symmetric matrices get no default solver."""
    import scipy.sparse.linalg as spla

    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    from pymor_b200.solvers import CgSolver, InversionError, cg
    dim = 300
    h = 1 / (dim + 1)
    K = (sps.diags([-np.ones(dim - 1), 2.0 * np.ones(dim), -np.ones(dim - 1)], [-1, 0, 1]) / h
         + sps.diags(np.linspace(1, 3, dim))).tocsr()
    sp = space_of(dim)
    op = CudaCsrOperator(K, sp)
    assert op.solver is None and isinstance(op.default_solver(), CgSolver)
    V = rng.standard_normal((dim, 5))
    V[:, 3] = 0                                               # a zero right-hand side converges at once
    ref = spla.spsolve(K.tocsc(), V)
    U, info = op.apply_inverse(sp.from_numpy(V), return_info=True)
    assert 0 < info['iterations'] <= dim
    np.testing.assert_allclose(U.to_numpy(), ref, atol=1e-9 * np.abs(ref).max())
    assert np.all(U.to_numpy()[:, 3] == 0)
    np.testing.assert_allclose(op.apply_inverse_adjoint(sp.from_numpy(V)).to_numpy(), ref, atol=1e-9 * np.abs(ref).max())
    # warm start and explicit solver; without the preconditioner more iterations, same answer
    U2, info2 = op.apply_inverse(sp.from_numpy(V), initial_guess=U, return_info=True, solver=CgSolver(tol=1e-10))
    assert info2['iterations'] <= 2
    U3, info3 = CgSolver(tol=1e-12, jacobi=False).solve(op, sp.from_numpy(V), return_info=True)
    np.testing.assert_allclose(U3.to_numpy(), ref, atol=1e-9 * np.abs(ref).max())
    w = rng.uniform(0.5, 1.5, dim)
    D = CudaDiagonalOperator(w, sp)
    np.testing.assert_allclose(D.apply_inverse(sp.from_numpy(V)).to_numpy(), V / w[:, None], rtol=1e-15)
    assert CudaCsrOperator(K + sps.diags(np.ones(dim - 1), 1), sp).default_solver() is None
    with pytest.raises(InversionError):                       # indefinite operator: non-positive curvature
        cg(CudaCsrOperator(sps.diags(np.where(np.arange(dim) % 2, 1.0, -1.0)).tocsr(), sp).apply, sp.from_numpy(V))
    with pytest.raises(InversionError):
        CgSolver(tol=1e-14, maxiter=3, jacobi=False).solve(op, sp.from_numpy(V))


def test_block_variants_keep_orthogonality_on_ill_conditioned_input(emu, rng):
    """"Orthogonality error no worse than the reference's" (north star): on inputs of condition 1e2 ... 1e13 the
    classical (cgs2) and block (bcgs2) variants with re-orthogonalisation stay at O(eps) like the reference's
    re-iterated MGS, and A = Q R holds to rounding."""
    from pymor_b200.operators import CudaDiagonalOperator
    dim, k = 600, 40
    sp = space_of(dim)
    prod = CudaDiagonalOperator(rng.uniform(0.5, 1.5, dim), sp)
    for cond in (1e2, 1e8, 1e13):
        U, _ = np.linalg.qr(rng.standard_normal((dim, k)))
        V, _ = np.linalg.qr(rng.standard_normal((k, k)))
        A = U @ np.diag(np.logspace(0, -np.log10(cond), k)) @ V.T
        ref_Q = alg.gram_schmidt(sp.from_numpy(A), product=prod)
        ref = np.abs(ref_Q.gramian(prod) - np.eye(len(ref_Q))).max()
        for kw in (dict(variant='cgs2'), dict(variant='bcgs2'), dict(variant='mgs', sweep=True)):
            Q, R = alg.gram_schmidt(sp.from_numpy(A), product=prod, return_R=True, **kw)
            assert len(Q) == len(ref_Q) == k
            assert np.abs(Q.gramian(prod) - np.eye(k)).max() <= max(4 * ref, 5e-15), (cond, kw)
            np.testing.assert_allclose(Q.to_numpy() @ R, A, atol=1e-14)
        Q, R = alg.block_gram_schmidt(sp.from_numpy(A), product=prod, return_R=True, panel=(16, 4))
        assert np.abs(Q.gramian(prod) - np.eye(k)).max() <= max(4 * ref, 5e-15)
