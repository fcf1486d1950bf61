"""Parity of the CUDA path (through the C ABI, real kernels) with the CPU oracle and the golden
fixtures produced by the reference.  All tests here need a B200 (``-m gpu``).

Tolerances (north star): Gram and lincomb entries relative 1e-12 (relative to ||a_i|| ||b_j|| resp.
||A[d,:]|| ||c_j||), singular values relative 1e-10, subspace angles < 1e-8,
orthogonality error no worse than the reference's.  Index / byte work (amax indices, dofs, copies) is
bit-exact.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sps

from conftest import assert_gram_schmidt_golden, csr_from, load_golden
from oracle import numpy_path as O

pytestmark = pytest.mark.gpu

GRAM_TOL = 1e-12
LINCOMB_TOL = 1e-12


@pytest.fixture(scope='module')
def loaded_lib():
    """The product path must be the native library, loaded from the tree."""
    from pymor_b200 import _lib
    _lib.set_library_for_testing(None)
    lib = _lib.load_library()
    assert 'libpymor_b200.so' in str(lib._name)
    return lib


def space_of(dim):
    from pymor_b200.vectorarray import CudaVectorSpace
    return CudaVectorSpace(dim)


def test_all_symbols_and_fp64_peak(loaded_lib):
    from pymor_b200 import _lib
    for name in _lib.EXPORTED_SYMBOLS:
        assert hasattr(loaded_lib, name)
    ctx = _lib.get_context('cuda')
    assert ctx.sm_count >= 100
    dmma = ctx.measure_fp64_peak(True, 50.0)
    dfma = ctx.measure_fp64_peak(False, 50.0)
    print(f'FP64 peak: DMMA {dmma:.1f} TFLOP/s, DFMA {dfma:.1f} TFLOP/s')
    assert dmma > 5 and dfma > 5


# (dim, kA, kB): edge sizes, ragged tiles, k below / at / above the 128 tile, n not a multiple of 16
SHAPES = [(1, 1, 1), (15, 3, 2), (16, 16, 16), (1000, 17, 130), (4099, 128, 128), (5000, 129, 257),
          (33333, 200, 64), (70001, 500, 500)]


@pytest.mark.parametrize('dim,ka,kb', SHAPES)
@pytest.mark.parametrize('weighted', [False, True])
def test_gram_parity(loaded_lib, rng, dim, ka, kb, weighted):
    sp = space_of(dim)
    A, B = rng.standard_normal((dim, ka)), rng.standard_normal((dim, kb))
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    w = rng.uniform(0.5, 1.5, dim) if weighted else None
    prod = None
    if weighted:
        from pymor_b200.operators import CudaDiagonalOperator
        prod = CudaDiagonalOperator(w, sp)
    for generic in (False, True):                       # TMA+DMMA path and the generic kernel
        sp.force_generic_kernels(generic)
        G = VA.inner(VB, prod)
        assert G.shape == (ka, kb) and G.flags.f_contiguous
        assert O.gram_rel_error(G, O.inner(A, B, w), A, B, w) < GRAM_TOL
        S = VA.gramian(prod)
        assert O.gram_rel_error(S, O.gramian(A, w), A, A, w) < GRAM_TOL
        np.testing.assert_array_equal(S, S.T)           # bitwise symmetric like the reference's dsyrk
        S2 = VA.gramian(prod)
        np.testing.assert_array_equal(S, S2)            # fixed summation order: reproducible
        # same buffer, different columns = plain GEMM (gram_schmidt's check)
        off = ka // 3
        G2 = VA[off:].inner(VA, prod)
        assert O.gram_rel_error(G2, O.inner(A[:, off:], A, w), A[:, off:], A, w) < GRAM_TOL
    sp.force_generic_kernels(False)


# Second-generation Gram tile kernel (pymor_b200/csrc/gram2.cu, gram_opts bit 16) against the oracle and the first
# kernel, for every ablation bit incl. both ways of applying the diagonal weight (bit 256: in shared memory).
def _block_error_map(G, G_ref, scale):
    """Max scaled error per 32x32 block (the unit a warp of the tile kernels owns) -- printed when a Gram parity
    assertion fails, so that one GPU session is enough to tell WHICH block / tile kind is wrong."""
    err = np.abs(G - G_ref) / scale
    nb_i, nb_j = -(-G.shape[0] // 32), -(-G.shape[1] // 32)
    m = np.zeros((nb_i, nb_j))
    for bi in range(nb_i):
        for bj in range(nb_j):
            m[bi, bj] = np.nanmax(np.nan_to_num(err[32 * bi:32 * bi + 32, 32 * bj:32 * bj + 32], nan=np.inf))
    with np.printoptions(precision=1, linewidth=250, threshold=100000):
        return f'max scaled error per 32x32 block (rows = block row i, 4 blocks = one 128 tile):\n{m}'


@pytest.mark.parametrize('opts', [16, 16 | 32, 16 | 64, 16 | 128, 16 | 32 | 64 | 128, 16 | 256, 16 | 64 | 256, 16 | 512,
                                  16 | 64 | 512, 16 | 1024, 16 | 64 | 1024, 16 | 2048, 16 | 64 | 2048])
def test_gram2_variant_parity(loaded_lib, rng, opts):
    from pymor_b200._lib import get_context
    from pymor_b200.operators import CudaDiagonalOperator
    ctx = get_context('cuda')
    shapes = SHAPES[2:] + [(2100, 232, 232), (3000, 1000, 1000), (200003, 256, 256)]
    try:
        for dim, ka, kb in shapes:
            sp = space_of(dim)
            A, B = rng.standard_normal((dim, ka)), rng.standard_normal((dim, kb))
            VA, VB = sp.from_numpy(A), sp.from_numpy(B)
            w = rng.uniform(0.5, 1.5, dim)
            for prod, wv in ((None, None), (CudaDiagonalOperator(w, sp), w)):
                ctx.set_option('gram_opts', 0)                  # first-generation kernel
                G0, S0 = VA.inner(VB, prod), VA.gramian(prod)
                ctx.set_option('gram_opts', opts)
                G, S = VA.inner(VB, prod), VA.gramian(prod)
                tag = f'dim={dim} ka={ka} kb={kb} weighted={wv is not None} opts={opts}'
                na, nb = O.norm(A, wv), O.norm(B, wv)
                assert O.gram_rel_error(G, O.inner(A, B, wv), A, B, wv) < GRAM_TOL, \
                    f'general product, {tag}\n' + _block_error_map(G, O.inner(A, B, wv), np.outer(na, nb))
                assert O.gram_rel_error(S, O.gramian(A, wv), A, A, wv) < GRAM_TOL, \
                    f'SYRK, {tag}\n' + _block_error_map(S, O.gramian(A, wv), np.outer(na, na))
                np.testing.assert_array_equal(S, S.T)
                np.testing.assert_array_equal(S, VA.gramian(prod))
                # same products, same order inside a split: without the K-splits (bit 64) both results are bitwise
                # the first kernel's; with them the split blocks (two per full diagonal tile, every block of a
                # narrow or ragged tile) add their K parts in a different order
                if opts & 64:
                    np.testing.assert_array_equal(G, G0)
                    np.testing.assert_array_equal(S, S0)
                else:
                    assert O.gram_rel_error(G, G0, A, B, wv) < 1e-13
                    assert O.gram_rel_error(S, S0, A, A, wv) < 1e-13
    finally:
        ctx.set_option('gram_opts', 16)                     # the default


@pytest.mark.parametrize('dim,k,m', [(1, 1, 1), (17, 5, 3), (16, 4, 16), (1000, 130, 17), (4099, 128, 128),
                                     (5000, 257, 129), (33333, 64, 200), (70001, 300, 300)])
def test_lincomb_parity(loaded_lib, rng, dim, k, m):
    sp = space_of(dim)
    A, C = rng.standard_normal((dim, k)), rng.standard_normal((k, m))
    VA = sp.from_numpy(A)
    ref = O.lincomb(A, C)
    for generic in (False, True):
        sp.force_generic_kernels(generic)
        R = VA.lincomb(C).to_numpy()
        assert O.lincomb_rel_error(R, ref, A, C) < LINCOMB_TOL
        R2 = VA[::2].lincomb(C[::2]).to_numpy()          # strided view (TMA stride = 2 columns)
        assert O.lincomb_rel_error(R2, O.lincomb(A[:, ::2], C[::2]), A[:, ::2], C[::2]) < LINCOMB_TOL
    sp.force_generic_kernels(False)


@pytest.mark.parametrize('dim', [1, 2, 3, 2047, 2048, 2049, 100003, 3000001])
def test_streaming_kernels_parity(loaded_lib, rng, dim):
    from pymor_b200.operators import CudaDiagonalOperator
    sp = space_of(dim)
    k = 5 if dim < 1000000 else 2
    A, B = rng.standard_normal((dim, k)), rng.standard_normal((dim, k))
    w = rng.uniform(0.5, 1.5, dim)
    VA, VB, prod = sp.from_numpy(A), sp.from_numpy(B), CudaDiagonalOperator(w, sp)
    scale = np.sqrt(dim)
    np.testing.assert_allclose(VA.pairwise_inner(VB), O.pairwise_inner(A, B), atol=1e-13 * dim, rtol=1e-13)
    np.testing.assert_allclose(VA.pairwise_inner(VB, prod), O.pairwise_inner(A, B, w), atol=1e-13 * dim, rtol=1e-13)
    np.testing.assert_allclose(VA.norm2(), O.norm2(A), rtol=1e-12)
    np.testing.assert_allclose(VA.norm(prod), O.norm(A, w), rtol=1e-12)
    np.testing.assert_allclose(VA[1].pairwise_inner(VB[0]), O.pairwise_inner(A[:, [1]], B[:, [0]]),
                               atol=1e-13 * dim)
    al = rng.standard_normal(k)
    Y = VA.copy(); Y.axpy(al, VB)
    np.testing.assert_allclose(Y.to_numpy(), A + B * al, rtol=0, atol=1e-15 * scale)
    Y = VA.copy(); Y.axpy(-0.75, VB[0])
    np.testing.assert_allclose(Y.to_numpy(), A - 0.75 * B[:, [0]], rtol=0, atol=1e-15 * scale)
    Y = VA.copy(); Y[::-1].scal(al)
    np.testing.assert_array_equal(Y.to_numpy(), A * al[::-1])
    np.testing.assert_array_equal(VA[[k - 1, 0, 0]].copy().to_numpy(), A[:, [k - 1, 0, 0]])
    np.testing.assert_array_equal(prod.apply(VA).to_numpy(), w[:, None] * A)
    idx, val = VA.amax()
    oi, ov = O.amax(A)
    np.testing.assert_array_equal(idx, oi)
    np.testing.assert_array_equal(val, ov)
    dofs = [0, dim - 1, dim // 2, dim // 2]
    np.testing.assert_array_equal(VA.dofs(dofs), O.dofs(A, dofs))


def test_amax_ties_first_index(loaded_lib):
    sp = space_of(5000)
    A = np.zeros((5000, 3))
    A[[7, 4000], 0] = [-2.0, 2.0]       # tie: first index wins
    A[4999, 1] = 1.0
    VA = sp.from_numpy(A)
    idx, val = VA.amax()
    np.testing.assert_array_equal(idx, [7, 4999, 0])
    np.testing.assert_array_equal(val, [2.0, 1.0, 0.0])


def test_misaligned_dlpack_views_take_generic_path(loaded_lib, rng):
    """Odd leading dimension / misaligned base: every kernel must still be correct."""
    import torch
    dim, k = 1000, 6
    host = rng.standard_normal((k, dim + 1))
    t = torch.from_numpy(host).cuda()[:, 1:]          # base misaligned by 8 bytes, ld odd
    from pymor_b200.vectorarray import CudaVectorArray, CudaVectorArrayImpl
    sp = space_of(dim)
    V = CudaVectorArray(sp, CudaVectorArrayImpl(sp, t, k))     # force adoption without the alignment check
    A = host[:, 1:].T
    assert O.gram_rel_error(V.gramian(), O.gramian(A), A, A) < GRAM_TOL
    np.testing.assert_allclose(V.norm2(), O.norm2(A), rtol=1e-13)
    C = rng.standard_normal((k, 20))
    assert O.lincomb_rel_error(V.lincomb(C).to_numpy(), O.lincomb(A, C), A, C) < LINCOMB_TOL
    Y = V.copy(); Y.axpy(2.0, V[::-1])
    np.testing.assert_allclose(Y.to_numpy(), A + 2 * A[:, ::-1], atol=1e-14)


def test_vecops_golden(loaded_lib):
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('vecops')
    A, B, C, alpha = g['A'], g['B'], g['C'], g['alpha']
    sp = space_of(A.shape[0])
    VA, VB = sp.from_numpy(A), sp.from_numpy(B)
    Wd, Wm = CudaDiagonalOperator(g['w'], sp), CudaCsrOperator(csr_from(g, 'M'), sp)
    tight = dict(rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(VA.inner(VB), g['inner'], **tight)
    np.testing.assert_allclose(VA.inner(VB, Wd), g['inner_diag'], **tight)
    np.testing.assert_allclose(VA.inner(VB, Wm), g['inner_csr'], **tight)
    np.testing.assert_allclose(VA.gramian(), g['gramian'], **tight)
    np.testing.assert_allclose(VA.gramian(Wd), g['gramian_diag'], **tight)
    np.testing.assert_allclose(VA.gramian(Wm), g['gramian_csr'], **tight)
    lb = B.shape[1]
    np.testing.assert_allclose(VA[:lb].pairwise_inner(VB), g['pairwise'], **tight)
    np.testing.assert_allclose(VA[:lb].pairwise_inner(VB, Wd), g['pairwise_diag'], **tight)
    np.testing.assert_allclose(VA[:lb].pairwise_inner(VB, Wm), g['pairwise_csr'], **tight)
    np.testing.assert_allclose(VA.lincomb(C).to_numpy(), g['lincomb'], **tight)
    np.testing.assert_allclose(VA.norm(), g['norm'], **tight)
    np.testing.assert_allclose(VA.norm(Wd), g['norm_diag'], **tight)
    np.testing.assert_allclose(VA.norm2(Wm), g['norm2_csr'], **tight)
    Y = VA.copy(); Y.axpy(0.37, VA[::-1]); np.testing.assert_allclose(Y.to_numpy(), g['axpy_scalar_rev'], **tight)
    Y = VA.copy(); Y.axpy(alpha, VB[0]); np.testing.assert_allclose(Y.to_numpy(), g['axpy_vec_bcast'], **tight)
    Y = VA.copy(); Y[[5, 1, 3]].axpy(alpha[:3], VB[[0, 4, 2]]); np.testing.assert_allclose(Y.to_numpy(), g['axpy_list'], **tight)
    Y = VA.copy(); Y.scal(alpha); np.testing.assert_allclose(Y.to_numpy(), g['scal_vec'], **tight)
    Y = VA.copy(); Y[1:6:2].scal(-2.5); np.testing.assert_allclose(Y.to_numpy(), g['scal_slice'], **tight)
    np.testing.assert_array_equal(VA[[6, 0, 2]].dofs(g['dof_idx']), g['dofs'])
    ai, av = VA.amax()
    np.testing.assert_array_equal(ai, g['amax_idx'])
    np.testing.assert_array_equal(av, g['amax_val'])
    np.testing.assert_allclose(VA[[6, 0, 0, 2]].inner(VB[3:0:-1]), g['view_inner'], **tight)
    Z = VA.copy(); Z.append(VB[[1, 1, 4]]); del Z[[0, 2]]
    np.testing.assert_array_equal(Z.to_numpy(), g['append_del'])
    np.testing.assert_allclose(Wm.apply(VA).to_numpy(), g['apply_csr'], **tight)


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_gram_schmidt_golden(loaded_lib, name):
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp), 'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    Q, R = alg.gram_schmidt(sp.from_numpy(g['A']), product=prod, return_R=True)
    Qh, Qref = Q.to_numpy(), g[f'Q_{name}']
    assert Qh.shape == Qref.shape                          # the same vectors were removed
    # Column 4 of the input is column 1 plus 1e-9 noise: its orthogonalised direction q_4 is the
    # normalised O(1e-9) remainder, i.e. conditioned like 1e9 * eps w.r.t. ANY change of summation
    # order (the reference itself moves by that much between BLAS builds); the coefficients R[4, :]
    # along it and, through them, the later vectors inherit a perturbation of that size.  So this
    # (deliberately nasty) input is compared at cond * eps; the well-conditioned sub-problem below and
    # the invariants are pinned tightly.
    np.testing.assert_allclose(R, g[f'R_{name}'], atol=1e-5)
    np.testing.assert_allclose(Qh, Qref, atol=1e-6)
    first = [0, 1, 2, 3]                                   # before the near-dependency: exact parity
    np.testing.assert_allclose(Qh[:, first], Qref[:, first], atol=1e-12)
    np.testing.assert_allclose(R[:4, :4], g[f'R_{name}'][:4, :4], atol=1e-11)
    np.testing.assert_allclose(Qh @ R, g['A'], atol=1e-12)     # A = Q R holds to rounding (removed columns too)
    # orthogonality error no worse than the reference's (up to a factor for rounding noise)
    assert O.orthogonality_error(Q.to_numpy(), W) <= max(4 * O.orthogonality_error(g[f'Q_{name}'], W), 1e-14)
    if name == 'euclid':
        # well-conditioned sub-problem incl. the `offset` path used by reductors' extend_basis
        Q0 = alg.gram_schmidt(sp.from_numpy(g['A'][:, :4]))
        Q0.append(sp.from_numpy(g['A'][:, 10:]))
        Q1 = alg.gram_schmidt(Q0, offset=4, copy=False)
        np.testing.assert_allclose(Q1.to_numpy(), g['Q_offset'], atol=1e-11)


@pytest.mark.parametrize('name', ['euclid', 'diag', 'csr'])
def test_pod_golden(loaded_lib, name):
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('pod')
    sp = space_of(g['A'].shape[0])
    W = {'euclid': None, 'diag': g['w'], 'csr': csr_from(g, 'M')}[name]
    prod = {'euclid': None, 'diag': CudaDiagonalOperator(g['w'], sp), 'csr': CudaCsrOperator(csr_from(g, 'M'), sp)}[name]
    VA = sp.from_numpy(g['A'])
    U, s, Vh = alg.method_of_snapshots(VA, product=prod)
    np.testing.assert_allclose(s, g[f'mos_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'mos_U_{name}'], U.to_numpy(), W) < 1e-8
    U, s, _ = alg.qr_svd(VA, product=prod)
    np.testing.assert_allclose(s, g[f'qr_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'qr_U_{name}'], U.to_numpy(), W) < 1e-8
    U, s = alg.pod(VA, product=prod, modes=10, method='method_of_snapshots')
    np.testing.assert_allclose(s, g[f'pod_s_{name}'], rtol=1e-10)
    assert O.subspace_angle(g[f'pod_U_{name}'], U.to_numpy(), W) < 1e-8
    assert O.orthogonality_error(U.to_numpy(), W) <= max(4 * O.orthogonality_error(g[f'pod_U_{name}'], W), 1e-13)


def test_projection_and_thermalblock_golden(loaded_lib):
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator
    g = load_golden('projection')
    sp = space_of(g['V'].shape[0])
    op, prod = CudaCsrOperator(csr_from(g, 'Mop'), sp), CudaCsrOperator(csr_from(g, 'P'), sp)
    V, U = sp.from_numpy(g['V']), sp.from_numpy(g['U'])
    np.testing.assert_allclose(alg.project(op, V, U).matrix, g['proj_both'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(alg.project(op, V, U, product=prod).matrix, g['proj_both_product'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(alg.project(op, U, U).matrix, g['proj_galerkin'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(alg.project(op, None, U).array.to_numpy(), g['proj_source_only'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(alg.project(op, V, None).array.inner(U), g['proj_range_only'], rtol=1e-12, atol=1e-12)
    t = load_golden('thermalblock')
    S, P = t['snapshots'], csr_from(t, 'P')
    sp = space_of(S.shape[0])
    prod = CudaCsrOperator(P, sp)
    basis, svals = alg.pod(sp.from_numpy(S), modes=5, product=prod)
    np.testing.assert_allclose(svals, t['pod_qr_s'], rtol=1e-10)
    assert O.subspace_angle(t['pod_qr_U'], basis.to_numpy(), P) < 1e-8
    basis2, svals2 = alg.pod(sp.from_numpy(S), modes=5, product=prod, method='method_of_snapshots')
    np.testing.assert_allclose(svals2, t['pod_mos_s'], rtol=1e-10)
    B = sp.from_numpy(t['pod_qr_U'])
    for i in range(int(t['n_ops'])):
        n = len(t[f'op{i}_indptr']) - 1
        Mi = sps.csr_matrix((t[f'op{i}_data'], t[f'op{i}_indices'], t[f'op{i}_indptr']), shape=(n, n))
        np.testing.assert_allclose(alg.project(CudaCsrOperator(Mi, sp), B, B).matrix, t[f'op{i}_proj'], rtol=1e-12, atol=1e-12)


def test_csr_twoform_multi_panel(loaded_lib, rng):
    """CSR two-form across several row panels (the fused V^T (M U) path) vs SciPy."""
    from pymor_b200.operators import CudaCsrOperator
    dim, kv, ku = 60000, 40, 600          # 600 columns -> ~10k-row panels -> 6 panels
    nz = dim
    M = (sps.diags([-np.ones(dim - 1), 2.5 * np.ones(dim), -np.ones(dim - 1)], [-1, 0, 1])
         + sps.coo_matrix((rng.standard_normal(nz), (rng.integers(0, dim, nz), rng.integers(0, dim, nz))),
                          shape=(dim, dim))).tocsr()
    sp = space_of(dim)
    V, U = rng.standard_normal((dim, kv)), rng.standard_normal((dim, ku))
    op = CudaCsrOperator(M, sp)
    G = op.apply2(sp.from_numpy(V), sp.from_numpy(U))
    MU = M @ U
    assert O.gram_rel_error(G, V.T @ MU, V, MU) < GRAM_TOL
    np.testing.assert_allclose(op.apply(sp.from_numpy(U[:, :7])).to_numpy(), MU[:, :7], rtol=1e-13, atol=1e-13)


def test_csr_pairwise_and_symmetric_twoform(loaded_lib, rng):
    """The fused pairwise two-form (pmc_csr_pairwise: no n x k temporary, fixed summation tree) and the
    lower-triangle-only two-form of a symmetric operator (PMC_LOWER_ONLY) against the oracle; the per-rank
    constructor `from_local_rows`."""
    from pymor_b200.operators import CudaCsrOperator
    for dim, k in ((1, 3), (777, 5), (60000, 300), (400003, 130)):
        M = (sps.diags([-np.ones(dim - 1), 2.5 * np.ones(dim), -np.ones(dim - 1)], [-1, 0, 1])).tocsr()
        if dim > 1000:
            # (sps.random draws without replacement from dim^2 candidates: terabytes at this size)
            nz = 4 * dim
            R = sps.coo_matrix((rng.standard_normal(nz), (rng.integers(0, dim, nz), rng.integers(0, dim, nz))),
                               shape=(dim, dim))
            M = (M + R + R.T).tocsr()                                  # symmetric, irregular rows
        sp = space_of(dim)
        V, U = rng.standard_normal((dim, k)), rng.standard_normal((dim, k))
        VV, UU = sp.from_numpy(V), sp.from_numpy(U)
        op = CudaCsrOperator(M, sp)
        MU = M @ U
        p = op.pairwise_apply2(VV, UU)
        ref = np.einsum('ij,ij->j', V, MU)
        scale = np.linalg.norm(V, axis=0) * np.linalg.norm(MU, axis=0)
        assert np.max(np.abs(p - ref) / scale) < GRAM_TOL
        np.testing.assert_array_equal(p, op.pairwise_apply2(VV, UU))      # fixed summation order
        np.testing.assert_allclose(UU.norm(op), np.sqrt(np.einsum('ij,ij->j', U, MU)), rtol=1e-12)
        assert op.is_symmetric()
        G = op.apply2(UU, UU)                                             # lower tiles only for k > 128
        assert O.gram_rel_error(G, U.T @ MU, U, MU) < GRAM_TOL
        if k > 128:
            np.testing.assert_array_equal(G, G.T)                         # mirrored => bitwise symmetric
        loc = CudaCsrOperator.from_local_rows(M, sp, symmetric=True)
        np.testing.assert_array_equal(loc.apply2(UU, UU), G)
        np.testing.assert_array_equal(loc.apply(UU[:3]).to_numpy(), op.apply(UU[:3]).to_numpy())
        G2 = op.apply2(VV, UU)                                            # general pair: full product
        assert O.gram_rel_error(G2, V.T @ MU, V, MU) < GRAM_TOL
        if dim == 60000:                                                  # the first-generation tile kernel as well
            sp.ctx.set_option('gram_opts', 0)
            try:
                G1 = op.apply2(UU, UU)
            finally:
                sp.ctx.set_option('gram_opts', 16)
            assert O.gram_rel_error(G1, U.T @ MU, U, MU) < GRAM_TOL
            np.testing.assert_array_equal(G1, G1.T)


@pytest.mark.parametrize('dim', [1, 7, 1001, 40000])
def test_delete_compacts_in_place(loaded_lib, rng, dim):
    """`del A[...]` (interface.py:252-259, numpy.py:47-57): the kept vectors move down inside the buffer, one
    pmc_shift_cols launch per run whatever the overlap -- bit-exact against NumPy's delete."""
    k = 300
    sp = space_of(dim)
    A = rng.standard_normal((dim, k))
    for drop in ([3], [0, 5, 6, 100, 299], list(range(1, 300, 2)), slice(10, 250), [-1, 0]):
        VA = sp.from_numpy(A)
        l0 = sp.ctx.launch_count
        del VA[drop]
        keep = np.delete(np.arange(k), np.arange(k)[drop] if isinstance(drop, slice) else [d % k for d in drop])
        np.testing.assert_array_equal(VA.to_numpy(), A[:, keep])
        runs = 1 + int(np.sum(np.diff(keep) > 1))
        assert sp.ctx.launch_count - l0 <= runs + 1


def test_kernel_timing_log(loaded_lib, rng):
    """pmc_ctx_kernel_times: one (kind, ms) pair per launch of the tensor-core / SpMM kernels while the option is on."""
    from pymor_b200._lib import get_context
    ctx = get_context('cuda')
    sp = space_of(50000)
    A = sp.from_numpy(rng.standard_normal((50000, 64)))
    ctx.kernel_times()
    A.gramian()
    assert ctx.kernel_times() == []                                       # off by default
    ctx.set_option('kernel_timing', 1)
    try:
        A.gramian(); A.lincomb(np.eye(64)); A.gramian()
        log = ctx.kernel_times()
    finally:
        ctx.set_option('kernel_timing', 0)
    assert [kind for kind, _ in log] == ['gram', 'lincomb', 'gram'] and all(0 < ms < 100 for _, ms in log)
    assert ctx.kernel_times() == []


def test_full_size_properties(loaded_lib):
    """Size-independent properties at a BASELINE-scale shape (config 2: 2M DOF x 500 snapshots):
    linearity of the Gramian, ||A c|| consistency between lincomb and Gramian, POD orthonormality and
    reconstruction of a known spectrum."""
    import torch
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    n, k = 2_000_000, 500
    sp = space_of(n)
    gen = torch.Generator(device='cuda').manual_seed(7)
    Gn = torch.randn((k, n), generator=gen, device='cuda', dtype=torch.float64) / np.sqrt(n)
    rng = np.random.default_rng(7)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    sig = 10.0 ** (-2 * np.arange(k) / (k - 1))
    base = sp.make_array(Gn)
    A = base.lincomb(np.diag(sig) @ Qk.T)            # A = Gn diag(sig) Qk^T
    del base, Gn
    w = torch.rand(n, generator=gen, device='cuda', dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, sp)
    G = A.gramian(prod)
    np.testing.assert_array_equal(G, G.T)
    # Gramian consistency: c^T G c == || A c ||_W^2
    c = rng.standard_normal((k, 3))
    Ac = A.lincomb(c)
    np.testing.assert_allclose(np.einsum('ia,ij,ja->a', c, G, c), Ac.norm2(prod), rtol=1e-11)
    # scaling linearity: gramian(2A) == 4 gramian(A) exactly (power of two)
    A2 = A.copy(); A2.scal(2.0)
    np.testing.assert_array_equal(A2.gramian(prod), 4 * G)
    del A2, Ac
    U, s = alg.pod(A, product=prod, method='method_of_snapshots')
    assert len(U) == k
    err = np.max(np.abs(U.gramian(prod) - np.eye(k)))
    assert err < 1e-10, err
    # the weighted spectrum is close to the prescribed one (random W ~ U(0.5,1.5) has mean 1)
    assert abs(s[0] / sig[0] - 1) < 0.05 and abs(s[-1] / sig[-1] - 1) < 0.2
    # reconstruction: A == U (U^T W A)
    coeff = U.inner(A[:8], prod)
    R = U.lincomb(coeff)
    R.axpy(-1.0, A[:8])
    assert np.max(R.norm(prod) / A[:8].norm(prod)) < 1e-9


def test_shifted_chol_qr_golden(loaded_lib):
    """"Next" row N1 on the device: CholeskyQR3 = 3 Gramians + 3 block lincombs, no per-pair host sync."""
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('chol_qr')
    for tag, tolQ in (('mild', 1e-11), ('ill', 1e-5)):
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
        assert O.orthogonality_error(Q.to_numpy(), w) < 1e-13
        np.testing.assert_allclose(Q.to_numpy() @ R, A, atol=1e-13)
    # at scale: 1M x 128, orthonormal to rounding, A = QR
    import torch
    n, k = 1_000_000, 128
    sp = space_of(n)
    A = sp.random_device(k, seed=5, scale=1.0 / np.sqrt(n))
    A.axpy(2.0, sp.random_device(1, seed=6, scale=1.0 / np.sqrt(n)))
    w = torch.rand(n, device='cuda', dtype=torch.float64) + 0.5
    prod = CudaDiagonalOperator(w, sp)
    Q, R = alg.shifted_chol_qr(A, product=prod, return_R=True, product_norm=float(np.sqrt(1.5)))
    assert np.max(np.abs(Q.gramian(prod) - np.eye(k))) < 1e-13
    D = Q.lincomb(R)
    D.axpy(-1.0, A)
    assert np.max(D.norm(prod) / A.norm(prod)) < 1e-13


@pytest.mark.parametrize('dim', [5, 2048, 100003, 2000000])
def test_fused_axpy_dot_bitwise_equals_separate_kernels(loaded_lib, rng, dim):
    """The deferred axpy merged with the following dot/norm (pmc_axpy_dot) must give the very bits of
    pmc_axpy + pmc_pairwise_inner / pmc_norm2 -- Gram-Schmidt results do not depend on the fusion."""
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    sp = space_of(dim)
    k = 6
    A = rng.standard_normal((dim, k)) + 2 * rng.standard_normal((dim, 1))
    w = rng.uniform(0.5, 1.5, dim)
    prod = CudaDiagonalOperator(w, sp)
    ctx = sp.ctx
    results = []
    for fuse in (True, False):
        ctx.fuse_axpy = fuse
        before = ctx.launch_count
        Q, R = alg.gram_schmidt(sp.from_numpy(A), product=prod, return_R=True)
        results.append((Q.to_numpy(), R, ctx.launch_count - before))
    ctx.fuse_axpy = True
    np.testing.assert_array_equal(results[0][0], results[1][0])
    np.testing.assert_array_equal(results[0][1], results[1][1])
    assert results[0][2] < results[1][2]                     # fewer launches with the fusion
    Qo, Ro = O.gram_schmidt(A, w, return_R=True)
    np.testing.assert_allclose(results[0][1], Ro, rtol=1e-11, atol=1e-11)


def test_complex_arrays_on_device(loaded_lib, rng):
    """Complex arrays (pairs of real device arrays, pymor_b200/complexarray.py) against NumPy."""
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    dim, k = 4099, 20
    sp = space_of(dim)
    A = rng.standard_normal((dim, k)) + 1j * rng.standard_normal((dim, k))
    B = rng.standard_normal((dim, k)) + 1j * rng.standard_normal((dim, k))
    w = rng.uniform(0.5, 1.5, dim)
    VA, VB, prod = sp.from_numpy(A), sp.from_numpy(B), CudaDiagonalOperator(w, sp)
    scale = np.outer(np.linalg.norm(A, axis=0), np.linalg.norm(B, axis=0))
    assert np.max(np.abs(VA.inner(VB) - A.conj().T @ B) / scale) < 1e-12
    assert np.max(np.abs(VA.inner(VB, prod) - A.conj().T @ (w[:, None] * B)) / scale) < 1e-12
    np.testing.assert_allclose(VA.pairwise_inner(VB), np.sum(A.conj() * B, axis=0), atol=1e-10)
    np.testing.assert_allclose(VA.norm(), np.linalg.norm(A, axis=0), rtol=1e-13)
    C = rng.standard_normal((k, 17)) + 1j * rng.standard_normal((k, 17))
    np.testing.assert_allclose(VA.lincomb(C).to_numpy(), A @ C, atol=1e-11)
    Y = sp.from_numpy(A.real); Y.scal(1j)                       # in-place promotion
    np.testing.assert_allclose(Y.to_numpy(), 1j * A.real, atol=1e-15)
    Y = VA.copy(); Y.axpy(0.5 - 2j, VB)
    np.testing.assert_allclose(Y.to_numpy(), A + (0.5 - 2j) * B, atol=1e-13)
    idx, val = VA.amax()
    np.testing.assert_array_equal(idx, np.argmax(np.abs(A), axis=0))
    Q, R = alg.gram_schmidt(VA, product=prod, return_R=True)
    np.testing.assert_allclose(Q.lincomb(R).to_numpy(), A, atol=1e-11)
    np.testing.assert_allclose(Q.gramian(prod), np.eye(k), atol=1e-12)


def test_gramian_while_uploading(loaded_lib, rng, monkeypatch):
    """from_local_numpy of a pinned shard uploads on a side stream; the Gramian contracts the column
    blocks as they land.  Same result as the resident path, symmetric, and every other consumer
    (norm, lincomb, axpy fast path) sees complete data."""
    import torch
    from pymor_b200 import vectorarray as va
    from pymor_b200.operators import CudaDiagonalOperator
    monkeypatch.setattr(va, '_STREAMED_UPLOAD_BYTES', 1 << 16)
    n, k = 30011, 200
    sp = space_of(n)
    host = torch.empty((k, n), dtype=torch.float64, pin_memory=True)
    host.copy_(torch.from_numpy(rng.standard_normal((k, n))))
    A = host.numpy().T
    w = rng.uniform(0.5, 1.5, n)
    prod = CudaDiagonalOperator(w, sp)
    ref = sp.from_numpy(np.array(A)).gramian(prod)            # pageable source: plain path
    V = sp.from_local_numpy(A)
    assert V.impl._arrivals is not None and len(V.impl._arrivals) > 1
    G = V.gramian(prod)
    assert V.impl._arrivals is None
    np.testing.assert_array_equal(G, G.T)
    assert O.gram_rel_error(G, ref, A, A, w) < GRAM_TOL
    assert O.gram_rel_error(G, O.gramian(A, w), A, A, w) < GRAM_TOL
    for consumer in (lambda X: X.norm(), lambda X: X[3].pairwise_inner(X[5]), lambda X: X.lincomb(np.eye(k)[:, :2]).norm(),
                     lambda X: X[:5].inner(X[7:9])):
        X = sp.from_local_numpy(A)
        got = consumer(X)
        want = consumer(sp.from_numpy(np.array(A)))
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)
    X = sp.from_local_numpy(A)
    X[0].axpy(2.0, X[1])
    np.testing.assert_allclose(X[0].to_numpy()[:, 0], A[:, 0] + 2.0 * A[:, 1], atol=1e-14)


def test_next_rows_hapod_and_range_finder(loaded_lib, rng):
    """N3 (incremental HAPOD vs the reference's golden result) and N2 (randomized range finder) on the
    device through the stand-alone drivers."""
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
    g = load_golden('hapod')
    A, w = g['A'], g['w']
    sp = space_of(A.shape[0])
    prod = CudaDiagonalOperator(w, sp)
    for steps in (1, 4, 7):
        modes, svals, count = alg.inc_vectorarray_hapod(steps, sp.from_numpy(A), 1e-3, 0.75, product=prod)
        ref_modes, ref_svals = g[f'modes_diag_{steps}'], g[f'svals_diag_{steps}']
        # the truncation is a threshold decision on a tail sum: allow the count to differ by one
        assert count == A.shape[1] and abs(len(modes) - ref_modes.shape[1]) <= 1
        m = min(len(modes), ref_modes.shape[1])
        np.testing.assert_allclose(svals[:m], ref_svals[:m], rtol=1e-7)
        assert O.subspace_angle(ref_modes[:, :7], modes.to_numpy()[:, :7], w) < 1e-6      # the 7 signal modes
    for slices, arity in ((5, None), (7, 2)):                     # distributed trees (hapod.py:372)
        modes, svals, count = alg.dist_vectorarray_hapod(slices, sp.from_numpy(A), 1e-3, 0.75, arity=arity, product=prod)
        ref_modes, ref_svals = g[f'dist_modes_diag_{slices}_{arity}'], g[f'dist_svals_diag_{slices}_{arity}']
        assert count == A.shape[1] and abs(len(modes) - ref_modes.shape[1]) <= 1
        m = min(len(modes), ref_modes.shape[1])
        np.testing.assert_allclose(svals[:m], ref_svals[:m], rtol=1e-7)
        assert O.subspace_angle(ref_modes[:, :7], modes.to_numpy()[:, :7], w) < 1e-6
    dim = 20011
    L, Rr = rng.standard_normal((dim, 6)), rng.standard_normal((6, 40))
    B = L @ Rr                                                    # rank 6, placed in 40 columns of a sparse dim x dim
    cols = rng.choice(dim, 40, replace=False)
    rr, cc = np.meshgrid(np.arange(dim), cols, indexing='ij')
    Mfull = sps.coo_matrix((B.ravel(), (rr.ravel(), cc.ravel())), shape=(dim, dim)).tocsr()
    sp = space_of(dim)
    op = CudaCsrOperator(Mfull, sp)
    Q = alg.RandomizedRangeFinder(op, power_iterations=1, error_estimator='loo').find_range(basis_size=10)
    Qh = Q.to_numpy()
    assert O.orthogonality_error(Qh) < 1e-10
    resid = L - Qh @ (Qh.T @ L)
    assert np.linalg.norm(resid) < 1e-8 * np.linalg.norm(L)


def test_next_row_empirical_interpolation_golden(loaded_lib):
    """N5 on the device: EI-greedy and DEIM (amax / dofs / per-column axpy / norm / lincomb kernels) reproduce
    the reference's interpolation DOFs exactly and its collateral bases to 1e-9."""
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    g = load_golden('ei')
    A = g['A']
    sp = space_of(A.shape[0])
    for tag, kw in (('l2', dict(rtol=1e-6)), ('sup', dict(error_norm='sup', rtol=1e-6)),
                    ('max8', dict(max_interpolation_dofs=8)), ('nodal', dict(rtol=1e-4, nodal_basis=True))):
        V = sp.from_numpy(A)
        dofs, basis, data = alg.ei_greedy(V, **kw)
        np.testing.assert_array_equal(dofs, g[f'ei_{tag}_dofs'])                      # index work: exact
        np.testing.assert_allclose(basis.to_numpy(), g[f'ei_{tag}_basis'], rtol=0, atol=1e-8)   # 30 elimination steps amplify rounding
        np.testing.assert_allclose(data['errors'], g[f'ei_{tag}_errors'], rtol=1e-7)
        np.testing.assert_allclose(data['interpolation_matrix'], g[f'ei_{tag}_matrix'], rtol=0, atol=1e-8)
        np.testing.assert_allclose(V.lincomb(data['coefficients']).to_numpy(), basis.to_numpy(), atol=1e-6)
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


# Fused Gram-Schmidt sweep (pymor_b200/csrc/mgs.cu); host logic and packet protocol are also CPU-checked
# (tests/test_host_layer.py, tests/test_mgs_sweep_model.py).
def test_mgs_sweep_parity(loaded_lib, rng):
    from pymor_b200 import algorithms as alg
    from pymor_b200.operators import CudaDiagonalOperator
    for dim in (1, 5, 4097, 100003, 3000001):
        k = 1 if dim == 1 else min(9, dim)
        sp = space_of(dim)
        Qh = np.linalg.qr(rng.standard_normal((dim, k)))[0]
        w = rng.uniform(0.5, 1.5, dim)
        prod = CudaDiagonalOperator(w, sp)
        for mode in ('euclid', 'w-stream', 'pre-weighted'):
            weighted = mode != 'euclid'
            for nq in sorted({0, min(1, k - 1), k - 1}):
                A = np.asfortranarray(np.hstack([Qh[:, :k - 1], rng.standard_normal((dim, 1))]))
                V = sp.from_numpy(A)
                Z = prod.apply(V).impl if mode == 'pre-weighted' else None       # z_j = w * q_j
                coeffs, nrm2 = V.impl.mgs_sweep(nq, k - 1, prod._w.data_ptr() if weighted else 0, weighted=Z)
                v = A[:, -1].copy()
                want = np.zeros(nq)
                for j in range(nq):
                    want[j] = np.sum(A[:, j] * (w * v if weighted else v))
                    v -= want[j] * A[:, j]
                scale = np.linalg.norm(A[:, -1]) * (1.5 if weighted else 1.0)
                np.testing.assert_allclose(coeffs, want, rtol=0, atol=1e-12 * scale)
                np.testing.assert_allclose(nrm2, np.sum(v * (w * v if weighted else v)), rtol=1e-12)
                np.testing.assert_allclose(V.to_numpy()[:, -1], v, rtol=0, atol=1e-13 * scale)
                np.testing.assert_array_equal(V.to_numpy()[:, :k - 1], A[:, :k - 1])
    g = load_golden('gram_schmidt')
    sp = space_of(g['A'].shape[0])
    for name, prod in (('euclid', None), ('diag', CudaDiagonalOperator(g['w'], sp))):
        Q, R = alg.gram_schmidt(sp.from_numpy(g['A']), product=prod, return_R=True, sweep=True)
        assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, None if prod is None else g['w'])
    n, k = 2_000_000, 24
    sp = space_of(n)
    A = sp.random_device(k, seed=3, scale=1.0 / np.sqrt(n))
    A.axpy(2.0, sp.random_device(1, seed=4, scale=1.0 / np.sqrt(n)))
    import torch
    prod = CudaDiagonalOperator(torch.rand(n, device='cuda', dtype=torch.float64) + 0.5, sp)
    launches = sp.ctx.launch_count
    Qs, Rs = alg.gram_schmidt(A, product=prod, return_R=True, sweep=True)
    fused_launches = sp.ctx.launch_count - launches
    Qp, Rp = alg.gram_schmidt(A, product=prod, return_R=True, sweep=False)
    np.testing.assert_allclose(Rs, Rp, rtol=0, atol=1e-12)
    assert np.max(np.abs(Qs.gramian(prod) - np.eye(k))) < 1e-13
    D = Qs.copy()
    D.axpy(-1.0, Qp)
    assert np.max(D.norm(prod)) < 1e-12
    assert fused_launches < 8 * k                      # a handful of launches per vector, not 2 per pair
    # classical variant with re-orthogonalisation: composed from the streaming inner / block lincomb / axpy kernels
    Qc, Rc = alg.gram_schmidt(A, product=prod, return_R=True, variant='cgs2')
    np.testing.assert_allclose(Rc, Rp, rtol=0, atol=1e-11)
    assert np.max(np.abs(Qc.gramian(prod) - np.eye(k))) < 1e-13
    D = Qc.copy()
    D.axpy(-1.0, Qp)
    assert np.max(D.norm(prod)) < 1e-11
    Qb, Rb = alg.block_gram_schmidt(A, product=prod, return_R=True, panel=8)      # BCGS2: DMMA Gramian / lincomb panels
    np.testing.assert_allclose(Rb, Rp, rtol=0, atol=1e-11)
    assert np.max(np.abs(Qb.gramian(prod) - np.eye(k))) < 1e-13
    D = Qb.copy()
    D.axpy(-1.0, Qp)
    assert np.max(D.norm(prod)) < 1e-11
    for name, gprod in (('euclid', None), ('diag', CudaDiagonalOperator(g['w'], space_of(g['A'].shape[0])))):
        sp_g = space_of(g['A'].shape[0]) if gprod is None else gprod.source
        Q, R = alg.gram_schmidt(sp_g.from_numpy(g['A']), product=gprod, return_R=True, variant='cgs2')
        assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, None if gprod is None else g['w'], tight=1e-11)
        Q, R = alg.block_gram_schmidt(sp_g.from_numpy(g['A']), product=gprod, return_R=True, panel=4)
        assert_gram_schmidt_golden(Q.to_numpy(), R, g, name, None if gprod is None else g['w'], tight=1e-11)


def test_device_cg_solver(loaded_lib, rng):
    """apply_inverse of an SPD CSR operator by block conjugate gradients on the device (Riesz representatives of
    the reductors' error estimators): 2-D 5-point Laplacian + mass shift, 40k DOFs, 6 right-hand sides."""
    import scipy.sparse.linalg as spla

    from pymor_b200.operators import CudaCsrOperator
    nx = 200
    T = sps.diags([-np.ones(nx - 1), 2.0 * np.ones(nx), -np.ones(nx - 1)], [-1, 0, 1])
    K = (sps.kron(sps.eye(nx), T) + sps.kron(T, sps.eye(nx)) + 0.1 * sps.eye(nx * nx)).tocsr()
    sp = space_of(nx * nx)
    op = CudaCsrOperator(K, sp)
    V = rng.standard_normal((nx * nx, 6))
    U, info = op.apply_inverse(sp.from_numpy(V), return_info=True)
    ref = spla.splu(K.tocsc()).solve(V)
    assert info['iterations'] < 2000
    np.testing.assert_allclose(U.to_numpy(), ref, atol=1e-8 * np.abs(ref).max())
