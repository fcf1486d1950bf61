"""Generate the golden fixtures in this directory by running the REFERENCE itself.

Run in the build container only (the reference is not present on the GPU box):

    PYTHONPATH=/root/reference/src python tests/golden/generate.py

Every fixture holds seeded inputs and the outputs of pyMOR's own NumpyVectorArray /
NumpyMatrixOperator / gram_schmidt / method_of_snapshots / qr_svd / pod / project on them.  The
fixtures pin the oracle (tests/test_oracle_pinned.py) and are compared with the CUDA path by the
`-m gpu` tests.  Nothing here is copied from the reference: it is only called.
"""
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sps

sys.path.insert(0, '/root/reference/src')

from pymor.algorithms.gram_schmidt import gram_schmidt  # noqa: E402
from pymor.algorithms.pod import pod  # noqa: E402
from pymor.algorithms.projection import project  # noqa: E402
from pymor.algorithms.svd_va import method_of_snapshots, qr_svd  # noqa: E402
from pymor.operators.numpy import NumpyMatrixOperator  # noqa: E402
from pymor.vectorarrays.numpy import NumpyVectorSpace  # noqa: E402

OUT = Path(__file__).resolve().parent


def mass_matrix(n):
    """1-D P1 mass matrix h/6*[1 4 1] (SPD, tridiagonal) in CSR."""
    h = 1.0 / (n + 1)
    return (sps.diags([np.full(n - 1, h / 6), np.full(n, 4 * h / 6), np.full(n - 1, h / 6)],
                      [-1, 0, 1]) * (n + 1)).tocsr()


def stiffness_matrix(n):
    return sps.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()


def vecops():
    rng = np.random.default_rng(1)
    dim, la, lb = 37, 7, 5
    A = np.asfortranarray(rng.standard_normal((dim, la)))
    B = np.asfortranarray(rng.standard_normal((dim, lb)))
    w = rng.uniform(0.5, 1.5, dim)
    M = mass_matrix(dim)
    C = rng.standard_normal((la, 4))
    alpha = rng.standard_normal(la)
    sp = NumpyVectorSpace(dim)
    VA, VB = sp.from_numpy(A.copy()), sp.from_numpy(B.copy())
    Wd = NumpyMatrixOperator(sps.diags(w).tocsr())
    Wm = NumpyMatrixOperator(M)
    out = dict(A=A, B=B, w=w, C=C, alpha=alpha,
               M_data=M.data, M_indices=M.indices, M_indptr=M.indptr)
    out['inner'] = VA.inner(VB)
    out['inner_diag'] = VA.inner(VB, Wd)
    out['inner_csr'] = VA.inner(VB, Wm)
    out['gramian'] = VA.gramian()
    out['gramian_diag'] = VA.gramian(Wd)
    out['gramian_csr'] = VA.gramian(Wm)
    out['pairwise'] = VA[:lb].pairwise_inner(VB)
    out['pairwise_diag'] = VA[:lb].pairwise_inner(VB, Wd)
    out['pairwise_csr'] = VA[:lb].pairwise_inner(VB, Wm)
    out['lincomb'] = VA.lincomb(C).to_numpy()
    out['norm'] = VA.norm()
    out['norm_diag'] = VA.norm(Wd)
    out['norm2'] = VA.norm2()
    out['norm2_csr'] = VA.norm2(Wm)
    Y = VA.copy(); Y.axpy(0.37, VA[::-1]); out['axpy_scalar_rev'] = Y.to_numpy()
    Y = VA.copy(); Y.axpy(alpha, VB[0]); out['axpy_vec_bcast'] = Y.to_numpy()
    Y = VA.copy(); Y[[5, 1, 3]].axpy(alpha[:3], VB[[0, 4, 2]]); out['axpy_list'] = Y.to_numpy()
    Y = VA.copy(); Y.scal(alpha); out['scal_vec'] = Y.to_numpy()
    Y = VA.copy(); Y[1:6:2].scal(-2.5); out['scal_slice'] = Y.to_numpy()
    out['dof_idx'] = np.array([0, 36, 5, 5, 17])
    out['dofs'] = VA[[6, 0, 2]].dofs(out['dof_idx'])
    ai, av = VA.amax(); out['amax_idx'] = ai; out['amax_val'] = av
    out['view_inner'] = VA[[6, 0, 0, 2]].inner(VB[3:0:-1])
    Z = VA.copy(); Z.append(VB[[1, 1, 4]]); del Z[[0, 2]]; out['append_del'] = Z.to_numpy()
    out['apply_csr'] = Wm.apply(VA).to_numpy()
    np.savez_compressed(OUT / 'vecops.npz', **out)


def gs():
    rng = np.random.default_rng(2)
    dim, k = 211, 14
    A = rng.standard_normal((dim, k))
    A[:, 4] = A[:, 1] + 1e-9 * rng.standard_normal(dim)       # forces a re-iteration
    A[:, 7] = 2 * A[:, 0] - A[:, 3]                           # linearly dependent -> removed
    A[:, 9] = 0.0                                             # zero vector -> removed
    A[:, 10:] += 3 * A[:, [0]]                                # strong overlap -> re-iterations
    A = np.asfortranarray(A)
    w = rng.uniform(0.5, 1.5, dim)
    M = mass_matrix(dim)
    sp = NumpyVectorSpace(dim)
    out = dict(A=A, w=w, M_data=M.data, M_indices=M.indices, M_indptr=M.indptr)
    for name, prod in (('euclid', None), ('diag', NumpyMatrixOperator(sps.diags(w).tocsr())),
                       ('csr', NumpyMatrixOperator(M))):
        Q, R = gram_schmidt(sp.from_numpy(A.copy()), product=prod, return_R=True)
        out[f'Q_{name}'] = Q.to_numpy()
        out[f'R_{name}'] = R
    # offset variant (reductors/basic.py:578-603 extend_basis)
    Q0 = gram_schmidt(sp.from_numpy(A[:, :4].copy()))
    Q0.append(sp.from_numpy(A[:, 10:].copy()))
    Q1 = gram_schmidt(Q0, offset=4)
    out['Q_offset'] = Q1.to_numpy()
    np.savez_compressed(OUT / 'gram_schmidt.npz', **out)


def pods():
    rng = np.random.default_rng(3)
    dim, k = 307, 24
    Gn = rng.standard_normal((dim, k)) / np.sqrt(dim)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    sig = 10.0 ** (-2 * np.arange(k) / (k - 1))
    A = np.asfortranarray(Gn @ np.diag(sig) @ Qk.T)
    w = rng.uniform(0.5, 1.5, dim)
    M = mass_matrix(dim)
    sp = NumpyVectorSpace(dim)
    out = dict(A=A, w=w, M_data=M.data, M_indices=M.indices, M_indptr=M.indptr)
    for name, prod in (('euclid', None), ('diag', NumpyMatrixOperator(sps.diags(w).tocsr())),
                       ('csr', NumpyMatrixOperator(M))):
        U, s, Vh = method_of_snapshots(sp.from_numpy(A.copy()), product=prod)
        out[f'mos_U_{name}'], out[f'mos_s_{name}'], out[f'mos_Vh_{name}'] = U.to_numpy(), s, Vh
        U, s, Vh = qr_svd(sp.from_numpy(A.copy()), product=prod)
        out[f'qr_U_{name}'], out[f'qr_s_{name}'] = U.to_numpy(), s
        U, s = pod(sp.from_numpy(A.copy()), product=prod, modes=10, method='method_of_snapshots')
        out[f'pod_U_{name}'], out[f'pod_s_{name}'] = U.to_numpy(), s
    # ill-conditioned input: method_of_snapshots loses orthogonality, pod() re-orthonormalises (pod.py:81-86)
    sig2 = 10.0 ** (-7 * np.arange(k) / (k - 1))
    A2 = np.asfortranarray(Gn @ np.diag(sig2) @ Qk.T)
    U, s = pod(sp.from_numpy(A2.copy()), rtol=1e-7, method='method_of_snapshots')
    out['A_illcond'], out['pod_U_illcond'], out['pod_s_illcond'] = A2, U.to_numpy(), s
    np.savez_compressed(OUT / 'pod.npz', **out)


def projection():
    rng = np.random.default_rng(4)
    dim, kv, ku = 151, 6, 9
    Mop = (stiffness_matrix(dim) + sps.random(dim, dim, density=0.03, random_state=5)).tocsr()
    V = np.asfortranarray(rng.standard_normal((dim, kv)))
    U = np.asfortranarray(rng.standard_normal((dim, ku)))
    Pm = mass_matrix(dim)
    sp = NumpyVectorSpace(dim)
    op = NumpyMatrixOperator(Mop)
    prod = NumpyMatrixOperator(Pm)
    VV, UU = sp.from_numpy(V.copy()), sp.from_numpy(U.copy())
    out = dict(V=V, U=U, Mop_data=Mop.data, Mop_indices=Mop.indices, Mop_indptr=Mop.indptr,
               P_data=Pm.data, P_indices=Pm.indices, P_indptr=Pm.indptr)
    out['proj_both'] = project(op, VV, UU).matrix
    out['proj_both_product'] = project(op, VV, UU, product=prod).matrix
    out['proj_galerkin'] = project(op, UU, UU).matrix
    out['proj_range_only'] = project(op, VV, None).apply(sp.from_numpy(U.copy())).to_numpy()
    out['proj_source_only'] = project(op, None, UU).as_range_array().to_numpy()
    out['apply2'] = op.apply2(VV, UU)
    out['pairwise_apply2'] = op.pairwise_apply2(VV, UU[:kv])
    np.savez_compressed(OUT / 'projection.npz', **out)


def thermalblock():
    """Config 1 in miniature: thermal block 2x2, 3^4 = 81 snapshots, POD w.r.t. the H1_0-semi product
    (src/pymordemos/thermalblock.py:362-393, `reduce_pod`), then Galerkin projection of the operator
    components (reductors/basic.py:177-186)."""
    from pymor.analyticalproblems.thermalblock import thermal_block_problem
    from pymor.core.logger import set_log_levels
    from pymor.discretizers.builtin import discretize_stationary_cg
    set_log_levels({'pymor': 'WARN'})
    problem = thermal_block_problem(num_blocks=(2, 2))
    fom, _ = discretize_stationary_cg(problem, diameter=1. / 12)
    training = problem.parameter_space.sample_uniformly(3)
    snaps = fom.solution_space.empty(reserve=len(training))
    for mu in training:
        snaps.append(fom.solve(mu))
    prod = fom.h1_0_semi_product
    P = prod.matrix.tocsr()
    P.sort_indices()
    out = dict(snapshots=snaps.to_numpy(), P_data=P.data, P_indices=P.indices, P_indptr=P.indptr)
    basis, svals = pod(snaps, modes=5, product=prod)                     # default method qr_svd
    out['pod_qr_U'], out['pod_qr_s'] = basis.to_numpy(), svals
    basis2, svals2 = pod(snaps, modes=5, product=prod, method='method_of_snapshots')
    out['pod_mos_U'], out['pod_mos_s'] = basis2.to_numpy(), svals2
    # projection of the affine components of the operator and of the rhs onto the POD basis
    comps = fom.operator.operators
    for i, c in enumerate(comps):
        m = c.matrix.tocsr()
        m.sort_indices()
        out[f'op{i}_data'], out[f'op{i}_indices'], out[f'op{i}_indptr'] = m.data, m.indices, m.indptr
        out[f'op{i}_proj'] = project(c, basis, basis).matrix
    out['n_ops'] = np.array(len(comps))
    out['rhs'] = fom.rhs.as_range_array().to_numpy()
    out['rhs_proj'] = project(fom.rhs, basis, None).matrix
    np.savez_compressed(OUT / 'thermalblock.npz', **out)


def chol_qr():
    """"Next" row N1: shifted CholeskyQR3 (src/pymor/algorithms/chol_qr.py:16)."""
    from pymor.algorithms.chol_qr import shifted_chol_qr
    rng = np.random.default_rng(6)
    dim, k = 257, 12
    Gn = rng.standard_normal((dim, k)) / np.sqrt(dim)
    Qk, _ = np.linalg.qr(rng.standard_normal((k, k)))
    out = {}
    # The ill-conditioned input needs the diagonal shift.  With the default kernel the reference
    # raises AttributeError there (BasicShiftedCholQRKernel reads `self.shift` before ever setting it,
    # chol_qr.py:227), so the shifted branch is pinned through recompute_shift=True.
    for tag, decades in (('mild', 3), ('ill', 9)):
        A = np.asfortranarray(Gn @ np.diag(10.0 ** (-decades * np.arange(k) / (k - 1))) @ Qk.T)
        w = rng.uniform(0.5, 1.5, dim)
        sp = NumpyVectorSpace(dim)
        out[f'A_{tag}'], out[f'w_{tag}'] = A, w
        rs = tag == 'ill'
        Q, R = shifted_chol_qr(sp.from_numpy(A.copy()), return_R=True, recompute_shift=rs)
        out[f'Q_euclid_{tag}'], out[f'R_euclid_{tag}'] = Q.to_numpy(), R
        Q, R = shifted_chol_qr(sp.from_numpy(A.copy()), product=NumpyMatrixOperator(sps.diags(w).tocsr()),
                               return_R=True, product_norm=float(np.sqrt(w.max())), recompute_shift=rs)
        out[f'Q_diag_{tag}'], out[f'R_diag_{tag}'] = Q.to_numpy(), R
    Q0 = shifted_chol_qr(sp.from_numpy(out['A_mild'][:, :5].copy()))
    Q0.append(sp.from_numpy(out['A_mild'][:, 5:].copy()))
    Q1, R1 = shifted_chol_qr(Q0, offset=5, return_R=True)
    out['Q_offset'], out['R_offset'] = Q1.to_numpy(), R1
    np.savez_compressed(OUT / 'chol_qr.npz', **out)


def hapod():
    """"Next" row N3: incremental HAPOD (src/pymor/algorithms/hapod.py:333)."""
    from pymor.algorithms.hapod import inc_vectorarray_hapod
    from pymor.core.logger import set_log_levels
    set_log_levels({'pymor': 'WARN'})
    rng = np.random.default_rng(7)
    dim, k = 193, 60
    A = np.asfortranarray(rng.standard_normal((dim, 7)) @ rng.standard_normal((7, k)) * np.sqrt(1.0 / dim)
                          + 1e-4 * rng.standard_normal((dim, k)))
    w = rng.uniform(0.5, 1.5, dim)
    sp = NumpyVectorSpace(dim)
    out = dict(A=A, w=w)
    for name, prod in (('euclid', None), ('diag', NumpyMatrixOperator(sps.diags(w).tocsr()))):
        for steps in (1, 4, 7):
            modes, svals, count = inc_vectorarray_hapod(steps, sp.from_numpy(A.copy()), 1e-3, 0.75, product=prod)
            out[f'modes_{name}_{steps}'], out[f'svals_{name}_{steps}'] = modes.to_numpy(), svals
            assert count == k
    from pymor.algorithms.hapod import dist_vectorarray_hapod
    for name, prod in (('euclid', None), ('diag', NumpyMatrixOperator(sps.diags(w).tocsr()))):
        for slices, arity in ((1, None), (5, None), (7, 2)):
            modes, svals, count = dist_vectorarray_hapod(slices, sp.from_numpy(A.copy()), 1e-3, 0.75, arity=arity,
                                                         product=prod)
            out[f'dist_modes_{name}_{slices}_{arity}'], out[f'dist_svals_{name}_{slices}_{arity}'] = modes.to_numpy(), svals
            assert count == k
    np.savez_compressed(OUT / 'hapod.npz', **out)


def ei():
    """"Next" row N5: empirical interpolation data, EI-greedy and DEIM (src/pymor/algorithms/ei.py:30, :182):
    amax / dofs / vector-alpha axpy / norm / lincomb on the backend, the small solves on the host."""
    from pymor.algorithms.ei import deim, ei_greedy
    from pymor.core.logger import set_log_levels
    set_log_levels({'pymor': 'WARN'})
    rng = np.random.default_rng(11)
    dim, k = 211, 30
    x = np.linspace(0, 1, dim)[:, None]
    mu = rng.uniform(0.1, 2.0, (1, k))
    A = np.asfortranarray(np.exp(-mu * x) * np.sin(3 * mu * x + 1) + 1e-3 * rng.standard_normal((dim, k)))
    sp = NumpyVectorSpace(dim)
    out = dict(A=A)
    for tag, kw in (('l2', dict(rtol=1e-6)), ('sup', dict(error_norm='sup', rtol=1e-6)),
                    ('max8', dict(max_interpolation_dofs=8)), ('nodal', dict(rtol=1e-4, nodal_basis=True))):
        dofs, basis, data = ei_greedy(sp.from_numpy(A.copy()), **kw)
        out[f'ei_{tag}_dofs'], out[f'ei_{tag}_basis'] = dofs, basis.to_numpy()
        out[f'ei_{tag}_errors'], out[f'ei_{tag}_coefficients'] = np.array(data['errors']), data['coefficients']
        out[f'ei_{tag}_matrix'] = data['interpolation_matrix']
    w = rng.uniform(0.5, 1.5, dim)
    out['w'] = w
    for tag, prod in (('euclid', None), ('diag', NumpyMatrixOperator(sps.diags(w).tocsr()))):
        dofs, basis, data = deim(sp.from_numpy(A.copy()), modes=9, product=prod)
        out[f'deim_{tag}_dofs'], out[f'deim_{tag}_basis'], out[f'deim_{tag}_svals'] = dofs, basis.to_numpy(), data['svals']
    dofs, basis, data = deim(sp.from_numpy(A[:, :6].copy()), pod=False)
    out['deim_nopod_dofs'], out['deim_nopod_basis'] = dofs, basis.to_numpy()
    np.savez_compressed(OUT / 'ei.npz', **out)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    vecops()
    gs()
    pods()
    projection()
    thermalblock()
    chol_qr()
    hapod()
    ei()
    for f in sorted(OUT.glob('*.npz')):
        print(f.name, f.stat().st_size)
