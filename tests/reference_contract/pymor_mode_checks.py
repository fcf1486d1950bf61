"""pyMOR's OWN algorithms on CudaVectorArray (collected only through run_contract.py, which puts the
reference on sys.path first so that the CUDA classes subclass pyMOR's real base classes).

Each check runs the unmodified reference function twice -- on NumpyVectorArray/NumpyMatrixOperator
and on CudaVectorArray/Cuda*Operator built from the same bits -- and compares.
"""
import numpy as np
import pytest
import scipy.sparse as sps
from pymor.algorithms.chol_qr import shifted_chol_qr
from pymor.algorithms.gram_schmidt import gram_schmidt
from pymor.algorithms.pod import pod
from pymor.algorithms.projection import project
from pymor.algorithms.svd_va import method_of_snapshots, qr_svd
from pymor.core.logger import set_log_levels
from pymor.operators.constructions import LincombOperator, VectorOperator
from pymor.operators.interface import Operator
from pymor.operators.numpy import NumpyMatrixOperator
from pymor.vectorarrays.interface import VectorArray, VectorSpace
from pymor.vectorarrays.numpy import NumpyVectorSpace

from pymor_b200.interface import HAVE_PYMOR, reference_root
from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator
from pymor_b200.vectorarray import CudaVectorArray, CudaVectorSpace

set_log_levels({'pymor': 'ERROR'})
DEVICE = 'cuda' if __import__('torch').cuda.is_available() else 'cpu'


def mass(n):
    h = 1.0 / (n + 1)
    return sps.diags([np.full(n - 1, h / 6), np.full(n, 4 * h / 6), np.full(n - 1, h / 6)], [-1, 0, 1]).tocsr()


def make(dim, k, seed, kind):
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.standard_normal((dim, k)) + 2 * rng.standard_normal((dim, 1)))
    nsp, csp = NumpyVectorSpace(dim), CudaVectorSpace(dim, device=DEVICE)
    if kind == 'euclid':
        return A, nsp, csp, None, None
    if kind == 'diag':
        w = rng.uniform(0.5, 1.5, dim)
        return A, nsp, csp, NumpyMatrixOperator(sps.diags(w).tocsr()), CudaDiagonalOperator(w, csp)
    M = mass(dim)
    return A, nsp, csp, NumpyMatrixOperator(M), CudaCsrOperator(M, csp)


def test_subclassing():
    assert HAVE_PYMOR
    assert issubclass(CudaVectorArray, VectorArray) and issubclass(CudaVectorSpace, VectorSpace)
    assert issubclass(CudaCsrOperator, Operator) and issubclass(CudaDiagonalOperator, Operator)


@pytest.mark.parametrize('kind', ['euclid', 'diag', 'csr'])
def test_gram_schmidt(kind):
    A, nsp, csp, nprod, cprod = make(157, 12, 1, kind)
    Qn, Rn = gram_schmidt(nsp.from_numpy(A.copy()), product=nprod, return_R=True)
    Qc, Rc = gram_schmidt(csp.from_numpy(A), product=cprod, return_R=True)
    np.testing.assert_allclose(Qc.to_numpy(), Qn.to_numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(Rc, Rn, rtol=0, atol=1e-11)


@pytest.mark.parametrize('kind', ['euclid', 'diag', 'csr'])
@pytest.mark.parametrize('method', ['method_of_snapshots', 'qr_svd'])
def test_pod(kind, method):
    A, nsp, csp, nprod, cprod = make(203, 9, 2, kind)
    Un, sn = pod(nsp.from_numpy(A.copy()), product=nprod, method=method)
    Uc, sc = pod(csp.from_numpy(A), product=cprod, method=method)
    np.testing.assert_allclose(sc, sn, rtol=1e-10)
    G = nsp.from_numpy(Uc.to_numpy()).inner(Un, nprod)
    np.testing.assert_allclose(np.abs(G), np.eye(len(Un)), atol=1e-8)
    svd = {'method_of_snapshots': method_of_snapshots, 'qr_svd': qr_svd}[method]
    U, s, Vh = svd(csp.from_numpy(A), product=cprod)
    np.testing.assert_allclose(U.lincomb(np.diag(s) @ Vh).to_numpy(), A, atol=1e-9)


def test_project_and_lincomb_operator():
    A, nsp, csp, nprod, cprod = make(120, 7, 3, 'csr')
    rng = np.random.default_rng(5)
    M1 = (sps.random(120, 120, density=0.05, random_state=1) + sps.eye(120)).tocsr()
    M2 = mass(120)
    V = np.asfortranarray(rng.standard_normal((120, 5)))
    for ops_n, ops_c in [((NumpyMatrixOperator(M1),), (CudaCsrOperator(M1, csp),)),
                         ((NumpyMatrixOperator(M1), NumpyMatrixOperator(M2)),
                          (CudaCsrOperator(M1, csp), CudaCsrOperator(M2, csp)))]:
        if len(ops_n) == 1:
            on, oc = ops_n[0], ops_c[0]
        else:
            on, oc = LincombOperator(ops_n, [1.5, -0.5]), LincombOperator(ops_c, [1.5, -0.5])
        pn = project(on, nsp.from_numpy(V.copy()), nsp.from_numpy(A.copy()), product=nprod)
        pc = project(oc, csp.from_numpy(V), csp.from_numpy(A), product=cprod)
        mn = pn.matrix if hasattr(pn, 'matrix') else pn.assemble().matrix
        mc = pc.matrix if hasattr(pc, 'matrix') else pc.assemble().matrix
        np.testing.assert_allclose(mc, mn, rtol=1e-11, atol=1e-12)
    # right-hand side functional: project(VectorOperator(f), RB, None)
    f = rng.standard_normal((120, 1))
    pn = project(VectorOperator(nsp.from_numpy(f.copy())), nsp.from_numpy(A.copy()), None)
    pc = project(VectorOperator(csp.from_numpy(f)), csp.from_numpy(A), None)
    np.testing.assert_allclose(pc.matrix, pn.matrix, rtol=1e-11, atol=1e-12)


def test_shifted_chol_qr_next_row_N1():
    A, nsp, csp, nprod, cprod = make(300, 10, 4, 'diag')
    Qn, Rn = shifted_chol_qr(nsp.from_numpy(A.copy()), product=nprod, return_R=True)
    Qc, Rc = shifted_chol_qr(csp.from_numpy(A), product=cprod, return_R=True)
    np.testing.assert_allclose(Rc, Rn, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(Qc.to_numpy(), Qn.to_numpy(), atol=1e-9)


def test_hapod_next_row_N3():
    """pyMOR's incremental and distributed HAPOD: every tree node is a `pod` call on the backend."""
    from pymor.algorithms.hapod import dist_vectorarray_hapod, inc_vectorarray_hapod
    A, nsp, csp, nprod, cprod = make(180, 40, 7, 'diag')
    rng = np.random.default_rng(8)
    A = np.asfortranarray((rng.standard_normal((180, 6)) @ rng.standard_normal((6, 40))) + 1e-3 * A)
    for fn, arg in ((inc_vectorarray_hapod, 4), (dist_vectorarray_hapod, 4)):
        mn, sn, cn = fn(arg, nsp.from_numpy(A.copy()), 1e-2, 0.75, product=nprod)
        mc, sc, cc = fn(arg, csp.from_numpy(A), 1e-2, 0.75, product=cprod)
        assert cn == cc and len(mn) == len(mc)
        np.testing.assert_allclose(sc, sn, rtol=1e-8)
        G = nsp.from_numpy(mc.to_numpy()).inner(mn, nprod)
        np.testing.assert_allclose(np.abs(np.linalg.svd(G, compute_uv=False)), 1.0, atol=1e-7)


def test_randomized_range_finder_next_row_N2():
    """pyMOR's RandomizedRangeFinder / randomized_svd with a CUDA CSR operator as `A`: exercises
    Operator.apply / apply_adjoint, space.random (pyMOR's global RNG -> identical samples) and GS."""
    from pymor.algorithms.rand_la import RandomizedRangeFinder, randomized_svd
    from pymor.tools.random import new_rng
    dim = 140
    rng = np.random.default_rng(9)
    L, Rr = rng.standard_normal((dim, 5)), rng.standard_normal((5, dim))
    M = sps.csr_matrix(L @ Rr) + 1e-6 * sps.eye(dim)
    nsp, csp = NumpyVectorSpace(dim), CudaVectorSpace(dim, device=DEVICE)
    on, oc = NumpyMatrixOperator(M), CudaCsrOperator(M, csp)
    with new_rng(3):
        Qn = RandomizedRangeFinder(on, power_iterations=1).find_range(basis_size=8)
    with new_rng(3):
        Qc = RandomizedRangeFinder(oc, power_iterations=1).find_range(basis_size=8)
    # the 5 dominant directions agree tightly; directions 6-8 are resolved out of the 1e-6 perturbation after one
    # power iteration, i.e. out of components of relative size 1e-12: they are conditioned like 1e12 * eps w.r.t. any
    # change of rounding between the two backends and are compared through what they are for -- the range they add
    np.testing.assert_allclose(Qc.to_numpy()[:, :5], Qn.to_numpy()[:, :5], atol=1e-9)
    assert np.max(np.abs(Qc.gramian() - np.eye(8))) < 1e-12
    Md0 = M.toarray()
    res_c = np.linalg.norm(Md0 - Qc.to_numpy() @ (Qc.to_numpy().T @ Md0), 2)
    res_n = np.linalg.norm(Md0 - Qn.to_numpy() @ (Qn.to_numpy().T @ Md0), 2)
    assert res_c <= 1.01 * res_n + 1e-12 and res_c < 2e-6
    with new_rng(4):
        Un, sn, Vn = randomized_svd(on, 5)
    with new_rng(4):
        Uc, sc, Vc = randomized_svd(oc, 5)
    np.testing.assert_allclose(sc, sn, rtol=1e-8)
    np.testing.assert_allclose(np.abs(nsp.from_numpy(Uc.to_numpy()).inner(Un)), np.eye(5), atol=1e-6)
    # ADAPTIVE mode (rand_la.py:20-252 with tol): the basis grows block by block until the error estimator -- the
    # 'bs18' test-vector bound and the leave-one-out estimator 'loo' -- drops below tol; same samples (pyMOR's
    # global RNG) => same basis sizes, same estimates, same range on both backends
    Md = M.toarray()
    for est in ('bs18', 'loo'):
        with new_rng(5):
            fn = RandomizedRangeFinder(on, power_iterations=1, error_estimator=est, block_size=2)
            Bn = fn.find_range(tol=1e-3)
        with new_rng(5):
            fc = RandomizedRangeFinder(oc, power_iterations=1, error_estimator=est, block_size=2)
            Bc = fc.find_range(tol=1e-3)
        assert len(Bc) == len(Bn) and 5 <= len(Bc) <= 12
        np.testing.assert_allclose(fc.last_estimated_error, fn.last_estimated_error, rtol=1e-3, atol=1e-12)
        Bh = Bc.to_numpy()
        assert np.linalg.norm(Md - Bh @ (Bh.T @ Md), 2) < 1e-3
        np.testing.assert_allclose(Bh[:, :5], Bn.to_numpy()[:, :5], atol=1e-9)
    with new_rng(6):
        Un, sn, Vn = randomized_svd(on, 5, rtol=1e-4)           # adaptive range finder inside RandomizedSVD
    with new_rng(6):
        Uc, sc, Vc = randomized_svd(oc, 5, rtol=1e-4)
    np.testing.assert_allclose(sc, sn, rtol=1e-8)


def test_reductor_workflow_next_row_N4():
    """Config 1 as a workflow, with pyMOR's own reductor: thermal-block FOM solved on the CPU, snapshots
    moved to the device, POD there, `StationaryRBReductor` projecting DEVICE operators
    (reductors/basic.py:152-200 -> project -> Cuda*Operator.apply2 / apply_adjoint), ROM solve on the host,
    `reconstruct` = device lincomb.  Must reproduce the all-NumPy pipeline."""
    from pymor.analyticalproblems.thermalblock import thermal_block_problem
    from pymor.discretizers.builtin import discretize_stationary_cg
    from pymor.models.basic import StationaryModel
    from pymor.reductors.basic import StationaryRBReductor
    problem = thermal_block_problem(num_blocks=(2, 2))
    fom, _ = discretize_stationary_cg(problem, diameter=1. / 10)
    mus = problem.parameter_space.sample_uniformly(2)
    snaps = fom.solution_space.empty()
    for mu in mus:
        snaps.append(fom.solve(mu))
    dim = fom.solution_space.dim
    csp = CudaVectorSpace(dim, device=DEVICE)
    # the same FOM with every operator living on the device space
    dev_ops = [CudaCsrOperator(op.matrix, csp) for op in fom.operator.operators]
    dev_operator = LincombOperator(dev_ops, fom.operator.coefficients)
    dev_rhs = VectorOperator(csp.from_numpy(fom.rhs.as_range_array().to_numpy()))
    dev_product = CudaCsrOperator(fom.h1_0_semi_product.matrix, csp)
    dev_fom = StationaryModel(dev_operator, dev_rhs, products={'h1_0_semi': dev_product})
    # POD on both sides
    rb_n, sv_n = pod(snaps, modes=6, product=fom.h1_0_semi_product)
    rb_c, sv_c = pod(csp.from_numpy(snaps.to_numpy()), modes=6, product=dev_product)
    np.testing.assert_allclose(sv_c, sv_n, rtol=1e-10)
    rom_n = StationaryRBReductor(fom, rb_n, product=fom.h1_0_semi_product).reduce()
    red_c = StationaryRBReductor(dev_fom, rb_c, product=dev_product)
    rom_c = red_c.reduce()
    mu = problem.parameter_space.sample_randomly(1)[0]
    u_n, u_c = rom_n.solve(mu), rom_c.solve(mu)
    U_fom = fom.solve(mu)
    U_rec = red_c.reconstruct(u_c)
    assert isinstance(U_rec, CudaVectorArray)
    err = np.linalg.norm(U_rec.to_numpy() - U_fom.to_numpy()) / np.linalg.norm(U_fom.to_numpy())
    ref_err = np.linalg.norm(rb_n.lincomb(u_n.to_numpy()).to_numpy() - U_fom.to_numpy()) / np.linalg.norm(U_fom.to_numpy())
    assert err < 5e-2 and abs(err - ref_err) < 1e-8        # same ROM error as the all-NumPy pipeline
    # reduced operator matrices agree up to the sign/rotation freedom of the bases: compare outputs instead
    np.testing.assert_allclose(U_rec.to_numpy(), rb_n.lincomb(u_n.to_numpy()).to_numpy(), atol=1e-8)
    # a17: ProjectionBasedReductor.extend_basis (reductors/basic.py:125-149, :578-603) on the device basis --
    # append + gram_schmidt(offset=...) resp. POD of the projection error + the orthonormality check
    red_n = StationaryRBReductor(fom, rb_n, product=fom.h1_0_semi_product)
    for method in ('gram_schmidt', 'pod'):
        mu_new = problem.parameter_space.sample_randomly(1)[0]
        U_new = fom.solve(mu_new)
        red_n.extend_basis(U_new, method=method)
        red_c.extend_basis(csp.from_numpy(U_new.to_numpy()), method=method)
        assert len(red_c.bases['RB']) == len(red_n.bases['RB']) and isinstance(red_c.bases['RB'], CudaVectorArray)
        Bc, Bn = red_c.bases['RB'].to_numpy(), red_n.bases['RB'].to_numpy()
        np.testing.assert_allclose(np.abs(np.sum(Bc[:, -1] * Bn[:, -1])) / np.sum(Bn[:, -1] ** 2), 1.0, atol=1e-6)
        u_c2 = red_c.reduce().solve(mu_new)
        rec = red_c.reconstruct(u_c2).to_numpy()
        assert np.linalg.norm(rec - U_new.to_numpy()) / np.linalg.norm(U_new.to_numpy()) < 1e-6   # snapshot is in the basis


def test_empirical_interpolation_next_row_N5():
    """pyMOR's own ei_greedy / deim on device arrays: same interpolation DOFs, same collateral bases."""
    from pymor.algorithms.ei import deim, ei_greedy
    rng = np.random.default_rng(11)
    dim, k = 211, 30
    x = np.linspace(0, 1, dim)[:, None]
    mu = rng.uniform(0.1, 2.0, (1, k))
    A = np.asfortranarray(np.exp(-mu * x) * np.sin(3 * mu * x + 1) + 1e-3 * rng.standard_normal((dim, k)))
    nsp, csp = NumpyVectorSpace(dim), CudaVectorSpace(dim, device=DEVICE)
    for kw in (dict(rtol=1e-6), dict(error_norm='sup', rtol=1e-6), dict(rtol=1e-4, nodal_basis=True)):
        dn, bn, datan = ei_greedy(nsp.from_numpy(A.copy()), **kw)
        dc, bc, datac = ei_greedy(csp.from_numpy(A), **kw)
        assert isinstance(bc, CudaVectorArray)
        np.testing.assert_array_equal(dn, dc)
        np.testing.assert_allclose(bc.to_numpy(), bn.to_numpy(), atol=1e-9)
        np.testing.assert_allclose(datac['interpolation_matrix'], datan['interpolation_matrix'], atol=1e-9)
    w = rng.uniform(0.5, 1.5, dim)
    dn, bn, datan = deim(nsp.from_numpy(A.copy()), modes=9, product=NumpyMatrixOperator(sps.diags(w).tocsr()))
    dc, bc, datac = deim(csp.from_numpy(A), modes=9, product=CudaDiagonalOperator(w, csp))
    np.testing.assert_array_equal(dn, dc)
    np.testing.assert_allclose(datac['svals'], datan['svals'], rtol=1e-10)


def test_reference_stored_golden_vector_thermalblock_pod():
    """The reference's OWN golden vector for this path: pymortests/testdata/check_results/
    test_thermalblock_results/c8a460e4... = results of `pymordemos/thermalblock.py --alg=pod 2 2 3 5`
    (test_thermalblock_results, pymortests/demos.py:386-402): 81 FEM snapshots (20 201 DOFs), POD with the
    H1_0-semi product, 5 modes, coercive RB reductor, errors of the ROM at 10 random parameters.  The demo
    module itself needs cyclopts/matplotlib (absent), so its `reduce_pod` pipeline (thermalblock.py:362-393) is
    re-driven here; with pyMOR's NumPy backend that reproduces the stored numbers to 1e-12, which validates the
    re-drive.  Then the POD is swapped for (a) the CPU oracle and (b) pyMOR's `pod` on device arrays with the
    CUDA product operator: the downstream ROM errors must still match the stored golden vector."""
    import pickle

    from oracle import numpy_path as O
    from pymor.algorithms.error import reduction_error_analysis
    from pymor.analyticalproblems.thermalblock import thermal_block_problem
    from pymor.discretizers.builtin import discretize_stationary_cg
    from pymor.parameters.functionals import ExpressionParameterFunctional
    from pymor.reductors.coercive import CoerciveRBReductor
    from pymor.tools.random import new_rng
    golden_file = (f'{reference_root()}/src/pymortests/testdata/check_results/test_thermalblock_results/'
                   'c8a460e412d9170f12508ab3608fd7e1e1750603')
    with open(golden_file, 'rb') as f:
        assert f.readline().decode().strip() == "['--alg=pod', 2, 2, 3, 5]"
        gold = pickle.load(f)

    problem = thermal_block_problem(num_blocks=(2, 2))
    fom, _ = discretize_stationary_cg(problem, diameter=1. / 100)
    space = fom.parameters.space(0.1, 1.)
    coercivity = ExpressionParameterFunctional('min(diffusion)', fom.parameters)
    snaps = fom.operator.source.empty()
    for mu in space.sample_uniformly(3):
        snaps.append(fom.solve(mu))
    prod = fom.h1_0_semi_product
    P = prod.matrix.tocsr()
    csp = CudaVectorSpace(fom.solution_space.dim, device=DEVICE)

    def pod_numpy():
        return pod(snaps, modes=5, product=prod)[0]

    def pod_oracle():
        U, _ = O.pod(snaps.to_numpy(), P, modes=5, method='qr_svd')
        return fom.solution_space.from_numpy(U)

    def pod_device():
        U, _ = pod(csp.from_numpy(snaps.to_numpy()), modes=5, product=CudaCsrOperator(P, csp))
        assert isinstance(U, CudaVectorArray)
        return fom.solution_space.from_numpy(U.to_numpy())

    for name, make_basis, tol in (('numpy', pod_numpy, 1e-11), ('oracle', pod_oracle, 1e-8), ('device', pod_device, 1e-8)):
        red = CoerciveRBReductor(fom, product=prod, coercivity_estimator=coercivity, check_orthonormality=False)
        red.extend_basis(make_basis(), method='trivial')
        rom = red.reduce()
        with new_rng(42):                     # the test parameters of the stored run (pyMOR's default seed)
            res = reduction_error_analysis(rom, fom=fom, reductor=red, error_estimator=True,
                                           error_norms=(fom.h1_0_semi_norm, fom.l2_norm), condition=True,
                                           test_mus=space.sample_randomly(10), basis_sizes=1)
        assert list(np.asarray(res['basis_sizes'])) == list(gold['basis_sizes']), name
        np.testing.assert_allclose(res['norms'], gold['norms'], rtol=1e-11, err_msg=name)
        for key in ('errors', 'max_errors', 'rel_errors', 'max_rel_errors', 'error_estimates'):
            np.testing.assert_allclose(res[key], gold[key], rtol=tol, err_msg=f'{name}: {key}')


def test_reference_stored_golden_vector_burgers_ei_greedy():
    """The reference's stored golden vector for the empirical-interpolation row (N5):
    pymortests/testdata/check_results/test_burgers_ei_results/* = `pymordemos/burgers_ei.py 1 2 2 5 2 5
    --grid=20` (test_burgers_ei_results, pymortests/demos.py:405-413): EI-greedy errors and triangularity errors
    for 202 operator evaluations of the FV Burgers model with the L2-product norm, 5 interpolation DOFs.  The
    demo module needs cyclopts/matplotlib, so `interpolate_operators`' evaluation loop (algorithms/ei.py:404-417)
    is re-driven; with pyMOR's NumPy backend that reproduces the stored numbers, then EI-greedy is swapped for
    the CPU oracle and for pyMOR's `ei_greedy` on device arrays with the CUDA diagonal product."""
    import glob
    import math
    import pickle

    from oracle import numpy_path as O
    from pymor.algorithms.ei import ei_greedy
    from pymor.analyticalproblems.burgers import burgers_problem_2d
    from pymor.discretizers.builtin import RectGrid, discretize_instationary_fv
    gold = None
    for fn in glob.glob(f'{reference_root()}/src/pymortests/testdata/check_results/test_burgers_ei_results/*'):
        with open(fn, 'rb') as f:
            if f.readline().decode().strip() == "['1', '2', '2', '5', '2', '5', '--grid=20']":
                gold = pickle.load(f)
    assert gold is not None
    problem = burgers_problem_2d(vx=1., vy=1., initial_data_type='sin', parameter_range=(1, 2), torus=True)
    fom, _ = discretize_instationary_fv(problem, diameter=1. / (20 / math.sqrt(2)), grid_type=RectGrid,
                                        num_flux='engquist_osher', lxf_lambda=1., nt=100)
    ev = fom.operator.range.empty()
    for mu in problem.parameter_space.sample_uniformly(2):
        ev.append(fom.operator.apply(fom.solve(mu), mu=mu))
    w = np.asarray(fom.l2_product.matrix.diagonal(), dtype=float)
    assert (fom.l2_product.matrix - sps.diags(w)).tocsr().nnz == 0       # the FV L2 product is diagonal
    E = ev.to_numpy()
    csp = CudaVectorSpace(ev.dim, device=DEVICE)
    cprod = CudaDiagonalOperator(w, csp)
    runs = {
        'numpy': lambda: ei_greedy(ev.copy(), error_norm=fom.l2_norm, max_interpolation_dofs=5)[2],
        'oracle': lambda: O.ei_greedy(E, error_norm=lambda M: O.norm(M, w), max_interpolation_dofs=5)[2],
        'device': lambda: ei_greedy(csp.from_numpy(E), error_norm=lambda U: U.norm(cprod), max_interpolation_dofs=5)[2],
    }
    for name, run in runs.items():
        data = run()
        np.testing.assert_allclose(data['errors'], gold['errors'], rtol=1e-12 if name == 'numpy' else 1e-9, err_msg=name)
        np.testing.assert_allclose(data['triangularity_errors'], gold['triangularity_errors'], rtol=0, atol=1e-14,
                                   err_msg=name)


@pytest.mark.parametrize('kind', ['euclid', 'diag', 'csr'])
def test_block_gram_schmidt_variant_vs_pymor_mgs(kind):
    """`pymor_b200.algorithms.gram_schmidt(variant='cgs2')` is written against the VectorArray interface only: on
    pyMOR's own NumpyVectorArray and on CudaVectorArray it reproduces pyMOR's (modified) gram_schmidt -- the QR
    factorisation is unique."""
    from pymor_b200 import algorithms as alg
    A, nsp, csp, nprod, cprod = make(157, 12, 1, kind)
    Qn, Rn = gram_schmidt(nsp.from_numpy(A.copy()), product=nprod, return_R=True)
    Q1, R1 = alg.gram_schmidt(nsp.from_numpy(A.copy()), product=nprod, return_R=True, variant='cgs2')
    Q2, R2 = alg.gram_schmidt(csp.from_numpy(A), product=cprod, return_R=True, variant='cgs2')
    Q3, R3 = alg.block_gram_schmidt(nsp.from_numpy(A.copy()), product=nprod, return_R=True, panel=5)
    Q4, R4 = alg.block_gram_schmidt(csp.from_numpy(A), product=cprod, return_R=True, panel=5)
    for Q, R in ((Q1, R1), (Q2, R2), (Q3, R3), (Q4, R4)):
        np.testing.assert_allclose(Q.to_numpy(), Qn.to_numpy(), rtol=0, atol=1e-11)
        np.testing.assert_allclose(R, Rn, rtol=0, atol=1e-11)


def test_numpy_conversion_operator_bridges_cpu_and_device():
    """N4 bridging: pyMOR's own `NumpyConversionOperator` (operators/constructions.py:1384) works on
    `CudaVectorSpace` as it is (`to_numpy` / `from_numpy`), so a CPU model can wrap a device operator --
    `to_numpy o CudaCsrOperator o from_numpy` acts like the NumpyMatrixOperator of the same matrix -- and CPU
    solves can feed device snapshot arrays."""
    from pymor.operators.constructions import ConcatenationOperator, NumpyConversionOperator
    dim = 97
    M = (sps.random(dim, dim, density=0.08, random_state=4) + 2 * sps.eye(dim)).tocsr()
    nsp, csp = NumpyVectorSpace(dim), CudaVectorSpace(dim, device=DEVICE)
    up, down = NumpyConversionOperator(csp, 'from_numpy'), NumpyConversionOperator(csp, 'to_numpy')
    assert up.source == nsp and up.range == csp and down.source == csp and down.range == nsp
    wrapped = ConcatenationOperator([down, CudaCsrOperator(M, csp), up])
    ref = NumpyMatrixOperator(M)
    rng = np.random.default_rng(2)
    U, V = nsp.from_numpy(rng.standard_normal((dim, 4))), nsp.from_numpy(rng.standard_normal((dim, 3)))
    np.testing.assert_allclose(wrapped.apply(U).to_numpy(), ref.apply(U).to_numpy(), atol=1e-12)
    np.testing.assert_allclose(wrapped.apply_adjoint(V).to_numpy(), ref.apply_adjoint(V).to_numpy(), atol=1e-12)
    np.testing.assert_allclose(wrapped.apply2(V, U), ref.apply2(V, U), atol=1e-12)
    snaps = csp.empty()
    for i in range(len(U)):                                   # CPU "solves" appended to a device snapshot array
        snaps.append(up.apply(U[i]))
    assert isinstance(snaps, CudaVectorArray) and len(snaps) == 4
    np.testing.assert_array_equal(snaps.to_numpy(), U.to_numpy())


def test_device_solvers_and_coercive_reductor_error_estimator_N4():
    """`apply_inverse` on the CUDA operators (device CG / reciprocal weights through pyMOR's Solver interface) and,
    built on it, pyMOR's `CoerciveRBReductor` with its residual-based error estimator on a DEVICE model: the Riesz
    representatives `product.apply_inverse(residual)` (reductors/residual.py:151, algorithms/image.py:106) and
    their Gram-Schmidt run on the device; estimates and errors agree with the all-NumPy pipeline."""
    from pymor.analyticalproblems.thermalblock import thermal_block_problem
    from pymor.discretizers.builtin import discretize_stationary_cg
    from pymor.models.basic import StationaryModel
    from pymor.parameters.functionals import ExpressionParameterFunctional
    from pymor.reductors.coercive import CoerciveRBReductor

    from pymor_b200.solvers import CgSolver, cg
    # solvers on their own
    dim = 200
    K = sps.diags([-np.ones(dim - 1), 2.2 * np.ones(dim), -np.ones(dim - 1)], [-1, 0, 1]).tocsr()
    csp, nsp = CudaVectorSpace(dim, device=DEVICE), NumpyVectorSpace(dim)
    op = CudaCsrOperator(K, csp)
    assert op.solver is None and isinstance(op.default_solver(), CgSolver)
    rng = np.random.default_rng(0)
    Vh = rng.standard_normal((dim, 3))
    V = csp.from_numpy(Vh)
    assert np.abs(K @ op.apply_inverse(V).to_numpy() - Vh).max() < 1e-8
    assert np.abs(K @ op.apply_inverse_adjoint(V).to_numpy() - Vh).max() < 1e-8
    assert np.abs(K @ op.with_(solver=CgSolver(tol=1e-13)).apply_inverse(V).to_numpy() - Vh).max() < 1e-11
    X, _ = cg(NumpyMatrixOperator(K).apply, nsp.from_numpy(Vh), tol=1e-13)      # interface-only: NumPy classes too
    assert np.abs(K @ X.to_numpy() - Vh).max() < 1e-11
    D = CudaDiagonalOperator(rng.uniform(0.5, 1.5, dim), csp)
    assert np.abs(D.apply(D.apply_inverse(V)).to_numpy() - Vh).max() < 1e-14
    assert CudaCsrOperator(K + sps.diags(np.ones(dim - 1), 1), csp).default_solver() is None   # pyMOR's default

    # the reductor with error estimator
    problem = thermal_block_problem(num_blocks=(2, 2))
    fom, _ = discretize_stationary_cg(problem, diameter=1. / 10)
    coercivity = ExpressionParameterFunctional('min(diffusion)', fom.parameters)
    space = fom.parameters.space(0.1, 1.)
    snaps = fom.solution_space.empty()
    for mu in space.sample_uniformly(2):
        snaps.append(fom.solve(mu))
    csp = CudaVectorSpace(fom.solution_space.dim, device=DEVICE)
    tight = CgSolver(tol=1e-13)
    dev_ops = [CudaCsrOperator(o.matrix, csp) for o in fom.operator.operators]
    dev_fom = StationaryModel(LincombOperator(dev_ops, fom.operator.coefficients),
                              VectorOperator(csp.from_numpy(fom.rhs.as_range_array().to_numpy())))
    dev_product = CudaCsrOperator(fom.h1_0_semi_product.matrix, csp, solver=tight)
    rb_n, _ = pod(snaps, modes=5, product=fom.h1_0_semi_product)
    rb_c, _ = pod(csp.from_numpy(snaps.to_numpy()), modes=5, product=dev_product)
    rom_n = CoerciveRBReductor(fom, rb_n, product=fom.h1_0_semi_product, coercivity_estimator=coercivity).reduce()
    red_c = CoerciveRBReductor(dev_fom, rb_c, product=dev_product, coercivity_estimator=coercivity)
    rom_c = red_c.reduce()
    assert isinstance(red_c.bases['RB'], CudaVectorArray)
    for mu in space.sample_randomly(3):
        u_n, u_c = rom_n.solve(mu), rom_c.solve(mu)
        est_n, est_c = rom_n.estimate_error(mu), rom_c.estimate_error(mu)
        np.testing.assert_allclose(est_c, est_n, rtol=1e-6)
        U_fom = fom.solve(mu)
        err_c = (red_c.reconstruct(u_c).to_numpy() - U_fom.to_numpy())
        err_n = (rb_n.lincomb(u_n.to_numpy()).to_numpy() - U_fom.to_numpy())
        np.testing.assert_allclose(np.linalg.norm(err_c), np.linalg.norm(err_n), rtol=1e-6)


def test_device_model_solve_and_rb_greedy_N4():
    """A thermal-block model whose operators live on the device: `LincombOperator.assemble(mu)` folds the affine
    components into one `CudaCsrOperator` (`_assemble_lincomb`), `fom.solve(mu)` is one device CG solve, and pyMOR's
    own weak greedy (`rb_greedy` with `CoerciveRBReductor`, Gram-Schmidt basis extension, error estimator) runs with
    every vector operation on the backend.  Same selected parameters and errors as the all-NumPy run."""
    from pymor.algorithms.greedy import rb_greedy
    from pymor.analyticalproblems.thermalblock import thermal_block_problem
    from pymor.discretizers.builtin import discretize_stationary_cg
    from pymor.models.basic import StationaryModel
    from pymor.parameters.functionals import ExpressionParameterFunctional
    from pymor.reductors.coercive import CoerciveRBReductor

    from pymor_b200.solvers import CgSolver
    problem = thermal_block_problem(num_blocks=(2, 2))
    fom, _ = discretize_stationary_cg(problem, diameter=1. / 10)
    coercivity = ExpressionParameterFunctional('min(diffusion)', fom.parameters)
    space = fom.parameters.space(0.1, 1.)
    csp = CudaVectorSpace(fom.solution_space.dim, device=DEVICE)
    tight = CgSolver(tol=1e-13)
    dev_ops = [CudaCsrOperator(o.matrix, csp, solver=tight) for o in fom.operator.operators]
    dev_operator = LincombOperator(dev_ops, fom.operator.coefficients)
    dev_product = CudaCsrOperator(fom.h1_0_semi_product.matrix, csp, solver=tight)
    dev_fom = StationaryModel(dev_operator, VectorOperator(csp.from_numpy(fom.rhs.as_range_array().to_numpy())),
                              products={'h1_0_semi': dev_product})
    mu = space.sample_randomly(1)[0]
    assembled = dev_operator.assemble(mu)
    assert isinstance(assembled, CudaCsrOperator)
    U_dev, U_ref = dev_fom.solve(mu), fom.solve(mu)
    assert isinstance(U_dev, CudaVectorArray)
    np.testing.assert_allclose(U_dev.to_numpy(), U_ref.to_numpy(), atol=1e-9 * np.abs(U_ref.to_numpy()).max())

    training = space.sample_uniformly(2)
    red_n = CoerciveRBReductor(fom, product=fom.h1_0_semi_product, coercivity_estimator=coercivity)
    red_c = CoerciveRBReductor(dev_fom, product=dev_product, coercivity_estimator=coercivity)
    g_n = rb_greedy(fom, red_n, training, max_extensions=4)
    g_c = rb_greedy(dev_fom, red_c, training, max_extensions=4)
    # (the 2x2 thermal block is symmetric: parameters that are mirror images of each other tie, so the SELECTED
    # parameters may differ by a symmetry; the error decay does not)
    assert sorted(sorted(m.to_numpy().tolist()) for m in g_c['max_err_mus']) == \
        sorted(sorted(m.to_numpy().tolist()) for m in g_n['max_err_mus'])
    np.testing.assert_allclose(g_c['max_errs'], g_n['max_errs'], rtol=1e-5)
    assert isinstance(red_c.bases['RB'], CudaVectorArray) and len(red_c.bases['RB']) == len(red_n.bases['RB']) == 4


@pytest.mark.parametrize('kind', ['euclid', 'diag'])
def test_gram_schmidt_biorth_runs_unchanged(kind):
    """The biorthogonal variant in the same reference file (gram_schmidt.py:125-222) uses the same primitives."""
    from pymor.algorithms.gram_schmidt import gram_schmidt_biorth
    A, nsp, csp, nprod, cprod = make(120, 7, 5, kind)
    B = np.asfortranarray(A + 0.3 * np.random.default_rng(6).standard_normal(A.shape))
    Vn, Wn = gram_schmidt_biorth(nsp.from_numpy(A.copy()), nsp.from_numpy(B.copy()), product=nprod)
    Vc, Wc = gram_schmidt_biorth(csp.from_numpy(A), csp.from_numpy(B), product=cprod)
    np.testing.assert_allclose(Vc.to_numpy(), Vn.to_numpy(), atol=1e-10)
    np.testing.assert_allclose(Wc.to_numpy(), Wn.to_numpy(), atol=1e-10)
    np.testing.assert_allclose(Wc.inner(Vc, cprod), np.eye(7), atol=1e-12)


def test_more_reference_algorithms_run_unchanged():
    """Other pyMOR algorithms that only touch the backend through the VectorArray / Operator interface: the
    implicitly restarted Arnoldi eigensolver (`algorithms/eigs.py`), dynamic mode decomposition (`algorithms/dmd.py`,
    built on the SVD of svd_va.py) and the block Arnoldi process with `E^{-1}` (`algorithms/krylov.py:12`, device
    solver behind `apply_inverse`)."""
    from pymor.algorithms.dmd import dmd
    from pymor.algorithms.eigs import eigs
    from pymor.algorithms.krylov import arnoldi
    from pymor.tools.random import new_rng
    dim = 150
    K = sps.diags([-np.ones(dim - 1), 2.2 * np.ones(dim), -np.ones(dim - 1)], [-1, 0, 1]).tocsr()
    w = np.linspace(1, 2, dim)
    nsp, csp = NumpyVectorSpace(dim), CudaVectorSpace(dim, device=DEVICE)
    with new_rng(0):
        ew_n, _ = eigs(NumpyMatrixOperator(K), k=3, which='LM')
    with new_rng(0):
        ew_c, ev_c = eigs(CudaCsrOperator(K, csp), k=3, which='LM')
    np.testing.assert_allclose(np.sort(ew_c.real), np.sort(ew_n.real), rtol=1e-8)
    assert isinstance(ev_c, CudaVectorArray)
    rng = np.random.default_rng(1)
    Aop = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    X = np.zeros((dim, 12))
    X[:, 0] = rng.standard_normal(dim)
    for i in range(1, 12):
        X[:, i] = Aop @ X[:, i - 1]
    _, om_n = dmd(nsp.from_numpy(X), modes=5)
    _, om_c = dmd(csp.from_numpy(X), modes=5)
    np.testing.assert_allclose(np.sort_complex(om_c), np.sort_complex(om_n), rtol=1e-7)
    b = rng.standard_normal((dim, 1))
    Vn = arnoldi(NumpyMatrixOperator(K), NumpyMatrixOperator(sps.diags(w).tocsr()), nsp.from_numpy(b), 5)
    Vc = arnoldi(CudaCsrOperator(K, csp), CudaDiagonalOperator(w, csp), csp.from_numpy(b), 5)
    assert isinstance(Vc, CudaVectorArray) and len(Vc) == len(Vn)
    G = nsp.from_numpy(Vc.to_numpy()).inner(Vn)
    np.testing.assert_allclose(np.abs(np.linalg.svd(G, compute_uv=False)), 1.0, atol=1e-6)


def test_accelerate_pymor_routes_device_arrays_to_block_gram_schmidt():
    """`pymor_b200.accelerate.accelerate_pymor()`: an unchanged pyMOR program -- `pod` with its default `qr_svd`
    method, `extend_basis` of a reductor -- runs its Gram-Schmidt through the block variant when the arrays are
    device arrays, and through pyMOR's own function otherwise; `restore_pymor()` undoes the patch."""
    import pymor.algorithms.gram_schmidt as gs_mod
    import pymor.algorithms.pod as pod_mod
    import pymor.algorithms.svd_va as svd_mod

    from pymor_b200 import algorithms as alg
    from pymor_b200.accelerate import accelerate_pymor, restore_pymor
    A, nsp, csp, nprod, cprod = make(157, 12, 1, 'diag')
    original = gs_mod.gram_schmidt
    calls = []
    real_bgs = alg.block_gram_schmidt

    def spy(*a, **kw):
        calls.append('bcgs2')
        return real_bgs(*a, **kw)
    alg.block_gram_schmidt = spy
    try:
        accelerate_pymor('bcgs2')
        assert svd_mod.gram_schmidt is not original and pod_mod.gram_schmidt is svd_mod.gram_schmidt
        Un, sn = pod(nsp.from_numpy(A.copy()), product=nprod)                 # NumPy arrays: pyMOR's own path
        assert not calls
        Uc, sc = pod_mod.pod(csp.from_numpy(A), product=cprod)                # default method qr_svd -> block GS
        assert calls and isinstance(Uc, CudaVectorArray)
        np.testing.assert_allclose(sc, sn, rtol=1e-10)
        G = nsp.from_numpy(Uc.to_numpy()).inner(Un, nprod)
        np.testing.assert_allclose(np.abs(np.linalg.svd(G, compute_uv=False)), 1.0, atol=1e-8)
    finally:
        restore_pymor()
        alg.block_gram_schmidt = real_bgs
    assert svd_mod.gram_schmidt is original and gs_mod.gram_schmidt is original


def test_accelerate_pymor_selects_the_divide_and_conquer_eigensolver():
    """`accelerate_pymor(eigh_driver='evd')`: the full eigensolve of `method_of_snapshots` (svd_va.py:82) runs with
    LAPACK's divide-and-conquer driver, a call for a subset (modes given) keeps pyMOR's choice, the results agree
    with the unpatched run, and `restore_pymor()` puts scipy.linalg back."""
    import scipy.linalg as spla
    import pymor.algorithms.svd_va as svd_mod

    from pymor_b200.accelerate import accelerate_pymor, restore_pymor
    A, nsp, csp, nprod, cprod = make(211, 14, 3, 'diag')
    U0, s0, _ = method_of_snapshots(csp.from_numpy(A), product=cprod)
    U0m, s0m, _ = method_of_snapshots(csp.from_numpy(A), product=cprod, modes=5)
    seen = []
    real = spla.eigh

    def spy(a, *args, **kw):
        seen.append((kw.get('driver'), kw.get('subset_by_index')))
        return real(a, *args, **kw)
    spla.eigh = spy
    try:
        accelerate_pymor(eigh_driver='evd')
        assert svd_mod.spla is not spla
        U1, s1, _ = svd_mod.method_of_snapshots(csp.from_numpy(A), product=cprod)
        U1m, s1m, _ = svd_mod.method_of_snapshots(csp.from_numpy(A), product=cprod, modes=5)
    finally:
        restore_pymor()
        spla.eigh = real
    assert svd_mod.spla is spla
    assert seen[0] == ('evd', None) and seen[1][0] is None and seen[1][1] is not None
    np.testing.assert_allclose(s1, s0, rtol=1e-12)
    np.testing.assert_allclose(s1m, s0m, rtol=1e-12)
    G = U1.inner(U0, cprod)
    np.testing.assert_allclose(np.abs(np.diag(G)), 1.0, atol=1e-9)
