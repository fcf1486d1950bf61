"""Run the REFERENCE's Operator contract tests (src/pymortests/operators/operators.py) on the CUDA
operators (CudaDiagonalOperator, CudaCsrOperator and pyMOR constructions built from them).

Needs the reference tree (baseline/_ref, placed by __graft_entry__.build(); else /root/reference).  The reference parametrises these tests over generator
lists in ``pymortests.fixtures.operator`` that are frozen into the fixtures at import; external backends
(FEniCS, dolfinx) append their generators there (fixtures/operator.py:617-627, :766-768).  The reference
tree is read-only, so the same registration is applied from outside: the two fixtures are re-created
with the CUDA generators before pytest collects them.  The C ABI is served by tests/emulated_lib.py
unless ``--real`` is given (then the kernels run).

usage: python tests/reference_contract/run_operator_contract.py [--real] [pytest args...]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests'))
from pymor_b200.interface import reference_root  # noqa: E402

REF = reference_root()            # baseline/_ref (placed by __graft_entry__.build(), travels to the GPU box)
assert REF is not None, 'the contract suite needs the reference tree (src/pymortests): run __graft_entry__.build()'
sys.path.insert(0, str(REF / 'src'))

import numpy as np  # noqa: E402
import pytest  # noqa: E402
import scipy.sparse as sps  # noqa: E402

args = sys.argv[1:]
real = '--real' in args
if real:
    args.remove('--real')
    device = 'cuda'
else:
    from emulated_lib import EmulatedLib

    from pymor_b200 import _lib
    _lib.set_library_for_testing(EmulatedLib())
    device = 'cpu'

import pymortests.fixtures.operator as fo  # noqa: E402
from pymor.operators.constructions import LincombOperator  # noqa: E402

from pymor_b200.operators import CudaCsrOperator, CudaDiagonalOperator  # noqa: E402
from pymor_b200.vectorarray import CudaVectorSpace  # noqa: E402


def _matrices(dim, seed):
    r = np.random.default_rng(seed)
    w = r.uniform(0.5, 1.5, dim)
    M = (sps.random(dim, dim, density=min(1.0, 6.0 / max(dim, 1)), random_state=seed) + 3 * sps.eye(dim)).tocsr()
    h = 1.0 / (dim + 1)
    P = sps.diags([np.full(dim - 1, h / 6), np.full(dim, 4 * h / 6), np.full(dim - 1, h / 6)], [-1, 0, 1]).tocsr()
    return w, M, P


def _diag(rng, dim=23, count=4):
    sp = CudaVectorSpace(dim, device=device)
    w, _, _ = _matrices(dim, 1)
    return CudaDiagonalOperator(w, sp), None, sp.random(count), sp.random(count)


def _csr(rng, dim=31, count=3):
    sp = CudaVectorSpace(dim, device=device)
    _, M, _ = _matrices(dim, 2)
    return CudaCsrOperator(M, sp), None, sp.random(count), sp.random(count)


def _lincomb(rng, dim=17, count=5):
    sp = CudaVectorSpace(dim, device=device)
    w, M, P = _matrices(dim, 3)
    op = LincombOperator([CudaCsrOperator(M, sp), CudaDiagonalOperator(w, sp), CudaCsrOperator(P, sp)], [1.5, -0.5, 2.0])
    return op, None, sp.random(count), sp.random(count)


def _with_products(gen):
    def make(rng):
        op, mu, U, V = gen(rng)
        sp = op.source
        w, _, P = _matrices(sp.dim, 7)
        return op, mu, U, V, CudaDiagonalOperator(w, sp), CudaCsrOperator(P, sp)
    return make


GENERATORS = [_diag, _csr, _lincomb]


@pytest.fixture(params=GENERATORS)
def operator_with_arrays(rng, request):
    return request.param(rng)


@pytest.fixture(params=[_with_products(g) for g in GENERATORS])
def operator_with_arrays_and_products(rng, request):
    return request.param(rng)


@pytest.fixture(params=[lambda rng, g=g: g(rng)[0:2] for g in GENERATORS])
def operator(rng, request):
    return request.param(rng)


fo.operator_with_arrays = operator_with_arrays
fo.operator_with_arrays_and_products = operator_with_arrays_and_products
fo.operator = operator

import os  # noqa: E402

os.chdir(REF)
# tests that build their own NumPy operators (no fixture) run unchanged and are not of interest here: select
# the fixture-driven ones; pickling is not part of the device backend's contract
DEFAULT = ['src/pymortests/operators/operators.py', 'src/pymortests/algorithms/projection.py', '-k',
           'not picklable and not pickle']
sys.exit(pytest.main(['-p', 'no:cacheprovider', '-q', '--rootdir', str(REF), '-c', str(REF / 'pyproject.toml'),
                      *(args or DEFAULT)]))
