"""The reference's OWN test-suite and its OWN algorithms on the real kernels (B200).

pyMOR (unmodified, ``baseline/_ref``) is the host library on the GPU box; these tests hand the device backend to

* the reference's contract suite -- ``src/pymortests/vectorarray.py`` (the VectorArray contract, hypothesis-driven,
  float64 and complex128) and ``src/pymortests/algorithms/{gram_schmidt,svd_va,pod,chol_qr,basic}.py`` with the
  backend ``'cuda'`` registered (tests/reference_contract/run_contract.py --real);
* the reference's Operator contract and projection suites -- ``src/pymortests/operators/operators.py``,
  ``src/pymortests/algorithms/projection.py`` with the device operators as fixture generators
  (run_operator_contract.py --real);
* pyMOR's own ``gram_schmidt`` / ``pod`` / ``method_of_snapshots`` / ``qr_svd`` / ``project`` / ``shifted_chol_qr`` (N1) /
  ``RandomizedRangeFinder`` (adaptive) + ``randomized_svd`` (N2) / HAPOD (N3) / ``StationaryRBReductor`` with
  ``extend_basis`` + ``reconstruct`` and ``CoerciveRBReductor`` with device-CG Riesz representatives (N4, a17) /
  ``ei_greedy`` + ``deim`` (N5), each run twice -- NumPy classes and CUDA classes from the same bits -- and compared
  (tests/reference_contract/pymor_mode_checks.py), including the reference's stored golden vectors
  (``pymortests/testdata/check_results``).

Each runs in a subprocess (the reference's conftest.py and hypothesis profile must be the ones in charge).
"""
import re
import subprocess
import sys

import pytest
from conftest import HAVE_REFERENCE, ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAVE_REFERENCE, reason='reference tree not present')]


def _run(script, *args):
    r = subprocess.run([sys.executable, str(ROOT / 'tests' / 'reference_contract' / script), '--real', *args],
                       capture_output=True, text=True, timeout=2400)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert ' passed' in tail and not re.search(r'\d+ (failed|error)', tail), tail
    return tail


def test_reference_vectorarray_and_algorithm_suites_on_kernels():
    _run('run_contract.py')


def test_reference_operator_and_projection_suites_on_kernels():
    _run('run_operator_contract.py')


def test_pymor_algorithms_unchanged_on_kernels():
    _run('run_contract.py', str(ROOT / 'tests' / 'reference_contract' / 'pymor_mode_checks.py'))
