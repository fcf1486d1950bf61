"""pytest configuration.

Markers
-------
gpu : needs a real B200 (run with ``-m gpu``); everything else runs on CPU (``-m "not gpu"``).

pyMOR is the host library of the product; it is imported from ``baseline/_ref`` (the unmodified reference tree
that ``__graft_entry__.build()`` places there; it travels to the GPU box).  Tests marked ``reference`` additionally
need the reference's own test-suite (``src/pymortests``) from that tree.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from pymor_b200.interface import reference_root  # noqa: E402

# the reference tree (pyMOR + its own test-suite): baseline/_ref, placed by __graft_entry__.build() and shipped to
# the GPU box; /root/reference itself exists only in the build container and is never read by a -m gpu test
REFERENCE_ROOT = reference_root()
HAVE_REFERENCE = REFERENCE_ROOT is not None and (REFERENCE_ROOT / 'src' / 'pymortests' / 'vectorarray.py').exists()
REFERENCE_SRC = REFERENCE_ROOT / 'src' if HAVE_REFERENCE else Path('/root/reference/src')
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')
    config.addinivalue_line('markers', 'reference: test runs the reference test-suite (baseline/_ref/src/pymortests)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    skip_gpu = pytest.mark.skip(reason='no CUDA device')
    skip_ref = pytest.mark.skip(reason='reference tree (baseline/_ref) not present')
    for item in items:
        if 'gpu' in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)
        if 'reference' in item.keywords and not HAVE_REFERENCE:
            item.add_marker(skip_ref)


def load_golden(name):
    return dict(np.load(GOLDEN / f'{name}.npz'))


def csr_from(g, prefix):
    import scipy.sparse as sps
    n = len(g[f'{prefix}_indptr']) - 1
    return sps.csr_matrix((g[f'{prefix}_data'], g[f'{prefix}_indices'], g[f'{prefix}_indptr']), shape=(n, n))


@pytest.fixture
def rng():
    return np.random.default_rng(0)


def assert_gram_schmidt_golden(Qh, R, g, name, W, tight=1e-12):
    """Compare a Gram-Schmidt result with the golden one of tests/golden/gram_schmidt.npz for a code path whose
    rounding differs from the reference's.  Column 4 of that input is column 1 plus 1e-9 noise: its orthogonalised
    direction is conditioned like 1e9 * eps w.r.t. ANY change of rounding or summation order, and the later columns
    inherit it through R[4, :].  So: the same vectors are removed, the columns before the near-dependency agree
    tightly, everything agrees at cond * eps, A = Q R holds to rounding, and the orthogonality error is no worse
    than the reference's."""
    from oracle import numpy_path as O
    Qref, Rref = g[f'Q_{name}'], g[f'R_{name}']
    assert Qh.shape == Qref.shape
    np.testing.assert_allclose(R, Rref, atol=1e-5)
    np.testing.assert_allclose(Qh, Qref, atol=1e-6)
    np.testing.assert_allclose(Qh[:, :4], Qref[:, :4], atol=tight)
    np.testing.assert_allclose(R[:4, :4], Rref[:4, :4], atol=10 * tight)
    np.testing.assert_allclose(Qh @ R, g['A'], atol=1e-12)
    assert O.orthogonality_error(Qh, W) <= max(4 * O.orthogonality_error(Qref, W), 1e-14)
