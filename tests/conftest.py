"""pytest configuration.

Markers
-------
gpu : needs a real B200 (run with ``-m gpu``); everything else runs on CPU (``-m "not gpu"``).

The reference (pyMOR) is importable only in the build container (``/root/reference``); tests that
drive the reference live are skipped when it is absent, the committed fixtures in ``tests/golden``
stand in for it.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE_SRC = Path('/root/reference/src')
HAVE_REFERENCE = (REFERENCE_SRC / 'pymor' / '__init__.py').exists()
GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (B200)')
    config.addinivalue_line('markers', 'reference: test imports the reference from /root/reference')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    skip_gpu = pytest.mark.skip(reason='no CUDA device')
    skip_ref = pytest.mark.skip(reason='/root/reference not present')
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
