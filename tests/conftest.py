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
