"""Run the REFERENCE's own pytest/hypothesis contract suite against the CUDA backend's host layer.

Needs the reference tree (baseline/_ref, placed by __graft_entry__.build(); else /root/reference).  The reference parametrises its VectorArray tests over
backend names looked up in ``pymortests.strategies`` (src/pymortests/strategies.py:110-184); external
backends register by adding a ``_<name>_vector_spaces`` factory.  We register ``'cuda'`` from here
(the reference tree is read-only and stays untouched) and hand over to pytest; hypothesis draws
float64 and complex128 inputs exactly as for the reference's own backends.

Without a GPU the C ABI is served by tests/emulated_lib.py (NumPy), so what this validates is the
host-side view / copy-on-write / aliasing / growth logic and the pyMOR integration; with a GPU
(``--real``) the same suite runs on the kernels.

usage: python tests/reference_contract/run_contract.py [--real] [pytest args...]
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
from hypothesis import strategies as hyst  # noqa: E402

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

import pymortests.strategies as pyst  # noqa: E402

from pymor_b200.vectorarray import CudaVectorSpace  # noqa: E402

_spaces = {}


def _cuda_vector_spaces(draw, np_data_list, compatible, count, dims):
    out = []
    for d, ar in zip(dims, np_data_list, strict=True):
        if d not in _spaces:
            _spaces[d] = CudaVectorSpace(d, device=device)
        out.append((_spaces[d], ar))
    return out


pyst._cuda_vector_spaces = _cuda_vector_spaces
pyst._picklable_vector_space_types = []
pyst._other_vector_space_types = ['cuda']

import pytest  # noqa: E402

DEFAULT = ['src/pymortests/vectorarray.py', 'src/pymortests/algorithms/gram_schmidt.py',
           'src/pymortests/algorithms/svd_va.py', 'src/pymortests/algorithms/pod.py',
           'src/pymortests/algorithms/chol_qr.py', 'src/pymortests/algorithms/basic.py']
import os  # noqa: E402

os.chdir(REF)
sys.exit(pytest.main(['-p', 'no:cacheprovider', '-q', '--rootdir', str(REF), '-c', str(REF / 'pyproject.toml'),
                      *(args or DEFAULT)]))
