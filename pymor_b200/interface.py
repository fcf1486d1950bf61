"""pyMOR's plug-in base classes -- the boundary this backend sits behind.

pyMOR is the host library: ``CudaVectorArray`` is a genuine
``pymor.vectorarrays.interface.VectorArray`` (reference: src/pymor/vectorarrays/interface.py:14),
``CudaVectorArrayImpl`` a ``VectorArrayImpl`` (:1010), ``CudaVectorSpace`` a ``VectorSpace`` (:796) and the
device operators ``pymor.operators.interface.Operator`` subclasses (src/pymor/operators/interface.py:18), so
that pyMOR's own ``gram_schmidt`` / ``pod`` / ``method_of_snapshots`` / ``qr_svd`` / ``project`` / reductors
run on device arrays unchanged.  There is no stand-alone substitute for these classes: without pyMOR the
import fails.

pyMOR is looked up (1) as an installed package, (2) in ``baseline/_ref/src`` -- the unmodified pure-Python
reference tree that ``__graft_entry__.build()`` places there (git-ignored, travels to the GPU box with the
built library; the offline ``pip install`` of the reference fails on its ``hatchling`` build backend,
DESIGN.md §1), (3) in ``/root/reference/src`` (the build container).
"""
import importlib.util
import sys
from pathlib import Path

_ROOT = Path(__file__).resolve().parents[1]
REFERENCE_ROOTS = (_ROOT / 'baseline' / '_ref', Path('/root/reference'))


def reference_root():
    """Root of the reference checkout in use (holds ``src/pymor``, ``src/pymortests``, ``conftest.py``), or
    None when pyMOR is an installed package."""
    for root in REFERENCE_ROOTS:
        if (root / 'src' / 'pymor' / '__init__.py').exists():
            return root
    return None


if importlib.util.find_spec('pymor') is None:
    _root = reference_root()
    if _root is None:
        raise ImportError('pymor_b200 is a backend for pyMOR and needs it importable: install pyMOR or run '
                          '`python -c "import __graft_entry__ as g; g.build()"` where /root/reference exists '
                          '(it places the reference under baseline/_ref).  There is no stand-alone mode.')
    sys.path.insert(0, str(_root / 'src'))

from pymor.operators.interface import Operator  # noqa: E402
from pymor.vectorarrays.interface import VectorArray, VectorArrayImpl, VectorSpace  # noqa: E402
from pymor.vectorarrays.interface import _create_random_values as create_random_values  # noqa: E402

HAVE_PYMOR = True

__all__ = ['HAVE_PYMOR', 'Operator', 'VectorArray', 'VectorArrayImpl', 'VectorSpace', 'create_random_values',
           'reference_root']
