"""Base classes the CUDA backend derives from.

With pyMOR importable these ARE pyMOR's classes -- ``CudaVectorArray`` then is a genuine
``pymor.vectorarrays.interface.VectorArray`` and pyMOR's own ``gram_schmidt`` / ``pod`` /
``method_of_snapshots`` / ``qr_svd`` / ``project`` run on it unchanged.  Without pyMOR (the GPU box:
the reference cannot be pip-installed offline) the stand-alone mirror of the same
contract is used.
"""
try:  # pragma: no cover - depends on the environment
    from pymor.operators.interface import Operator
    from pymor.vectorarrays.interface import VectorArray, VectorArrayImpl, VectorSpace
    from pymor.vectorarrays.interface import _create_random_values as create_random_values
    HAVE_PYMOR = True
except Exception:  # ImportError or a broken partial installation
    from pymor_b200._mirror import Operator, VectorArray, VectorArrayImpl, VectorSpace, create_random_values
    HAVE_PYMOR = False

__all__ = ['HAVE_PYMOR', 'Operator', 'VectorArray', 'VectorArrayImpl', 'VectorSpace', 'create_random_values']
