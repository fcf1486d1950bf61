"""pymor_b200 -- device-resident fp64 snapshot linear algebra for pyMOR on NVIDIA B200 (sm_100a).

Public API::

    from pymor_b200 import CudaVectorSpace, CudaDiagonalOperator, CudaCsrOperator
    space = CudaVectorSpace(dim)                  # one GPU
    A = space.from_numpy(snapshots)               # (dim, len) host array -> HBM
    U, s = pod(A, product=CudaDiagonalOperator(w, space), method='method_of_snapshots')

where ``pod`` is pyMOR's own ``pymor.algorithms.pod.pod`` when pyMOR is installed, or
:func:`pymor_b200.algorithms.pod` otherwise.
"""
from pymor_b200.interface import HAVE_PYMOR  # noqa: F401


def __getattr__(name):
    # lazy: importing the package must not require torch / the built library
    if name in ('CudaVectorSpace', 'CudaVectorArray', 'CudaVectorArrayImpl'):
        from pymor_b200 import vectorarray
        return getattr(vectorarray, name)
    if name in ('CudaDiagonalOperator', 'CudaCsrOperator'):
        from pymor_b200 import operators
        return getattr(operators, name)
    raise AttributeError(name)
