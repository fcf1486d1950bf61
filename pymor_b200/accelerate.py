"""Opt-in acceleration of an UNCHANGED pyMOR program: route pyMOR's ``gram_schmidt`` to the device-friendly
variants when (and only when) it is called on a :class:`~pymor_b200.vectorarray.CudaVectorArray`.

pyMOR's modified Gram-Schmidt needs every projection coefficient on the host before the next pair
(src/pymor/algorithms/gram_schmidt.py:84-91): k(k-1) host-synchronised pair-steps, which is the one part of the hot
path that cannot run at device speed through the plug-in boundary alone.  Its callers (``pod`` with the default
``qr_svd`` method, ``ProjectionBasedReductor.extend_basis``, ``estimate_image_hierarchical``, ``rand_la``,
``krylov``, ...) import the function by name, so :func:`accelerate_pymor` swaps that name in every loaded
``pymor`` module for a dispatcher:

* arguments that are device arrays  ->  :func:`pymor_b200.algorithms.gram_schmidt` with the chosen ``variant``
  (``'bcgs2'``: projections on the tensor cores; ``'cgs2'``: two host syncs per pass; ``'mgs'``: the reference
  order, optionally as fused device sweeps) -- same signature, same contract, the same (unique) QR factorisation;
* anything else  ->  pyMOR's own function, untouched.

The second host-side lever is the small eigensolve of the method of snapshots (src/pymor/algorithms/svd_va.py:82:
``spla.eigh(B, overwrite_a=True, subset_by_index=...)``, LAPACK ``dsyevr``).  It does not shrink when the rows are
sharded over more GPUs, so it bounds the strong scaling of POD (Amdahl).  With ``eigh_driver='evd'`` the module's
``spla`` is replaced by a proxy whose ``eigh`` selects LAPACK's divide-and-conquer driver (``dsyevd``: BLAS-3 rich,
about 1.5x faster for all eigenpairs of a 1000 x 1000 Gramian) whenever ALL eigenpairs are requested; a call for a
subset keeps pyMOR's choice.  Same eigenpairs to rounding.

Nothing is patched unless this is called; :func:`restore_pymor` undoes it.
"""
from __future__ import annotations

import sys

_original = None
_patched_modules = []
_original_spla = None


class _SplaProxy:
    """``scipy.linalg`` with a different default LAPACK driver for full symmetric eigenproblems."""

    def __init__(self, module, driver):
        self._module, self._driver = module, driver

    def __getattr__(self, name):
        return getattr(self._module, name)

    def eigh(self, a, *args, **kwargs):
        if (self._driver and kwargs.get('driver') is None and kwargs.get('subset_by_index') is None
                and kwargs.get('subset_by_value') is None and len(args) == 0):
            kwargs['driver'] = self._driver
        return self._module.eigh(a, *args, **kwargs)


def accelerate_pymor(variant='bcgs2', sweep=None, eigh_driver='evd'):
    """Install the dispatcher (idempotent).  Returns the dispatcher."""
    global _original, _original_spla
    import pymor.algorithms.gram_schmidt as gs_mod
    import pymor.algorithms.svd_va as svd_va_mod

    from pymor_b200 import algorithms as alg
    from pymor_b200.vectorarray import CudaVectorArray
    if _original is None:
        _original = gs_mod.gram_schmidt
    original = _original

    def gram_schmidt(A, *args, **kwargs):
        if isinstance(A, CudaVectorArray):
            kwargs.setdefault('variant', variant)
            if sweep is not None:
                kwargs.setdefault('sweep', sweep)
            return alg.gram_schmidt(A, *args, **kwargs)
        return original(A, *args, **kwargs)

    gram_schmidt.__doc__ = original.__doc__
    gram_schmidt.__wrapped__ = original
    restore_pymor()
    _original = original
    if eigh_driver:
        _original_spla = svd_va_mod.spla
        svd_va_mod.spla = _SplaProxy(_original_spla, eigh_driver)
    for name, module in list(sys.modules.items()):
        if (name == 'pymor' or name.startswith('pymor.')) and module is not None \
                and getattr(module, 'gram_schmidt', None) is original:
            setattr(module, 'gram_schmidt', gram_schmidt)
            _patched_modules.append(module)
    return gram_schmidt


def restore_pymor():
    """Put pyMOR's own ``gram_schmidt`` back everywhere."""
    global _original, _original_spla
    if _original_spla is not None:
        import pymor.algorithms.svd_va as svd_va_mod
        svd_va_mod.spla = _original_spla
        _original_spla = None
    if _original is None:
        return
    while _patched_modules:
        setattr(_patched_modules.pop(), 'gram_schmidt', _original)
    _original = None
