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

Nothing is patched unless this is called; :func:`restore_pymor` undoes it.
"""
from __future__ import annotations

import sys

_original = None
_patched_modules = []


def accelerate_pymor(variant='bcgs2', sweep=None):
    """Install the dispatcher (idempotent).  Returns the dispatcher."""
    global _original
    import pymor.algorithms.gram_schmidt as gs_mod

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
    for name, module in list(sys.modules.items()):
        if (name == 'pymor' or name.startswith('pymor.')) and module is not None \
                and getattr(module, 'gram_schmidt', None) is original:
            setattr(module, 'gram_schmidt', gram_schmidt)
            _patched_modules.append(module)
    return gram_schmidt


def restore_pymor():
    """Put pyMOR's own ``gram_schmidt`` back everywhere."""
    global _original
    if _original is None:
        return
    while _patched_modules:
        setattr(_patched_modules.pop(), 'gram_schmidt', _original)
    _original = None
