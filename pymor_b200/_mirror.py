"""Host-side mirror of the pyMOR interfaces the CUDA backend plugs into.

When pyMOR is importable, :mod:`pymor_b200.interface` re-exports pyMOR's own ``VectorArray``,
``VectorSpace``, ``VectorArrayImpl`` and ``Operator`` and the CUDA classes subclass those (the real
drop-in).  pyMOR is pure Python but cannot be installed on the GPU box (its build backend is not in the
offline wheelhouse), so this module restates the part of the contract the hot path relies on -- same
names, argument meaning and error behaviour -- to let the backend run stand-alone:

* ``VectorArray``  : view indexing, copy-on-write reference counting, argument checks and the
  product dispatch (reference: src/pymor/vectorarrays/interface.py:14-794)
* ``VectorSpace``  : array factory base (interface.py:796-983)
* ``VectorArrayImpl``: the backend method table (interface.py:1010-1110)
* ``Operator``     : ``apply`` / ``apply2`` / ``pairwise_apply2`` / ``apply_adjoint`` boundary
  (src/pymor/operators/interface.py:63-172)
"""
from __future__ import annotations

from numbers import Integral, Number

import numpy as np


class VectorArrayImpl:
    """Backend method table.  ``ind`` / ``oind`` / ``xind`` are ``None``, a normalised ``slice`` or a
    list of non-negative ints (interface.py:1013-1020)."""

    def __len__(self):
        raise NotImplementedError

    def to_numpy(self, ensure_copy, ind):
        raise NotImplementedError

    def append(self, other, remove_from_other, oind):
        raise NotImplementedError

    def delete(self, ind):
        raise NotImplementedError

    def copy(self, deep, ind):
        raise NotImplementedError

    def scal(self, alpha, ind):
        raise NotImplementedError

    def axpy(self, alpha, x, ind, xind):
        raise NotImplementedError

    def inner(self, other, ind, oind):
        raise NotImplementedError

    def pairwise_inner(self, other, ind, oind):
        raise NotImplementedError

    def lincomb(self, coefficients, ind):
        raise NotImplementedError

    def norm2(self, ind):
        raise NotImplementedError

    def dofs(self, dof_indices, ind):
        raise NotImplementedError

    def amax(self, ind):
        raise NotImplementedError

    def real(self, ind):
        raise NotImplementedError

    def imag(self, ind):
        raise NotImplementedError

    def conj(self, ind):
        raise NotImplementedError

    # defaults expressed through the primitives (interface.py:1044-1075)
    def scal_copy(self, alpha, ind):
        out = self.copy(True, ind)
        out.scal(alpha, None)
        return out

    def axpy_copy(self, alpha, x, ind, xind):
        out = self.copy(True, ind)
        out.axpy(alpha, x, None, xind)
        return out

    def gramian(self, ind):
        return self.inner(self, ind, ind)

    def norm(self, ind):
        return np.sqrt(self.norm2(ind))

    def len_ind(self, ind):
        if ind is None:
            return len(self)
        if type(ind) is slice:
            return len(range(*ind.indices(len(self))))
        return len(ind) if hasattr(ind, '__len__') else 1


def _compose_index(base_len, outer, inner):
    """Index of ``base[outer][inner]`` expressed on ``base`` (interface.py:770-789)."""
    seq = range(*outer.indices(base_len)) if type(outer) is slice else list(outer)
    if type(inner) is slice:
        picked = seq[inner]
        if isinstance(picked, range):
            return slice(picked.start, picked.stop, picked.step)
        return list(picked)
    if hasattr(inner, '__len__'):
        return [seq[i] for i in inner]
    return [seq[inner]]


class VectorArray:
    """List of vectors with NumPy-like view indexing and copy-on-write copies."""

    __array_priority__ = 100.0
    __array_ufunc__ = None

    impl_type = None
    is_view = False
    base = None
    ind = None
    _impl = None

    def __init__(self, space, impl, base=None, ind=None, _len=None):
        assert impl is None or isinstance(impl, self.impl_type)
        assert base is None or impl is None
        assert ind is None or base is not None
        self.space = space
        if base is None:
            self._impl = impl
            self._refcount = [1]
            self._len = len(impl)
        else:
            self.is_view = True
            self.base = base
            self.ind = ind
            self._len = _len

    # -- plumbing -------------------------------------------------------------------------
    @property
    def impl(self):
        return self.base._impl if self.is_view else self._impl

    @property
    def dim(self):
        return self.space.dim

    def __len__(self):
        return self._len

    def zeros(self, count=1, reserve=0):
        return self.space.zeros(count, reserve=reserve)

    def ones(self, count=1, reserve=0):
        return self.space.full(1., count, reserve=reserve)

    def full(self, value, count=1, reserve=0):
        return self.space.full(value, count, reserve=reserve)

    def random(self, count=1, distribution='uniform', reserve=0, **kwargs):
        return self.space.random(count, distribution, reserve=reserve, **kwargs)

    def empty(self, reserve=0):
        return self.space.zeros(0, reserve=reserve)

    def __getitem__(self, ind):
        n = self._len
        if isinstance(ind, Integral):
            if not -n <= ind < n:
                raise IndexError('VectorArray index out of range')
            pos = int(ind) % n
            ind, count = slice(pos, pos + 1), 1
        elif type(ind) is slice:
            start, stop, step = ind.indices(n)
            count = len(range(start, stop, step))
            if count == 0:
                ind = slice(0, 0, 1)
            else:
                # keep the view length fixed when the base array grows later on
                ind = slice(start, None if stop < 0 else stop, step)
        else:
            assert isinstance(ind, (list, np.ndarray))
            assert all(-n <= i < n for i in ind)
            ind = [int(i) % n for i in ind]
            count = len(ind)
        if self.is_view:
            return type(self)(self.space, None, self.base, _compose_index(self.base._len, self.ind, ind), count)
        return type(self)(self.space, None, self, ind, count)

    def __delitem__(self, ind):
        if self.is_view:
            raise ValueError('Cannot delete items from VectorArray view')
        assert self.check_ind(ind)
        self._copy_impl_if_multiple_refs()
        self.impl.delete(ind)
        self._len = len(self.impl)

    def _copy_impl_if_multiple_refs(self):
        owner = self.base if self.is_view else self
        if owner._refcount[0] == 1:
            return
        owner._impl = owner._impl.copy(False, None)
        owner._refcount[0] -= 1
        owner._refcount = [1]

    def __del__(self):
        if not self.is_view and getattr(self, '_refcount', None) is not None:
            self._refcount[0] -= 1

    # -- storage ---------------------------------------------------------------------------
    def to_numpy(self, ensure_copy=False):
        return self.impl.to_numpy(ensure_copy, self.ind)

    def append(self, other, remove_from_other=False):
        assert other in self.space
        if self.is_view:
            raise ValueError('Cannot append to VectorArray view')
        if remove_from_other and self.impl is other.impl:
            raise ValueError('Cannot append VectorArray to itself with remove_from_other=True')
        self._copy_impl_if_multiple_refs()
        if remove_from_other:
            other._copy_impl_if_multiple_refs()
        self.impl.append(other.impl, remove_from_other, other.ind)
        self._len += other._len
        if remove_from_other:
            if other.is_view:
                other.base._len = len(other.base._impl)
            else:
                other._len = 0

    def copy(self, deep=False):
        if self.is_view or deep:
            return type(self)(self.space, self.impl.copy(deep, self.ind))
        twin = type(self)(self.space, self.impl)
        twin._refcount = self._refcount
        self._refcount[0] += 1
        return twin

    def __deepcopy__(self, memo):
        return self.copy(deep=True)

    # -- BLAS-1 ----------------------------------------------------------------------------
    def scal(self, alpha):
        assert not self.is_view or self.base.check_ind_unique(self.ind)
        assert isinstance(alpha, Number) or isinstance(alpha, np.ndarray) and alpha.shape == (len(self),)
        self._copy_impl_if_multiple_refs()
        self.impl.scal(alpha, self.ind)

    def axpy(self, alpha, x):
        assert not self.is_view or self.base.check_ind_unique(self.ind)
        assert isinstance(alpha, Number) or isinstance(alpha, np.ndarray) and alpha.shape == (len(self),)
        assert x in self.space
        assert len(self) == len(x) or len(x) == 1
        self._copy_impl_if_multiple_refs()
        self.impl.axpy(alpha, x.impl, self.ind, x.ind)

    # -- reductions ------------------------------------------------------------------------
    def inner(self, other, product=None):
        if product is not None:
            return product.apply2(self, other)
        assert self.space == other.space
        return self.impl.inner(other.impl, self.ind, other.ind)

    def pairwise_inner(self, other, product=None):
        if product is not None:
            return product.pairwise_apply2(self, other)
        assert self.space == other.space
        assert len(self) == len(other)
        return self.impl.pairwise_inner(other.impl, self.ind, other.ind)

    def gramian(self, product=None):
        if product is not None:
            return product.apply2(self, self)
        return self.impl.gramian(self.ind)

    def lincomb(self, coefficients):
        assert 1 <= coefficients.ndim <= 2
        if coefficients.ndim == 1:
            coefficients = coefficients[..., np.newaxis]
        assert coefficients.shape[0] == len(self)
        return type(self)(self.space, self.impl.lincomb(coefficients, self.ind))

    def norm(self, product=None, tol=None, raise_complex=None):
        if product is not None:
            kw = {}
            if tol is not None:
                kw['tol'] = tol
            if raise_complex is not None:
                kw['raise_complex'] = raise_complex
            return np.sqrt(self.norm2(product=product, **kw).real)
        result = self.impl.norm(self.ind)
        assert np.all(np.isrealobj(result))
        return result

    def norm2(self, product=None, tol=1e-10, raise_complex=True):
        if product is not None:
            sq = product.pairwise_apply2(self, self)
            if raise_complex and np.any(np.abs(sq.imag) > tol):
                raise ValueError(f'norm is complex (square = {sq})')
            return sq.real
        result = self.impl.norm2(self.ind)
        assert np.all(np.isrealobj(result))
        return result

    def sup_norm(self):
        if self.dim == 0:
            return np.zeros(len(self))
        return self.amax()[1]

    def dofs(self, dof_indices):
        idx = np.asarray(dof_indices, dtype=np.int64)
        assert idx.ndim == 1 and (idx.size == 0 or idx.min() >= 0)
        assert idx.size == 0 or idx.max() < self.dim
        return self.impl.dofs(idx, self.ind)

    def amax(self):
        assert self.dim > 0
        return self.impl.amax(self.ind)

    @property
    def real(self):
        return type(self)(self.space, self.impl.real(self.ind))

    @property
    def imag(self):
        return type(self)(self.space, self.impl.imag(self.ind))

    def conj(self):
        return type(self)(self.space, self.impl.conj(self.ind))

    # -- operator sugar --------------------------------------------------------------------
    def __add__(self, other):
        if isinstance(other, Number):
            assert other == 0
            return self.copy()
        assert other in self.space
        assert len(self) == len(other) or len(other) == 1
        return type(self)(self.space, self.impl.axpy_copy(1, other.impl, self.ind, other.ind))

    __radd__ = __add__

    def __iadd__(self, other):
        self.axpy(1, other)
        return self

    def __sub__(self, other):
        assert other in self.space
        assert len(self) == len(other) or len(other) == 1
        return type(self)(self.space, self.impl.axpy_copy(-1, other.impl, self.ind, other.ind))

    def __isub__(self, other):
        self.axpy(-1, other)
        return self

    def __mul__(self, other):
        assert isinstance(other, Number) or isinstance(other, np.ndarray) and other.shape == (len(self),)
        return type(self)(self.space, self.impl.scal_copy(other, self.ind))

    __rmul__ = __mul__

    def __imul__(self, other):
        self.scal(other)
        return self

    def __neg__(self):
        return type(self)(self.space, self.impl.scal_copy(-1, self.ind))

    # -- index checks ----------------------------------------------------------------------
    def check_ind(self, ind):
        n = len(self)
        if type(ind) is slice:
            return True
        if isinstance(ind, Number):
            return -n <= ind < n
        return isinstance(ind, (list, np.ndarray)) and all(-n <= i < n for i in ind)

    def check_ind_unique(self, ind):
        n = len(self)
        if type(ind) is slice:
            return True
        if isinstance(ind, Number):
            return -n <= ind < n
        if not isinstance(ind, (list, np.ndarray)):
            return False
        return len({i % n for i in ind if -n <= i < n}) == len(ind)

    def len_ind(self, ind):
        return self.impl.len_ind(ind)

    def len_ind_unique(self, ind):
        n = len(self)
        if type(ind) is slice:
            return len(range(*ind.indices(n)))
        if isinstance(ind, Number):
            return 1
        return len({i % n for i in ind})


class VectorSpace:
    """Factory of vector arrays of one kind and dimension."""

    id = None
    dim = None
    is_scalar = False

    def make_array(self, *args, **kwargs):
        raise NotImplementedError

    def zeros(self, count=1, reserve=0):
        raise NotImplementedError

    def ones(self, count=1, reserve=0):
        return self.full(1., count, reserve)

    def full(self, value, count=1, reserve=0):
        return self.from_numpy(np.full((self.dim, count), value))

    def random(self, count=1, distribution='uniform', reserve=0, **kwargs):
        return self.from_numpy(create_random_values((self.dim, count), distribution, **kwargs))

    def empty(self, reserve=0):
        return self.zeros(0, reserve=reserve)

    def from_numpy(self, data, ensure_copy=False):
        raise NotImplementedError

    def __eq__(self, other):
        return other is self

    def __ne__(self, other):
        return not (self == other)

    def __contains__(self, other):
        return self == getattr(other, 'space', None)

    def __hash__(self):
        return id(self)


_rng = np.random.default_rng(42)


def set_rng(rng):
    """Stand-alone replacement for pyMOR's global RNG (src/pymor/tools/random.py)."""
    global _rng
    _rng = rng


def create_random_values(shape, distribution, **kwargs):
    """Random (dim, count) values, drawn vector by vector (interface.py:986-1007)."""
    if distribution not in ('uniform', 'normal'):
        raise NotImplementedError
    if distribution == 'uniform':
        if not kwargs.keys() <= {'low', 'high'}:
            raise ValueError
        low, high = kwargs.get('low', 0.), kwargs.get('high', 1.)
        if high <= low:
            raise ValueError
        draw = _rng.uniform(low, high, shape)
    else:
        if not kwargs.keys() <= {'loc', 'scale'}:
            raise ValueError
        draw = _rng.normal(kwargs.get('loc', 0.), kwargs.get('scale', 1.), shape)
    return draw.ravel().reshape(shape, order='F')


class Operator:
    """Linear operator boundary: only what the hot path touches."""

    linear = True
    source = None
    range = None
    parametric = False
    name = None

    def apply(self, U, mu=None):
        raise NotImplementedError

    def apply_adjoint(self, V, mu=None):
        raise NotImplementedError

    def apply2(self, V, U, mu=None):
        assert isinstance(V, VectorArray) and isinstance(U, VectorArray)
        return V.inner(self.apply(U, mu=mu))

    def pairwise_apply2(self, V, U, mu=None):
        assert isinstance(V, VectorArray) and isinstance(U, VectorArray)
        assert len(U) == len(V)
        return V.pairwise_inner(self.apply(U, mu=mu))

    solver = None

    def apply_inverse(self, V, mu=None, initial_guess=None, return_info=False, solver=None):
        """Dispatch to the operator's solver (operators/interface.py:174-208)."""
        solver = solver or self.solver
        if solver is None:
            raise NotImplementedError('operator has no solver')
        return solver.solve(self, V, mu=mu, initial_guess=initial_guess, return_info=return_info)

    def apply_inverse_adjoint(self, U, mu=None, initial_guess=None, return_info=False, solver=None):
        solver = solver or self.solver
        if solver is None:
            raise NotImplementedError('operator has no solver')
        return solver.solve_adjoint(self, U, mu=mu, initial_guess=initial_guess, return_info=return_info)

    def assemble(self, mu=None):
        return self
