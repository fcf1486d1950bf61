"""Linear solvers for the CUDA operators: what ``Operator.apply_inverse`` needs on the device.

pyMOR computes Riesz representatives of residuals with ``product.apply_inverse`` (reference:
src/pymor/reductors/residual.py:151, src/pymor/algorithms/image.py:106) and dispatches ``apply_inverse`` to the
operator's ``solver`` (src/pymor/operators/interface.py:174-208, src/pymor/solvers/interface.py).  The reference's
NumPy operators hand the matrix to SciPy's sparse direct solver on the CPU; for device-resident operators the
natural solver is a Krylov method written against the Operator / VectorArray interface, so that every step is
one of the backend's kernels:

* :class:`CgSolver` -- conjugate gradients for symmetric positive definite operators (mass / stiffness /
  energy products), ALL right-hand sides at once: per iteration one SpMM (``Operator.apply`` on the whole
  block), two ``pairwise_inner``, one ``norm``, three per-column ``axpy`` / ``scal``; optional Jacobi
  preconditioner.  Converged columns are frozen (coefficient 0).
* :class:`DiagonalSolver` -- exact inverse of a :class:`~pymor_b200.operators.CudaDiagonalOperator`.

Both are ``pymor.solvers.interface.Solver`` subclasses (usable as ``op.with_(solver=...)`` or
``apply_inverse(..., solver=...)``).
They use nothing but the interface, so they also run on pyMOR's NumPy classes (that is how they are checked
against SciPy on the CPU).
"""
from __future__ import annotations

import numpy as np

from pymor_b200 import interface as _interface  # noqa: F401  (puts pyMOR on sys.path)
from pymor.core.exceptions import InversionError  # noqa: E402
from pymor.solvers.interface import Solver as _SolverBase  # noqa: E402


def cg(apply, V, tol=1e-10, maxiter=None, preconditioner=None, initial_guess=None):
    """Conjugate gradients for ``A X = V`` with symmetric positive definite ``A`` given as ``apply(X) -> A X``
    (|VectorArray| in, |VectorArray| out), for all ``len(V)`` right-hand sides at once.  A column stops when
    ``||r|| <= tol * ||v||``.  Returns ``X`` and ``{'iterations', 'residual_norms'}``; raises
    :class:`InversionError` if some column has not converged after ``maxiter`` iterations (default ``10 dim``)
    or a search direction of non-positive curvature shows up (operator not positive definite)."""
    k = len(V)
    if maxiter is None:
        maxiter = 10 * max(V.dim, 1)
    X = V.space.zeros(k) if initial_guess is None else initial_guess.copy()
    R = V.copy()
    if initial_guess is not None:
        R.axpy(-1., apply(X))
    bnorm = V.norm()
    rnorm = R.norm()
    target = tol * bnorm
    active = rnorm > target
    it = 0
    if k and active.any():
        Z = R if preconditioner is None else preconditioner(R)
        P = Z.copy()
        rz = R.pairwise_inner(Z)
        while True:
            AP = apply(P)
            pAp = P.pairwise_inner(AP)
            if np.any(pAp[active] <= 0):
                raise InversionError('cg: direction of non-positive curvature (operator not positive definite?)')
            alpha = np.where(active, rz / np.where(active, pAp, 1.), 0.)
            X.axpy(alpha, P)
            R.axpy(-alpha, AP)
            it += 1
            rnorm = R.norm()
            active = rnorm > target
            if not active.any():
                break
            if it >= maxiter:
                raise InversionError(f'cg failed to converge after {it} iterations (worst relative residual '
                                     f'{np.max(rnorm / np.where(bnorm > 0, bnorm, 1.)):.2e})')
            Z = R if preconditioner is None else preconditioner(R)
            rz_new = R.pairwise_inner(Z)
            beta = np.where(active, rz_new / np.where(rz != 0, rz, 1.), 0.)
            P.scal(beta)
            P.axpy(1., Z)
            rz = rz_new
    return X, {'iterations': it, 'residual_norms': rnorm}


class CgSolver(_SolverBase):
    """Conjugate gradients on the whole block of right-hand sides (see :func:`cg`).

    Parameters
    ----------
    tol
        Relative residual at which a column stops.
    maxiter
        Iteration limit (``None``: ten times the dimension).
    jacobi
        Precondition with the inverse diagonal when the operator provides ``jacobi_preconditioner()``
        (:class:`~pymor_b200.operators.CudaCsrOperator` does).
    """

    def __init__(self, tol=1e-10, maxiter=None, jacobi=True):
        self.tol, self.maxiter, self.jacobi = tol, maxiter, jacobi
        self.__auto_init(locals())       # ImmutableObject: stores the arguments and locks the instance

    def _solve(self, operator, V, mu, initial_guess):
        op = operator.assemble(mu)
        pre = None
        if self.jacobi and hasattr(op, 'jacobi_preconditioner'):
            jac = op.jacobi_preconditioner()
            pre = None if jac is None else jac.apply
        return cg(lambda X: op.apply(X), V, tol=self.tol, maxiter=self.maxiter, preconditioner=pre,
                  initial_guess=initial_guess)

    def _solve_adjoint(self, operator, U, mu, initial_guess):          # symmetric operators only
        return self._solve(operator, U, mu, initial_guess)


class DiagonalSolver(_SolverBase):
    """Exact inverse of a diagonal operator: multiply by the reciprocal weights."""

    def __init__(self):
        self.__auto_init(locals())

    def _solve(self, operator, V, mu, initial_guess):
        return operator.assemble(mu).inverse().apply(V), {}

    def _solve_adjoint(self, operator, U, mu, initial_guess):
        return self._solve(operator, U, mu, initial_guess)
