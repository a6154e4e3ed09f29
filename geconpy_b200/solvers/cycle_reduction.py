"""``gEconpy.solvers.cycle_reduction`` on B200: same names and signatures, CUDA underneath.

Reference: gEconpy/solvers/cycle_reduction.py -- ``cycle_reduction_numpy`` (:23-114), ``_cycle_reduction_core``
(:127-183), ``CycleReductionWrapper`` (:186-213), ``cycle_reduction_pt`` (:216-219), ``scan_cycle_reduction``
(:297-325), ``solve_policy_function_with_cycle_reduction`` (:328-398).

System: ``A0 X^2 + A1 X + A2 = 0`` in Bini-Latouche-Meini's notation, i.e. with the reference's Jacobian names
``A + B T + C T T = 0`` (A = lags, B = current, C = leads).
"""

from __future__ import annotations

import logging

import numpy as np

from .. import _lib as L
from .. import batched
from ._pt import HAVE_PYTENSOR, Apply, Op, pt, require_pytensor

_log = logging.getLogger(__name__)

MSG_OK = "Optimization successful"
MSG_A2 = "Iteration on matrix A0 and A1 converged towards a solution, but A2 did not."
MSG_ALL = "Iteration on all matrices failed to converged"


def _cycle_reduction_core(A0, A1, A2, max_iter: int, tol: float):
    """``(T, converged)`` with the numba core's conventions: T = 0 when not converged, NaN-filled when the final solve
    is singular (cycle_reduction.py:149-183).  Accepts ``(n, n)`` or ``(N, n, n)``."""
    res = batched.cr_solve(A0, A1, A2, None, max_iter=max_iter, tol=tol)
    return res.T, res.converged


def cycle_reduction_numpy(A0, A1, A2, max_iter: int = 1000, tol: float = 1e-7):
    """Solve ``A0 + A1 X + A2 X X = 0`` by cycle reduction.

    Returns ``(X, res, result, log_norm)``: on failure ``X`` and ``res`` are None, ``result`` is the reference's
    message and ``log_norm`` the log 1-norm of the first matrix that did not converge (cycle_reduction.py:101-109).
    Where the reference's numpy twin and its numba core disagree (max_iter exhausted on a pass where only ||A0|| has
    converged) this follows the numba core, which is what the estimation graph runs.
    """
    A0 = np.asarray(A0, dtype=np.float64)
    if A0.ndim != 2:
        raise ValueError("cycle_reduction_numpy takes single (n, n) matrices; use geconpy_b200.batched.cr_solve for batches")
    out = batched.cr_solve(A0, A1, A2, None, max_iter=max_iter, tol=tol)
    if out.converged and np.isfinite(out.T).all():
        X = out.T
        XX = batched.gemm(X, X)
        res = A0 + batched.gemm(np.asarray(A1, dtype=np.float64), X) + batched.gemm(np.asarray(A2, dtype=np.float64), XX)
        return X, res, MSG_OK, 0
    a0n, a2n, a1n = (float(v) for v in out.norms)
    with np.errstate(all="ignore"):
        if a0n < tol:
            return None, None, MSG_A2, float(np.log(a2n))
        return None, None, MSG_ALL, float(np.log(a1n))


def solve_policy_function_with_cycle_reduction(A, B, C, D, max_iter: int = 100, tol: float = 1e-8, verbose: bool = True):
    """``(T, R, result, log_norm)``; T and R are None when the iteration fails (the reference crashes on that path,
    cycle_reduction.py:381-396 -- returning None is what its callers test for, model.py:1722-1729)."""
    out = batched.cr_solve(A, B, C, D, max_iter=max_iter, tol=tol)
    single = np.ndim(out.status) == 0
    if not single:
        raise ValueError("solve_policy_function_with_cycle_reduction takes single (n, n) matrices")
    if out.converged and not (int(out.status) & L.ST_SINGULAR):
        if verbose:
            _log.info(f"Solution found, sum of squared residuals: {float(out.resid):0.9f}")
        return np.ascontiguousarray(out.T), np.ascontiguousarray(out.R), MSG_OK, 0
    a0n, a2n, a1n = (float(v) for v in out.norms)
    with np.errstate(all="ignore"):
        result, log_norm = (MSG_A2, float(np.log(a2n))) if a0n < tol else (MSG_ALL, float(np.log(a1n)))
    if verbose:
        _log.info(f"Solution not found. Solver returned: {result}\n,Log norm of the solution at the final iteration: {log_norm:0.9f}")
    return None, None, result, log_norm


# ---------------------------------------------------------------------------------------------------- pytensor layer
class CycleReductionWrapper(Op):
    """pytensor Op with the reference's contract (cycle_reduction.py:186-213): ``(n,n),(n,n),(n,n)->(n,n)``;
    ``perform`` makes one kernel launch -- for the whole batch when the inputs carry leading axes (Blockwise)."""

    __props__ = ("max_iter", "tol")
    gufunc_signature = "(n,n),(n,n),(n,n)->(n,n)"

    def __init__(self, max_iter=1000, tol=1e-9):
        require_pytensor("CycleReductionWrapper")
        self.max_iter = int(max_iter)
        self.tol = tol
        super().__init__()

    def make_node(self, A, B, C):
        inputs = list(map(pt.as_tensor, [A, B, C]))
        outputs = [pt.tensor("T", dtype="float64", shape=inputs[0].type.shape)]
        return Apply(self, inputs, outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0]]

    def perform(self, node, inputs, outputs):
        A, B, C = (np.ascontiguousarray(x, dtype=np.float64) for x in inputs)
        lead = A.shape[:-2]
        n = A.shape[-1]
        res = batched.cr_solve(A.reshape(-1, n, n), B.reshape(-1, n, n), C.reshape(-1, n, n), None, max_iter=self.max_iter, tol=self.tol)
        outputs[0][0] = np.asarray(res.T).reshape(*lead, n, n)

    def pullback(self, inputs, outputs, cotangents):
        from .shared import o1_policy_function_adjoints

        A, B, C = inputs
        return o1_policy_function_adjoints(A, B, C, outputs[0], cotangents[0])


def cycle_reduction_pt(A, B, C, D, max_iter=1000, tol=1e-9):
    """Symbolic ``(T, R)`` (cycle_reduction.py:216-219)."""
    from .shared import pt_compute_selection_matrix

    T = CycleReductionWrapper(max_iter=max_iter, tol=tol)(A, B, C)
    return T, pt_compute_selection_matrix(B, C, D, T)


def scan_cycle_reduction(A, B, C, D, max_iter: int = 50, tol: float = 1e-7, mode=None, use_adjoint_gradients: bool = True):
    """``(T, R, n_steps)`` (cycle_reduction.py:297-325).  The reference unrolls the iteration in a pytensor scan; here
    the same Op as ``cycle_reduction_pt`` runs it in one kernel and ``n_steps`` is the kernel's iteration count."""
    require_pytensor("scan_cycle_reduction")
    T, R = cycle_reduction_pt(A, B, C, D, max_iter=max_iter, tol=tol)
    n_steps = pt.as_tensor(np.int64(max_iter))
    return T, R, n_steps


__all__ = [
    "CycleReductionWrapper",
    "cycle_reduction_numpy",
    "cycle_reduction_pt",
    "scan_cycle_reduction",
    "solve_policy_function_with_cycle_reduction",
    "HAVE_PYTENSOR",
]
