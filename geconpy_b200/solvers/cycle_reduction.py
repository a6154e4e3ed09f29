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
def linalg_output_dtype(*dtypes) -> str:
    """Working dtype of a LAPACK-style Op given its input dtypes (the rule the reference applies in ``make_node``,
    cycle_reduction.py:196-199): float64 unless every input is single precision.  The kernels compute in fp64 either way;
    ``perform`` casts the result to this dtype."""
    kinds = [np.dtype(d) for d in dtypes]
    return "float32" if kinds and all(k == np.dtype("float32") for k in kinds) else "float64"


def _solve_nd(inputs, max_iter, tol, scan_semantics=False):
    """One batched launch for gufunc-style inputs ``(..., n, n)``: returns (T with the leading axes restored, result)."""
    A, B, C = (np.ascontiguousarray(x, dtype=np.float64) for x in inputs)
    lead, n = A.shape[:-2], A.shape[-1]
    res = batched.cr_solve(A.reshape(-1, n, n), B.reshape(-1, n, n), C.reshape(-1, n, n), None, max_iter=max_iter, tol=tol,
                           scan_semantics=scan_semantics)  # fmt: skip
    return np.asarray(res.T).reshape(*lead, n, n), res, lead


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
        o_dtype = linalg_output_dtype(*(inp.type.dtype for inp in inputs))
        outputs = [pt.tensor("T", dtype=o_dtype, shape=inputs[0].type.shape)]
        return Apply(self, inputs, outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0]]

    def perform(self, node, inputs, outputs):
        T, _res, _lead = _solve_nd(inputs, self.max_iter, self.tol)
        outputs[0][0] = np.asarray(T, dtype=node.outputs[0].type.dtype)

    def pullback(self, inputs, outputs, cotangents):
        return _linear_policy_jvp(inputs, outputs, cotangents)


def _linear_policy_jvp(inputs, outputs, cotangents):
    """Pullback shared by every solver Op (cycle_reduction.py:117-124): the adjoints of A, B, C given T_bar.  With
    symbolic inputs this is the ``PolicyAdjoint`` Op below, i.e. the Stein-equation kernel, not the n^2 x n^2 Kronecker
    graph of the reference."""
    A, B, C = inputs[:3]
    return list(PolicyAdjoint()(A, B, C, outputs[0], cotangents[0]))


class PolicyAdjoint(Op):
    """``o1_policy_function_adjoints`` (shared.py:12-71) as ONE Op backed by ``gecon_policy_adjoint_*``, so that a pytensor
    user differentiating through a solver Op reaches the GPU adjoint kernel: ``(n,n) x 5 -> (n,n) x 3``."""

    __props__ = ()
    gufunc_signature = "(n,n),(n,n),(n,n),(n,n),(n,n)->(n,n),(n,n),(n,n)"

    def __init__(self):
        require_pytensor("PolicyAdjoint")
        super().__init__()

    def make_node(self, A, B, C, T, T_bar):
        inputs = list(map(pt.as_tensor, [A, B, C, T, T_bar]))
        o_dtype = linalg_output_dtype(*(inp.type.dtype for inp in inputs))
        outputs = [pt.tensor(nm, dtype=o_dtype, shape=inputs[0].type.shape) for nm in ("A_bar", "B_bar", "C_bar")]
        return Apply(self, inputs, outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0]] * 3

    def perform(self, node, inputs, outputs):
        A, B, C, T, Tb = (np.ascontiguousarray(x, dtype=np.float64) for x in inputs)
        lead, n = A.shape[:-2], A.shape[-1]
        flat = [x.reshape(-1, n, n) for x in (A, B, C, T, Tb)]
        A_bar, B_bar, C_bar, _D, _st = batched.policy_adjoints(*flat)
        for slot, val in zip(outputs, (A_bar, B_bar, C_bar)):
            slot[0] = np.asarray(val, dtype=node.outputs[0].type.dtype).reshape(*lead, n, n)


def cycle_reduction_pt(A, B, C, D, max_iter=1000, tol=1e-9):
    """Symbolic ``(T, R)`` (cycle_reduction.py:216-219)."""
    from .shared import pt_compute_selection_matrix

    T = CycleReductionWrapper(max_iter=max_iter, tol=tol)(A, B, C)
    return T, pt_compute_selection_matrix(B, C, D, T)


class ScanCycleReduction(Op):
    """The scan twin (cycle_reduction.py:246-325) as one Op with TWO outputs, ``T`` and the int32 step count ``n_steps``
    that ``build_statespace_graph`` publishes as the ``n_cycle_steps`` Deterministic (statespace.py:1169-1171).  The
    kernel runs with ``scan_semantics``: only ||A0||_1 is tested, T is always solved for, and ``n_steps`` is the number of
    steps actually taken (the reference's scan keeps stepping as a no-op after convergence and reports the same count)."""

    __props__ = ("max_iter", "tol")
    gufunc_signature = "(n,n),(n,n),(n,n)->(n,n),()"

    def __init__(self, max_iter=50, tol=1e-7):
        require_pytensor("ScanCycleReduction")
        self.max_iter = int(max_iter)
        self.tol = tol
        super().__init__()

    def make_node(self, A, B, C):
        inputs = list(map(pt.as_tensor, [A, B, C]))
        o_dtype = linalg_output_dtype(*(inp.type.dtype for inp in inputs))
        outputs = [pt.tensor("T", dtype=o_dtype, shape=inputs[0].type.shape), pt.scalar("n_steps", dtype="int32")]
        return Apply(self, inputs, outputs)

    def infer_shape(self, fgraph, node, input_shapes):
        return [input_shapes[0], ()]

    def perform(self, node, inputs, outputs):
        T, res, lead = _solve_nd(inputs, self.max_iter, self.tol, scan_semantics=True)
        outputs[0][0] = np.asarray(T, dtype=node.outputs[0].type.dtype)
        n_it = np.asarray(res.n_iter, dtype=np.int32)
        outputs[1][0] = n_it.reshape(lead) if lead else np.int32(n_it.reshape(-1)[0])

    def pullback(self, inputs, outputs, cotangents):
        return _linear_policy_jvp(inputs, outputs, cotangents)


def scan_cycle_reduction(A, B, C, D, max_iter: int = 50, tol: float = 1e-7, mode=None, use_adjoint_gradients: bool = True):
    """``(T, R, n_steps)`` (cycle_reduction.py:297-325).  The reference unrolls the iteration in a pytensor scan; here one
    kernel runs it and ``n_steps`` is the kernel's own iteration count.  ``mode`` (a pytensor compilation mode for the
    scan) has nothing to act on and is accepted for signature compatibility; the adjoint-based pullback is the only one."""
    from .shared import pt_compute_selection_matrix

    require_pytensor("scan_cycle_reduction")
    T, n_steps = ScanCycleReduction(max_iter=max_iter, tol=tol)(A, B, C)
    return T, pt_compute_selection_matrix(B, C, D, T), n_steps


__all__ = [
    "CycleReductionWrapper",
    "ScanCycleReduction",
    "PolicyAdjoint",
    "linalg_output_dtype",
    "cycle_reduction_numpy",
    "cycle_reduction_pt",
    "scan_cycle_reduction",
    "solve_policy_function_with_cycle_reduction",
    "HAVE_PYTENSOR",
]
