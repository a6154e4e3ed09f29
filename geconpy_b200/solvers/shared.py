"""``gEconpy.solvers.shared`` on B200.

Reference: gEconpy/solvers/shared.py -- ``stabilize`` (:6-9), ``o1_policy_function_adjoints`` (:12-71),
``pt_compute_selection_matrix`` (:74-75).
"""

from __future__ import annotations

import numpy as np

from .. import batched
from ._pt import HAVE_PYTENSOR, pt


def _is_symbolic(*xs) -> bool:
    return HAVE_PYTENSOR and any(hasattr(x, "owner") and hasattr(x, "type") for x in xs)


def stabilize(x, jitter: float = 1e-16):
    """Add ``jitter`` to the diagonal."""
    if _is_symbolic(x):
        return x + pt.eye(x.shape[-1]) * jitter
    x = np.asarray(x, dtype=np.float64)
    return x + np.eye(x.shape[-1]) * jitter


def pt_compute_selection_matrix(B, C, D, T):
    """``R = -(C T + B)^-1 D``.  Symbolic inputs build the pytensor expression of the reference; numeric inputs
    (numpy / torch CUDA, optionally batched) run on the GPU: one DMMA product and one pivoted solve per draw."""
    if _is_symbolic(B, C, D, T):
        return -pt.linalg.solve(C @ T + B, D, assume_a="gen", check_finite=False)
    CT = batched.gemm(C, T)
    X, _st = batched.solve(CT + B, D)
    return -X


def o1_policy_function_adjoints(A, B, C, T, T_bar):
    """Adjoint of the matrix quadratic ``A + B T + C T T = 0``: returns ``[A_bar, B_bar, C_bar]`` (shared.py:12-71).

    The reference solves an n^2 x n^2 Kronecker system for the multipliers S; here they come from the equivalent Stein
    equation ``S = Q + G S T'`` by doubling (``gecon_policy_adjoint_*``).  Numeric inputs (numpy / torch CUDA, optionally
    with a leading draw axis) call the kernel directly; symbolic inputs go through the ``PolicyAdjoint`` Op, whose
    ``perform`` makes the same call -- so the pullback of every solver Op reaches the GPU kernel."""
    if _is_symbolic(A, B, C, T, T_bar):
        from .cycle_reduction import PolicyAdjoint

        return list(PolicyAdjoint()(A, B, C, T, T_bar))
    A_bar, B_bar, C_bar, _D_bar, _st = batched.policy_adjoints(A, B, C, T, T_bar)
    return [A_bar, B_bar, C_bar]
