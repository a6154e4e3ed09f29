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
    """Adjoint of ``A + B T + C T T = 0`` (an n^2 x n^2 Kronecker solve, shared.py:12-71).  Gradients are a "next" row
    of the scope table and are not implemented in round 1."""
    raise NotImplementedError("policy-function adjoints are a 'next' row of the hot-path scope (SURVEY.md 8f, rank 3)")
