"""``gEconpy.solvers.backward_looking`` on B200 (models without leads: C == 0).

Reference: gEconpy/solvers/backward_looking.py -- ``solve_backward_policy`` (:8-24), ``solve_backward_shock_matrix``
(:46-78), ``solve_policy_function_with_backward_direct`` (:102-133) and their ``_pt`` twins.  ``A + B T = 0`` and
``B R + D = 0`` are solved by the batched partial-pivoting solve kernel (``gecon_solve_*`` / ``gecon_cr_solve_*`` with
C = NULL); singular B gives a NaN-filled result, never an exception.
"""

from __future__ import annotations

import numpy as np

from .. import batched
from ._pt import pt, require_pytensor


def solve_backward_policy(A, B):
    """``T = solve(-B, A)``: the transition matrix of a purely backward-looking model."""
    X, _st = batched.solve(-np.asarray(B, dtype=np.float64), np.asarray(A, dtype=np.float64))
    return X


def solve_backward_shock_matrix(B, D):
    """``R = -solve(B, D)``."""
    X, _st = batched.solve(np.asarray(B, dtype=np.float64), np.asarray(D, dtype=np.float64))
    return -X


def solve_policy_function_with_backward_direct(A, B, C, D):
    """``(T, R)`` for a model with no forward-looking variables; ``C`` is accepted for signature parity and must be
    (numerically) zero."""
    if C is not None and np.abs(np.asarray(C)).max(initial=0.0) > 1e-12:
        raise ValueError("backward-direct solver requires C == 0 (no forward-looking variables)")
    out = batched.cr_solve(A, B, None, D)
    return out.T, out.R


def solve_backward_policy_pt(A, B):
    require_pytensor("solve_backward_policy_pt")
    return pt.linalg.solve(-B, A)


def solve_backward_shock_matrix_pt(B, D):
    require_pytensor("solve_backward_shock_matrix_pt")
    return -pt.linalg.solve(B, D)


def solve_policy_function_with_backward_direct_pt(A, B, C, D):
    require_pytensor("solve_policy_function_with_backward_direct_pt")
    return solve_backward_policy_pt(A, B), solve_backward_shock_matrix_pt(B, D)
