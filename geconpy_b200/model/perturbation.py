"""``gEconpy.model.perturbation`` on B200: linearisation entry point and Blanchard-Kahn checks.

Reference: gEconpy/model/perturbation.py -- ``linearize_model`` (:29-198), ``check_bk_condition`` (:508-583),
``check_bk_condition_pt`` (:586-625), ``compute_bk_eigenvalues(_pt)`` (:412-505).
"""

from __future__ import annotations

import logging

import numpy as np

from .. import _lib as L
from .. import batched
from .compiled import CompiledModel

_log = logging.getLogger(__name__)
_FLOAT_ZERO_TOL = 1e-8


def linearize_model(spec, log_linearize: bool = True, not_loglin_variables=()):
    """Model spec -> ``([A, B, C, D] evaluator, eq_order, var_order)``.

    The reference's ``linearize_model(variables, equations, shocks, ...)`` returns pytensor graph nodes for A, B, C, D
    (rows in ``eq_order``, columns in ``var_order``); its B200 counterpart returns a ``CompiledModel`` whose
    ``jacobian(theta)`` evaluates the same four matrices, in the same orderings, for a whole batch of draws with one
    generated CUDA kernel.  The symbolic front-end objects (TimeAwareSymbol lists) are replaced by this repo's model
    spec, which carries exactly those lists in serialised form (tests/golden/make_models.py)."""
    cm = CompiledModel(spec, log_linearize=log_linearize, not_loglin_variables=not_loglin_variables)
    return cm, cm.eq_order, cm.var_order


def check_bk_condition_pt(A, B, C, D, lead_var_idx):
    """``(bk_ok, n_forward, n_unstable)`` on the regularised pencil of the estimation graph (perturbation.py:586-625).
    Numeric arrays (numpy / torch CUDA), optionally with a leading draw axis; ``lead_var_idx`` are the (permuted)
    column positions of the structural lead variables (statespace.py:224-233, 769)."""
    lead = np.asarray(lead_var_idx, dtype=np.int32)
    nu, st = batched.bk_count(A, B, C, lead)
    ok = (st & (L.ST_BK | L.ST_BK_INCONCLUSIVE)) == 0
    return ok, int(lead.size), nu


def check_bk_condition(A, B, C, D, tol: float = 1e-8, verbose: bool = True, on_failure: str = "ignore", return_value="dataframe"):
    """Blanchard-Kahn check with the reference's signature (perturbation.py:508-583).

    The GPU kernel returns the COUNT of unstable generalized eigenvalues, not the eigenvalues themselves, so with
    ``return_value='dataframe'`` the frame holds one row (``n_forward``, ``n_unstable``, ``satisfied``) instead of the
    reference's per-eigenvalue table.  Forward-looking variables are the numerically non-zero columns of C, as in
    the reference's numpy variant (perturbation.py:441)."""
    if return_value not in ["dataframe", "bool", None]:
        raise ValueError(f'Unknown return type "{return_value}"')
    A, B, C = (np.ascontiguousarray(x, dtype=np.float64) for x in (A, B, C))
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > tol).astype(np.int32)
    nu, st = batched.bk_count(A, B, C, lead)
    n_forward, n_unstable = int(lead.size), int(nu)
    satisfied = (int(st) & (L.ST_BK | L.ST_BK_INCONCLUSIVE)) == 0
    msg = (
        f"Model solution has {n_unstable} eigenvalues greater than one in modulus and {n_forward} forward-looking variables."
        f"\nBlanchard-Kahn condition is{'' if satisfied else ' NOT'} satisfied."
    )
    if not satisfied and on_failure == "raise":
        raise ValueError(msg)
    if verbose:
        _log.info(msg)
    if return_value is None:
        return None
    if return_value == "bool":
        return bool(satisfied)
    import pandas as pd

    return pd.DataFrame({"n_forward": [n_forward], "n_unstable": [n_unstable], "satisfied": [bool(satisfied)]})


def compute_bk_eigenvalues(A, B, C, D, tol: float = 1e-8):
    raise NotImplementedError(
        "the B200 path counts the unstable eigenvalues (gecon_bk_count_*, matrix sign function) and never forms them; "
        "use check_bk_condition / check_bk_condition_pt"
    )


compute_bk_eigenvalues_pt = compute_bk_eigenvalues
