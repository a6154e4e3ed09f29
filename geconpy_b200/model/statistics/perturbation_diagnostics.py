"""``solvability_check`` on B200: every row of ``samples`` through steady state -> linearisation -> perturbation solve ->
Blanchard-Kahn -> residual norms in two kernel launches for the WHOLE frame (plus the exact BK kernel for uncertified
rows), instead of one Python call chain per draw over a fork pool.

Reference: gEconpy/model/statistics/perturbation_diagnostics.py -- ``_check_one_draw`` (:105-161),
``solvability_check`` (:362-450); norms as in gEconpy/model/perturbation.py:287-380.
"""

from __future__ import annotations

import numpy as np

from ... import _lib as L
from ... import batched
from ..compiled import CompiledModel


def solvability_check(
    model: CompiledModel,
    samples,
    *,
    cores: int = 1,
    solver: str = "cycle_reduction",
    steady_state_kwargs: dict | None = None,
    linearize_kwargs: dict | None = None,
    tol: float = 1e-8,
    max_iter: int = 100,
    norm_tol: float = 1e-8,
    progressbar: bool = True,
):
    """Same signature and return value as the reference: ``samples`` (a DataFrame whose columns are a subset of the
    model's free parameters) plus ``failure_step`` (None | "steady_state" | "perturbation" | "blanchard-kahn" |
    "deterministic_norm" | "stochastic_norm"), ``norm_deterministic``, ``norm_stochastic`` (NaN when not reached).
    ``cores`` / ``progressbar`` / the kwargs dictionaries are accepted for compatibility: the batch is one GPU pass."""
    if solver not in ("cycle_reduction", "gensys"):
        raise NotImplementedError(f"solver={solver!r}")
    unknown = [c for c in samples.columns if c not in model.param_names]
    if unknown:
        raise ValueError(f"samples has columns that are not free parameters of {model.name}: {unknown}")
    N = len(samples)
    theta = np.tile(model.theta_vector(), (N, 1))
    for c in samples.columns:
        theta[:, model.param_names.index(c)] = samples[c].to_numpy(dtype=np.float64)
    A, B, C, D, _xss, st_j = model.jacobian(theta)
    lead = model.permuted_lead_var_idx
    res = batched.cr_solve(A, B, C, D, max_iter=max_iter, tol=tol, lead_idx=lead, solvability_norms=True, trunc_tol=tol)
    status = res.status | st_j
    nu, status = batched.bk_count(A, B, C, lead, status=status, skip_mask=L.ST_BK_CERTIFIED | L.ST_JAC_NONFINITE, n_unstable=res.n_unstable)
    nd, ns = res.solv_norms[:, 0].copy(), res.solv_norms[:, 1].copy()
    step = np.full(N, None, dtype=object)
    reached = np.ones(N, dtype=bool)

    def mark(mask, name):
        nonlocal reached
        hit = mask & reached
        step[hit] = name
        reached &= ~hit
        return hit

    ss_fail = mark((status & L.ST_JAC_NONFINITE) != 0, "steady_state")
    pert_fail = mark((status & (L.ST_CR_NOT_CONVERGED | L.ST_CR_NAN | L.ST_SINGULAR)) != 0, "perturbation")
    bk_fail = mark((status & (L.ST_BK | L.ST_BK_INCONCLUSIVE)) != 0, "blanchard-kahn")
    not_reached = ss_fail | pert_fail | bk_fail
    nd[not_reached] = np.nan
    ns[not_reached] = np.nan
    with np.errstate(invalid="ignore"):
        mark(~(nd <= norm_tol) & ~not_reached, "deterministic_norm")
        mark(~(ns <= norm_tol) & ~not_reached, "stochastic_norm")
    out = samples.copy()
    import pandas as pd

    out["failure_step"] = pd.Series(step, index=out.index, dtype=object)  # None stays None (not NaN) on success
    out["norm_deterministic"] = nd
    out["norm_stochastic"] = ns
    return out
