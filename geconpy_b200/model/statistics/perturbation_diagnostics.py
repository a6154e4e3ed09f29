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
    # Everything between the parameter frame and the three result columns stays on the device: theta -> A, B, C, D (generated
    # kernel) -> cycle reduction + residual norms -> Blanchard-Kahn count, in chunks that bound the workspace; only
    # (status, two norms) per draw come back (16 bytes per draw instead of the 14.6 KB Jacobians of a medium model).
    L.require_device()
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    n, k = model.n, model.k
    lead = model.permuted_lead_var_idx
    chunk = min(N, 32768) or 1
    f64 = dict(dtype=torch.float64, device=dev)
    ws = [torch.empty((chunk, n, n), **f64) for _ in range(3)] + [torch.empty((chunk, n, k), **f64)]
    st_j = torch.empty((chunk,), dtype=torch.int32, device=dev)
    status = np.empty(N, dtype=np.int32)
    norms = np.empty((N, 2))
    stream = torch.cuda.current_stream(dev).cuda_stream
    for lo in range(0, N, chunk):
        cnt = min(chunk, N - lo)
        th_d = torch.as_tensor(theta[lo : lo + cnt], device=dev)
        A, B, C, D = (w[:cnt] for w in ws)
        model.jacobian_device(th_d, A, B, C, D, None, st_j[:cnt], stream)
        res = batched.cr_solve(A, B, C, D, max_iter=max_iter, tol=tol, lead_idx=lead, solvability_norms=True, trunc_tol=tol)
        st = res.status | st_j[:cnt]
        _nu, st = batched.bk_count(A, B, C, lead, status=st, skip_mask=L.ST_BK_CERTIFIED | L.ST_JAC_NONFINITE, n_unstable=res.n_unstable)
        status[lo : lo + cnt] = st.cpu().numpy()
        norms[lo : lo + cnt] = res.solv_norms.cpu().numpy()
    nd, ns = norms[:, 0].copy(), norms[:, 1].copy()
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


def prior_solvability_check(model: CompiledModel, n_samples: int, *, seed=None, param_subset=None, method: str = "lhs", hdi_prob: float = 0.99,
                            **kwargs):
    """Sample the free parameters over their prior bounds and check solvability (perturbation_diagnostics.py:526-579).

    The model spec carries each prior's bounds (``spec["bounds"]``, written by tests/golden/make_models.py from the GCN's
    priors), not the preliz distributions themselves, so the space-filling methods the reference recommends are available --
    ``"lhs"``, ``"sobol"``, ``"halton"``, ``"poisson_disk"`` (``scipy.stats.qmc`` over the bounds, as ``sample_uniform`` does,
    model/sampling.py:72-145) and ``"random"`` as independent uniforms on the same box -- while the inverse-CDF methods
    (``"*_ppf"``) need the distributions and raise.  ``hdi_prob`` is accepted for signature compatibility (the bounds are fixed
    in the spec).  Parameters outside ``param_subset`` (or without bounds) stay at their defaults.  ``**kwargs`` go to
    ``solvability_check``."""
    import pandas as pd

    from scipy.stats import qmc

    bounds = dict(model.lin.spec.get("bounds", {}))
    if param_subset is not None:
        unknown = sorted(set(param_subset) - set(bounds))
        if unknown:
            raise ValueError(f"param_subset contains names not found in model.param_priors: {unknown}")
        bounds = {k_: v for k_, v in bounds.items() if k_ in param_subset}
    if not bounds:
        raise ValueError(f"model {model.name} has no priors (spec['bounds'] is empty): nothing to sample")
    if method.endswith("_ppf"):
        raise NotImplementedError(f"method={method!r} needs the prior distributions; the model spec only carries their bounds")
    names = list(bounds)
    lo = np.array([bounds[p][0] for p in names], dtype=np.float64)
    hi = np.array([bounds[p][1] for p in names], dtype=np.float64)
    if method == "random":
        u = np.random.default_rng(seed).random((n_samples, len(names)))
    else:
        engines = {"lhs": qmc.LatinHypercube, "sobol": qmc.Sobol, "halton": qmc.Halton, "poisson_disk": qmc.PoissonDisk}
        if method not in engines:
            raise ValueError(f"unknown sampling method {method!r}; expected one of {sorted(engines)} or 'random'")
        u = engines[method](d=len(names), seed=seed).random(n_samples)
    samples = pd.DataFrame(qmc.scale(u, lo, hi), columns=names)
    return solvability_check(model, samples, **kwargs)
