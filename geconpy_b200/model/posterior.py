"""Posterior- / prior-batched statistics of the state space (SURVEY.md section 8(f) rank 4 leftovers):

* ``sample_autocorrelation_matrices``  gEconpy/model/statespace.py:1217-1303 (``DSGEStateSpace.sample_autocorrelation_matrices``):
  model-implied autocorrelation matrices for EVERY draw of a parameter population.  The reference builds one pytensor graph
  (``solve_discrete_lyapunov`` -> ``scan`` of ``T_step`` powers) and maps it over the posterior with
  ``pm.compute_deterministics``; here it is three batched launches for the whole population: cycle reduction (T, R),
  ``dlyap_kernel`` (Sigma) and ``propagate_kernel`` (``T_step^k Sigma`` for k = 0 .. n_lags).
* ``data_from_prior``  gEconpy/model/statespace.py:1324-1429: artificial data from prior-predictive trajectories -- prior draws,
  one unconditional trajectory per draw (``x_t = T x_{t-1} + R eps_t`` from the stationary distribution, observation noise
  added), one draw picked as "the truth", entries knocked out at random.  PyMC-free: the prior is the box of the model spec
  (the same one ``prior_solvability_check`` samples), the population is solved and simulated on the GPU.

Both take a configured ``BatchedStateSpace``; neither has a CPU fallback (the batched entry points raise without the library).
"""

from __future__ import annotations

import numpy as np

from .. import batched

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

__all__ = ["sample_autocorrelation_matrices", "data_from_prior"]


def _split_params(ss, theta_full):
    """[N, n_param] -> (theta [N, n_theta], sigma_shock [N, k], sigma_err [N, n_err]) in the layout ``loglik`` takes."""
    m = ss.model
    th = np.ascontiguousarray(np.atleast_2d(theta_full), dtype=np.float64)
    if ss.constant_params:  # constant parameters filled in at their defaults, as BatchedStateSpace._expand_params does on the device
        if th.shape[1] != ss.n_param:
            raise ValueError(f"theta has {th.shape[1]} columns, expected {ss.n_param}: {ss.param_names}")
        full = np.tile(np.asarray(ss._const_row, dtype=np.float64), (th.shape[0], 1))
        full[:, np.asarray(ss._free_cols, dtype=np.int64)] = th
        th = full
    n_err = len(ss.measurement_error)
    want = m.n_theta + m.k + n_err
    if th.shape[1] != want:
        raise ValueError(f"expected {want} columns (free parameters, sigma_<shock>, error_sigma_<state>), got {th.shape[1]}")
    return th[:, : m.n_theta], th[:, m.n_theta : m.n_theta + m.k], th[:, m.n_theta + m.k :]


def _design(ss):
    """(Z [p, n], selector rows or the constant dense design matrix) in model variable order; augmented configurations are
    not supported here (their stationary covariance is singular and the reference special-cases them too)."""
    if ss.n_aug != ss.n_filter or getattr(ss, "_obs_lib", None) is not None:
        raise NotImplementedError("autocorrelations / prior data with state augmentation or observation equations are not built")
    m = ss.model
    Z = np.zeros((ss.p, m.n))
    Z[np.arange(ss.p), [m.var_names.index(v) for v in ss.observed_states]] = 1.0  # variable order: what solve() returns T, R in
    return Z


def sample_autocorrelation_matrices(ss, theta_full, n_lags: int = 10, observed: bool = False, lag_step: int = 1, return_status: bool = False):
    """Autocorrelation matrices of every draw: ``[N, n_lags + 1, s, s]`` with s = the model's variables (``observed=False``) or the
    observed states (``observed=True``: ``Z (T_step^k Sigma) Z'``, measurement-error variances added to the lag-0 variance), each
    normalised by ``outer(std, std)`` of the lag-0 matrix -- statespace.py:1266-1298.  ``lag_step``: model periods per lag.
    Draws whose solution failed (status != 0) come back NaN; ``return_status=True`` also returns the status words."""
    if not ss.configured:
        raise RuntimeError("call configure(...) first")
    if n_lags < 0 or lag_step < 1:
        raise ValueError("n_lags must be >= 0 and lag_step >= 1")
    theta, sig, err = _split_params(ss, theta_full)
    Z = _design(ss) if observed else None
    sol = ss.solve(theta)
    T, R, status = sol["T"], sol["R"], np.asarray(sol["status"]).copy()
    Sigma, st_l, _ = batched.dlyap(T, R, sig**2)
    status |= np.asarray(st_l)
    T_step = T
    for _ in range(lag_step - 1):
        T_step = batched.gemm(T_step, T)
    acov = batched.propagate(T_step, X0=Sigma, n_steps=n_lags + 1, start_at_x0=True)  # [N, n_lags + 1, n, n]
    if observed:
        acov = np.einsum("pi,nlij,qj->nlpq", Z, acov, Z)
        h = np.zeros((theta.shape[0], ss.p))
        if err.shape[1]:
            h[:, np.asarray(ss.err_pos, dtype=np.int64)] = err**2
        acov[:, 0] += h[:, :, None] * np.eye(ss.p)[None]
    with np.errstate(all="ignore"):
        std = np.sqrt(np.diagonal(acov[:, 0], axis1=-2, axis2=-1))
        out = acov / (std[:, None, :, None] * std[:, None, None, :])
    out[status != 0] = np.nan
    return (out, status) if return_status else out


def data_from_prior(ss, n_timesteps: int = 180, n_samples: int = 500, pct_missing: float = 0.0, random_seed=None, sigma_shock=0.01,
                    sigma_err=1e-3, method: str = "random", index=None):
    """``(true_parameters, data, prior)`` -- statespace.py:1324-1429 without PyMC.

    * prior draws: ``n_samples`` points of the spec's prior box (``method``: "random" | "lhs" | "sobol" | "halton", as
      ``prior_solvability_check``); parameters without a prior stay at their defaults; the shock / error scales are the given
      constants (scalars or vectors);
    * every draw is solved on the GPU and ONE unconditional trajectory per draw is simulated by ``propagate_kernel`` from a state
      drawn from its stationary distribution (Sigma from ``dlyap_kernel``), the reference's ``sample_unconditional_prior``;
    * one solvable draw is picked at random as the truth; its observed series plus measurement noise is ``data``
      (``[n_timesteps, p]``; a pandas DataFrame over ``index`` -- default the reference's quarterly index -- when pandas imports);
      ``pct_missing`` of the rows of each column are set to NaN, each column independently.

    ``prior``: dict(theta [n_samples, n_theta], status [n_samples], observed [n_samples, n_timesteps, p], param_idx)."""
    if not ss.configured:
        raise RuntimeError("call configure(...) first")
    if not 0.0 <= pct_missing < 1.0:
        raise ValueError("pct_missing must be in [0, 1)")
    from scipy.stats import qmc

    m = ss.model
    rng = np.random.default_rng(random_seed)
    Z = _design(ss)
    if index is not None:
        n_timesteps = len(index)
    names = list(m.param_names)
    theta = np.tile(m.theta_vector(), (n_samples, 1))
    bounds = dict(m.lin.spec.get("bounds", {}))
    cols = [names.index(p_) for p_ in bounds if p_ in names]
    if not cols:
        raise ValueError(f"model {m.name} has no priors (spec['bounds'] is empty): nothing to sample")
    lo = np.array([bounds[names[c]][0] for c in cols], dtype=np.float64)
    hi = np.array([bounds[names[c]][1] for c in cols], dtype=np.float64)
    if method == "random":
        u = rng.random((n_samples, len(cols)))
    else:
        engines = {"lhs": qmc.LatinHypercube, "sobol": qmc.Sobol, "halton": qmc.Halton}
        if method not in engines:
            raise ValueError(f"unknown sampling method {method!r}; expected one of {sorted(engines)} or 'random'")
        u = engines[method](d=len(cols), seed=rng).random(n_samples)
    theta[:, cols] = qmc.scale(u, lo, hi)
    sig = np.broadcast_to(np.asarray(sigma_shock, dtype=np.float64), (m.k,)).copy()
    n_err = len(ss.measurement_error)
    err = np.broadcast_to(np.asarray(sigma_err, dtype=np.float64), (n_err,)).copy()
    with np.errstate(all="ignore"):
        sol = ss.solve(theta)
    T, R, status = sol["T"], sol["R"], np.asarray(sol["status"]).copy()
    Sigma, st_l, _ = batched.dlyap(np.nan_to_num(T), np.nan_to_num(R), np.tile(sig**2, (n_samples, 1)))
    status |= np.asarray(st_l)
    ok = status == 0
    if not ok.any():
        raise RuntimeError("no prior draw has a stable solution: nothing to simulate")
    # x_0 ~ N(0, Sigma) per draw (eigen-factor: Sigma is only semi-definite), then x_t = T x_{t-1} + R eps_t
    x0 = np.zeros((n_samples, m.n, 1))
    for i in np.flatnonzero(ok):
        w, V = np.linalg.eigh(0.5 * (Sigma[i] + Sigma[i].T))
        x0[i, :, 0] = V @ (np.sqrt(np.clip(w, 0.0, None)) * rng.standard_normal(m.n))
    E = (rng.standard_normal((n_samples, n_timesteps, m.k, 1)) * sig[None, None, :, None])
    Tn, Rn = np.where(ok[:, None, None], T, 0.0), np.where(ok[:, None, None], R, 0.0)
    X = batched.propagate(Tn, Rn, E=E, X0=x0)  # [N, L, n, 1]
    X = np.asarray(X)[..., 0]
    obs = np.einsum("pi,nli->nlp", Z, X)
    if n_err:
        noise = np.zeros_like(obs)
        noise[:, :, np.asarray(ss.err_pos, dtype=np.int64)] = rng.standard_normal((n_samples, n_timesteps, n_err)) * err
        obs = obs + noise
    obs[~ok] = np.nan
    idx = int(rng.choice(np.flatnonzero(ok)))
    data = obs[idx].copy()
    if pct_missing > 0:
        n_missing = int(n_timesteps * pct_missing)
        for c in range(data.shape[1]):
            data[rng.choice(n_timesteps, size=n_missing, replace=False), c] = np.nan
    true_parameters = dict(zip(names, theta[idx]))
    true_parameters["param_idx"] = idx
    try:
        import pandas as pd

        if index is None:
            index = pd.date_range(start="1980-01-01", periods=n_timesteps, freq="QS-OCT")
        data = pd.DataFrame(data, index=index, columns=list(ss.observed_states))
    except ImportError:  # pragma: no cover
        pass
    return true_parameters, data, dict(theta=theta, status=status, observed=obs, param_idx=idx)
