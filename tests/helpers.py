"""Shared test helpers: seeded draws, oracle Jacobians and simulated observations for the model-spec fixtures."""

from __future__ import annotations

import functools

import numpy as np

from oracle import solvers as osol
from oracle import statespace as oss
from oracle.model import OracleModel

# sigma of every shock (sigma_<shock>, statespace.py:255-258) and measurement-error sigma used in the tests
SIGMA_SHOCK = 0.01
SIGMA_ERR = 1e-3


@functools.lru_cache(maxsize=None)
def model(name: str) -> OracleModel:
    return OracleModel(name)


def draws(mod: OracleModel, N: int, seed: int = 0, width: float = 0.05, valid: bool = False) -> np.ndarray:
    """theta[N, n_theta]: defaults perturbed uniformly by +-width (relative), clipped into the GCN's prior bounds.
    Row 0 is the default parameter vector.  ``valid=True`` also keeps the discount factor and the AR coefficients
    strictly inside the unit interval and pins steady-state targets, so that (almost) every draw is solvable;
    ``valid=False`` deliberately leaves NaN steady states and unit roots in the population."""
    rng = np.random.default_rng(seed)
    th0 = mod.theta_vector()
    th = th0 * (1.0 + width * (2.0 * rng.random((N, th0.size)) - 1.0))
    bounds = mod.spec.get("bounds", {})
    for j, pname in enumerate(mod.param_names):
        if pname in bounds:
            lo, hi = bounds[pname]
            eps = 1e-6 * (hi - lo)
            th[:, j] = np.clip(th[:, j], lo + eps, hi - eps)
    if valid:
        for j, pname in enumerate(mod.param_names):
            if pname == "beta":
                th[:, j] = np.minimum(th[:, j], 0.999)
            elif pname.startswith("rho_"):
                th[:, j] = np.minimum(th[:, j], 0.99)
            elif pname in ("pi_bar", "phi_pi_obj"):
                th[:, j] = th0[j]
    th[0] = th0
    return th


def jacobian_batch(mod: OracleModel, thetas: np.ndarray):
    A, B, C, D = [], [], [], []
    for th in thetas:
        a, b, c, d = mod.jacobians(th, mode="statespace")
        A.append(a), B.append(b), C.append(c), D.append(d)
    return tuple(np.ascontiguousarray(np.stack(x)) for x in (A, B, C, D))


def observed_idx(mod: OracleModel, observed=None, permuted=False) -> np.ndarray:
    observed = observed or mod.spec["observed_default"]
    idx = np.array([mod.var_names.index(v) for v in observed], dtype=np.int32)
    if permuted:
        idx = mod.inv_var_order[idx].astype(np.int32)
    return idx


def simulate_obs(mod: OracleModel, Tobs: int, observed=None, seed: int = 0, sigma_err: float = 0.0) -> np.ndarray:
    """Y[Tobs, p] simulated at the default parameters (gEconpy/model/simulate.py:171-183 recursion)."""
    th = mod.theta_vector()
    A, B, C, D = mod.jacobians(th, mode="statespace")
    T, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-12)
    assert conv
    R = osol.selection_matrix(B, C, D, T)
    T, R = mod.unpermute_policy(T, R)
    x = oss.simulate(T, R, np.full(mod.k, SIGMA_SHOCK), Tobs, seed=seed)
    Y = x[:, observed_idx(mod, observed)]
    if sigma_err > 0:
        Y = Y + np.random.default_rng(seed + 1).standard_normal(Y.shape) * sigma_err
    return np.ascontiguousarray(Y)


def rel_fro(X, Xref):
    den = np.linalg.norm(Xref)
    return np.linalg.norm(X - Xref) / (den if den > 0 else 1.0)
