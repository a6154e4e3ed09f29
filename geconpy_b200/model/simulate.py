"""Posterior-batched impulse responses and simulations on B200.

Same names, argument meaning and shock-selection rules as gEconpy/model/simulate.py, with a leading draw axis on
``T[N, n, n]`` / ``R[N, n, k]`` and plain arrays instead of xarray:

* ``impulse_response_function`` (simulate.py:201-318) -> ``[N, n_shocks, time, n]`` (individual shocks) or ``[N, time, n]``
* ``simulate``                  (simulate.py:320-412) -> ``[N, n_simulations, time, n]``

The recursion ``out[0] = R e_0, out[t] = T out[t-1] + R e_t`` (``_simulate_linear_system``, simulate.py:171-183) runs in
``gecon_propagate_*``; the shock panels are built on the host exactly as the reference builds them (same RNG calls).
"""

from __future__ import annotations

import numpy as np

from .. import batched
from .statistics.covariance import build_Q_matrix

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _names(model):
    return list(model.var_names), list(model.shock_names)


def _shock_vector(size, shock_names):
    n = len(shock_names)
    if size is None:
        return np.ones(n)
    if isinstance(size, (int, float)):
        return np.full(n, float(size))
    if isinstance(size, dict):
        return np.array([float(size.get(name, 0.0)) for name in shock_names])
    arr = np.asarray(size, dtype=float)
    if arr.shape != (n,):
        raise ValueError(f"shock_size array must have shape ({n},); got {arr.shape}.")
    return arr


def _orthogonal_factor(cov):
    return np.linalg.cholesky(cov)


def _to_out_layout(out, squeeze_m):
    # kernel layout [N, L, n, m] -> [N, m, L, n]
    perm = out.permute(0, 3, 1, 2) if torch is not None and isinstance(out, torch.Tensor) else np.transpose(out, (0, 3, 1, 2))
    return perm[:, 0] if squeeze_m else perm


def impulse_response_function(model, T, R, simulation_length=40, shock_size=None, shock_cov=None, shock_trajectory=None,
                              return_individual_shocks=None, orthogonalize_shocks=False, random_seed=None):
    """Impulse responses of every draw.  Shock options as in the reference: at most one of ``shock_size`` (float, array,
    dict of selected shocks), ``shock_cov`` (one draw e_0 ~ N(0, cov)), ``shock_trajectory`` ([time, k])."""
    given = [name for name, v in (("shock_size", shock_size), ("shock_cov", shock_cov), ("shock_trajectory", shock_trajectory)) if v is not None]
    if len(given) > 1:
        raise ValueError(f"Only one of {', '.join(given)} may be specified, got {len(given)}.")
    rng = np.random.default_rng(random_seed)
    _, shock_names = _names(model)
    k = len(shock_names)
    mode = "trajectory" if shock_trajectory is not None else "cov" if shock_cov is not None else "size"
    selected = shock_names
    if mode == "size" and isinstance(shock_size, dict):
        if len(shock_size) == 0:
            raise ValueError("Shock size cannot be empty.")
        unknown = set(shock_size) - set(shock_names)
        if unknown:
            raise ValueError(f"shock_size dict contains unknown shock names: {unknown}")
        selected = [s for s in shock_names if s in shock_size]
    idxs = [shock_names.index(s) for s in selected]
    # base trajectory [time, k]
    if mode == "trajectory":
        base = np.asarray(shock_trajectory, dtype=float)
        if base.ndim != 2 or base.shape[1] != k:
            raise ValueError(f"shock_trajectory must have shape (T, {k}); got {base.shape}.")
        simulation_length = base.shape[0]
    elif mode == "cov":
        Q = np.asarray(shock_cov, dtype=float)
        if Q.shape != (k, k):
            raise ValueError(f"shock_cov must be ({k}, {k}); got {Q.shape}.")
        base = np.zeros((simulation_length, k))
        base[0] = _orthogonal_factor(Q) @ rng.standard_normal(k)
    else:
        base = np.zeros((simulation_length, k))
        base[0] = _shock_vector(shock_size, shock_names)
    if return_individual_shocks is None:
        if mode == "size":
            individual = isinstance(shock_size, (int, float, dict)) or shock_size is None or np.asarray(shock_size).shape in ((), (k,))
        elif mode == "cov":
            individual = bool(np.allclose(shock_cov, np.diag(np.diag(shock_cov))))
        else:
            individual = False
    else:
        individual = bool(return_individual_shocks)
    if individual:
        E = np.zeros((simulation_length, k, len(idxs)))  # one column (trajectory) per selected shock
        for c, i in enumerate(idxs):
            E[:, i, c] = base[:, i]
    else:
        E = base[:, :, None]
    out = batched.propagate(T, R, E=E)
    return _to_out_layout(out, squeeze_m=not individual)


def simulate(model, T, R, n_simulations=1, simulation_length=40, shock_std_dict=None, shock_cov_matrix=None, shock_std=None,
             random_seed=None, per_draw_shocks=False):
    """Simulated trajectories of every draw: ``[N, n_simulations, time, n]``.  The shocks are drawn on the host with the
    reference's call (``rng.multivariate_normal(..., method="svd")``, simulate.py:388-393); by default every draw sees the
    SAME shock panels (common random numbers across the posterior); ``per_draw_shocks=True`` draws N independent sets."""
    rng = np.random.default_rng(random_seed)
    _, shock_names = _names(model)
    k = len(shock_names)
    Q = build_Q_matrix(shock_names, shock_std_dict, shock_cov_matrix, shock_std)
    N = T.shape[0] if T.ndim == 3 else 1
    size = (N, n_simulations, simulation_length) if per_draw_shocks else (n_simulations, simulation_length)
    eps = rng.multivariate_normal(mean=np.zeros(k), cov=Q, size=size, method="svd")  # [..., S, L, k]
    E = np.ascontiguousarray(np.moveaxis(eps, -3, -1))  # [..., L, k, S]
    out = batched.propagate(T, R, E=E)
    return _to_out_layout(out, squeeze_m=False)
