"""Posterior-batched second moments on B200.

Same names and argument meaning as gEconpy/model/statistics/covariance.py, with a leading draw axis on ``T`` / ``R``:

* ``build_Q_matrix``                 (covariance.py:31-65)   shock covariance from a dict / a matrix / a common std
* ``stationary_covariance_matrix``   (covariance.py:68-128)  Sigma = T Sigma T' + R Q R' -> ``gecon_dlyap_*`` (Smith doubling)
* ``autocovariance_matrix`` / ``autocorrelation_matrix`` (covariance.py:131-237)  Gamma_h = T^h Sigma -> ``gecon_propagate_*``

``model`` only supplies names (``shock_names`` / ``var_names``: a ``CompiledModel`` or any object with those
attributes); solving the model is the caller's job (``BatchedStateSpace.solve``), so ``T`` and ``R`` are required.
Arrays in, arrays out (numpy on the host path, torch CUDA tensors on the device path); no DataFrame / xarray wrapping.
"""

from __future__ import annotations

import functools as ft

import numpy as np

from ... import batched

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _shock_names(model):
    return list(getattr(model, "shock_names"))


def _validate_shock_options(shock_std_dict, shock_cov_matrix, shock_std, shocks):
    """Exactly one way of describing the shocks (gEconpy/model/statistics/covariance.py `_validate_shock_options`)."""
    given = [x is not None for x in (shock_std_dict, shock_cov_matrix, shock_std)]
    if sum(given) != 1:
        raise ValueError("Exactly one of shock_std_dict, shock_cov_matrix, or shock_std should be provided")
    k = len(shocks)
    if shock_cov_matrix is not None and tuple(np.shape(shock_cov_matrix)) != (k, k):
        raise ValueError(f"Incorrect covariance matrix shape. Expected ({k}, {k}), found {np.shape(shock_cov_matrix)}")
    if shock_std_dict is not None:
        missing = [s for s in shocks if s not in shock_std_dict]
        extra = [s for s in shock_std_dict if s not in shocks]
        if missing:
            raise ValueError(f"If shock_std_dict is specified, it must give values for all shocks. The following shocks were not found: {', '.join(missing)}")
        if extra:
            raise ValueError(f"Unexpected shocks in shock_std_dict. The following names were not found among the model shocks: {', '.join(extra)}")


def build_Q_matrix(model_shocks, shock_std_dict=None, shock_cov_matrix=None, shock_std=None) -> np.ndarray:
    """(k, k) shock covariance matrix; same rules as the reference (covariance.py:31-65)."""
    shocks = list(model_shocks)
    _validate_shock_options(shock_std_dict, shock_cov_matrix, shock_std, shocks)
    k = len(shocks)
    if shock_cov_matrix is not None:
        return np.asarray(shock_cov_matrix, dtype=np.float64)
    if shock_std_dict is not None:
        Q = np.zeros((k, k))
        for name, value in shock_std_dict.items():
            i = shocks.index(name)
            Q[i, i] = float(value) ** 2
        return Q
    std = np.asarray(shock_std, dtype=np.float64)
    return np.eye(k) * std**2 if std.ndim == 0 else np.diag(std**2)


def _factor_Q(Q):
    """(L, q): R Q R' = (R L) diag(q) (R L)'; L = None when Q is diagonal (the dlyap kernel takes variances)."""
    Q = np.asarray(Q, dtype=np.float64)
    if np.array_equal(Q, np.diag(np.diag(Q))):
        return None, np.ascontiguousarray(np.diag(Q))
    return np.linalg.cholesky(Q), np.ones(Q.shape[0])


def stationary_covariance_matrix(model, T, R, shock_std_dict=None, shock_cov_matrix=None, shock_std=None, return_status=False):
    """Sigma[N, n, n] solving Sigma = T Sigma T' + R Q R' for every draw (``T[N, n, n]``, ``R[N, n, k]``)."""
    Q = build_Q_matrix(_shock_names(model), shock_std_dict, shock_cov_matrix, shock_std)
    Lq, q = _factor_Q(Q)
    if Lq is not None:
        R = R @ (torch.as_tensor(Lq, device=R.device) if torch is not None and isinstance(R, torch.Tensor) else Lq)
    Sigma, status, _ = batched.dlyap(T, R, q)
    return (Sigma, status) if return_status else Sigma


def _compute_autocovariance_matrix(T, Sigma, n_lags=5, correlation=True):
    """acov[N, n_lags, n, n] with acov[:, h] = T^h Sigma (/ outer(std, std) when ``correlation``)."""
    out = batched.propagate(T, X0=Sigma, n_steps=n_lags, start_at_x0=True)
    if correlation:
        is_t = torch is not None and isinstance(Sigma, torch.Tensor)
        diag = torch.diagonal(Sigma, dim1=-2, dim2=-1) if is_t else np.diagonal(Sigma, axis1=-2, axis2=-1)
        std = diag.sqrt() if is_t else np.sqrt(diag)
        norm = std[..., :, None] * std[..., None, :]
        out = out / (norm[:, None] if out.ndim == 4 else norm[None])
    return out


def autocovariance_matrix(model, T, R, shock_std_dict=None, shock_cov_matrix=None, shock_std=None, n_lags=10, correlation=False):
    """Autocovariances (or autocorrelations) of every draw: ``[N, n_lags, n, n]``."""
    Sigma = stationary_covariance_matrix(model, T, R, shock_std_dict, shock_cov_matrix, shock_std)
    return _compute_autocovariance_matrix(T, Sigma, n_lags=n_lags, correlation=correlation)


autocorrelation_matrix = ft.partial(autocovariance_matrix, correlation=True)
autocorrelation_matrix.__doc__ = autocovariance_matrix.__doc__
