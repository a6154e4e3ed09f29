"""Posterior-batched moments, impulse responses and simulations against golden vectors produced by the
REFERENCE's own functions (tests/golden/make_moment_goldens.py -> ref_moments.npz)."""

from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import model

pytestmark = pytest.mark.gpu

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "ref_moments.npz")


def _stack(name, key):
    return np.stack([GOLD[f"{name}/{d}/{key}"] for d in range(3)])


def _fake(name):
    mod = model(name)
    return SimpleNamespace(var_names=mod.var_names, shock_names=mod.shock_names), mod


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_stationary_covariance_and_autocovariance_match_the_reference(name):
    from geconpy_b200.model.statistics.covariance import (
        _compute_autocovariance_matrix, autocorrelation_matrix, autocovariance_matrix, stationary_covariance_matrix)

    fake, mod = _fake(name)
    T, R = _stack(name, "T"), _stack(name, "R")
    for key, kw in (("Sigma_std", dict(shock_std=0.01)), ("Sigma_cov", dict(shock_cov_matrix=GOLD[f"{name}/Qfull"]))):
        S, st = stationary_covariance_matrix(fake, T, R, return_status=True, **kw)
        ref = _stack(name, key)
        assert (st == 0).all()
        # scipy's bilinear solve and the doubling iteration both carry ~cond * eps error as rho(T) -> 1
        assert np.abs(S - ref).max() <= 1e-9 * np.abs(ref).max()
    S = _stack(name, "Sigma_std")
    for corr, key in ((False, "acov"), (True, "acorr")):
        ac = _compute_autocovariance_matrix(T, S, n_lags=6, correlation=corr)
        ref = _stack(name, key)
        assert ac.shape == ref.shape
        np.testing.assert_allclose(ac, ref, rtol=1e-11, atol=1e-13 * np.abs(ref).max())
    ac2 = autocovariance_matrix(fake, T, R, shock_std=0.01, n_lags=6)
    assert np.abs(ac2 - _stack(name, "acov")).max() <= 1e-9 * np.abs(S).max()
    assert np.abs(autocorrelation_matrix(fake, T, R, shock_std=0.01, n_lags=3)[:, 0].diagonal(axis1=-2, axis2=-1) - 1.0).max() <= 1e-12
    sd = stationary_covariance_matrix(fake, T, R, shock_std_dict={s: 0.01 for s in mod.shock_names})
    assert np.abs(sd - _stack(name, "Sigma_std")).max() <= 1e-9 * np.abs(S).max()


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_impulse_responses_and_simulation_match_the_reference(name):
    import torch

    from geconpy_b200.model.simulate import impulse_response_function, simulate

    fake, mod = _fake(name)
    T, R = _stack(name, "T"), _stack(name, "R")
    cases = [
        ("irf_unit", dict(simulation_length=25, shock_size=1.0)),
        ("irf_sizes_joint", dict(simulation_length=25, shock_size=GOLD[f"{name}/sizes"], return_individual_shocks=False)),
        ("irf_dict", dict(simulation_length=10, shock_size={mod.shock_names[-1]: 2.0})),
        ("irf_traj", dict(shock_trajectory=GOLD[f"{name}/traj"])),
        ("irf_cov", dict(simulation_length=8, shock_cov=GOLD[f"{name}/Qfull"], random_seed=9)),
    ]
    for key, kw in cases:
        out = impulse_response_function(fake, T, R, **kw)
        ref = _stack(name, key)
        assert out.shape == ref.shape, (key, out.shape, ref.shape)
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-14 * max(1.0, np.abs(ref).max()), err_msg=key)
    sim = simulate(fake, T, R, n_simulations=3, simulation_length=20, shock_std=0.01, random_seed=5)
    ref = _stack(name, "sim")
    assert sim.shape == ref.shape
    np.testing.assert_allclose(sim, ref, rtol=1e-11, atol=1e-15)
    # device path: torch tensors in, torch tensors out, same numbers
    Td, Rd = torch.as_tensor(T, device="cuda"), torch.as_tensor(R, device="cuda")
    out_d = impulse_response_function(fake, Td, Rd, simulation_length=25, shock_size=1.0)
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), impulse_response_function(fake, T, R, simulation_length=25, shock_size=1.0))
    with pytest.raises(ValueError):
        impulse_response_function(fake, T, R, shock_size=1.0, shock_cov=np.eye(mod.k))
    sim_n = simulate(fake, T, R, n_simulations=2, simulation_length=5, shock_std=0.01, random_seed=1, per_draw_shocks=True)
    assert sim_n.shape == (3, 2, 5, mod.n) and not np.allclose(sim_n[0], sim_n[1])


def test_propagate_sizes(rng):
    from geconpy_b200 import batched as B

    for n, k, m, L in ((1, 1, 1, 3), (9, 2, 5, 7), (45, 13, 13, 10), (64, 8, 64, 4)):
        N = 3
        T = rng.standard_normal((N, n, n)) / np.sqrt(n)
        R = rng.standard_normal((N, n, k))
        E = rng.standard_normal((N, L, k, m))
        X0 = rng.standard_normal((N, n, m))
        out = B.propagate(T, R, E=E, X0=X0)
        for i in range(N):
            x = X0[i]
            for t in range(L):
                x = T[i] @ x + R[i] @ E[i, t]
                np.testing.assert_allclose(out[i, t], x, rtol=1e-11, atol=1e-12)


# ---------------------------------------------------------------------------- posterior-batched ACF, prior-predictive data
def _configured(name, observed, meas):
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    return BatchedStateSpace(CompiledModel(name)).configure(observed_states=observed, measurement_error=meas, tol=1e-13, max_iter=1000)


@pytest.mark.parametrize("name,observed,lag_step", [("rbc", ["Y", "C"], 1), ("rbc", ["Y", "C"], 4), ("full_nk", ["Y", "pi", "r_G"], 1)])
def test_sample_autocorrelation_matrices_matches_the_reference_formula(name, observed, lag_step):
    """statespace.py:1266-1298 restated with scipy on the oracle's T, R: Sigma = dlyap(T, R Q R'), T_step = T^lag_step,
    acov_k = T_step^k Sigma (Z . Z' and + H at lag 0 when observed), normalised by the lag-0 standard deviations."""
    import scipy.linalg as sla

    from helpers import SIGMA_ERR, SIGMA_SHOCK, draws
    from oracle import solvers as osol

    from geconpy_b200.model.posterior import sample_autocorrelation_matrices

    mod = model(name)
    ss = _configured(name, observed, observed)
    N, n_lags = 5, 6
    th = draws(mod, N, seed=11, width=0.02, valid=True)
    sig = np.full((N, mod.k), SIGMA_SHOCK) * (1.0 + 0.2 * np.arange(N))[:, None]
    err = np.full((N, len(observed)), SIGMA_ERR)
    full = np.hstack([th, sig, err])
    got_all, st = sample_autocorrelation_matrices(ss, full, n_lags=n_lags, lag_step=lag_step, return_status=True)
    got_obs = sample_autocorrelation_matrices(ss, full, n_lags=n_lags, lag_step=lag_step, observed=True)
    assert (st == 0).all() and got_all.shape == (N, n_lags + 1, mod.n, mod.n) and got_obs.shape == (N, n_lags + 1, len(observed), len(observed))
    Z = np.zeros((len(observed), mod.n))
    Z[np.arange(len(observed)), [mod.var_names.index(v) for v in observed]] = 1.0
    for i in range(N):
        A, B, C, D = mod.jacobians(th[i], mode="statespace")
        T = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-13)[0]
        T, R = mod.unpermute_policy(T, osol.selection_matrix(B, C, D, T))
        Sigma = sla.solve_discrete_lyapunov(T, R @ np.diag(sig[i] ** 2) @ R.T, method="direct")
        Ts = np.linalg.matrix_power(T, lag_step)
        acov = np.stack([np.linalg.matrix_power(Ts, k_) @ Sigma for k_ in range(n_lags + 1)])
        std = np.sqrt(np.diag(acov[0]))
        with np.errstate(all="ignore"):
            ref_all = acov / np.outer(std, std)[None]
        ok = np.isfinite(ref_all)  # variables with zero variance (0 / 0) are NaN on both sides
        assert np.array_equal(np.isfinite(got_all[i]), ok) and np.abs(got_all[i][ok] - ref_all[ok]).max() <= 1e-8
        aco = Z @ acov @ Z.T
        aco[0] = Z @ Sigma @ Z.T + np.diag(err[i] ** 2)
        so = np.sqrt(np.diag(aco[0]))
        assert np.abs(got_obs[i] - aco / np.outer(so, so)[None]).max() <= 1e-8
    assert np.allclose(np.diagonal(got_obs[:, 0], axis1=-2, axis2=-1), 1.0)


def test_data_from_prior_is_reproducible_and_consistent_with_the_likelihood():
    """statespace.py:1324-1429: prior draws, one unconditional trajectory each, one of them the truth.  Same seed -> same data; the
    requested share of every column is missing; the truth's own likelihood of its data is finite and beats a clearly wrong draw."""
    from geconpy_b200.model.posterior import data_from_prior

    ss = _configured("rbc", ["Y", "C"], ["Y", "C"])
    truth, data, prior = data_from_prior(ss, n_timesteps=120, n_samples=64, pct_missing=0.1, random_seed=7)
    truth2, data2, prior2 = data_from_prior(ss, n_timesteps=120, n_samples=64, pct_missing=0.1, random_seed=7)
    D, D2 = np.asarray(data), np.asarray(data2)
    assert D.shape == (120, 2) and np.array_equal(np.isnan(D), np.isnan(D2)) and np.array_equal(np.nan_to_num(D), np.nan_to_num(D2))
    assert truth == truth2 and list(getattr(data, "columns", ["Y", "C"])) == ["Y", "C"]
    assert (np.isnan(D).sum(axis=0) == 12).all()
    ok = prior["status"] == 0
    assert ok[prior["param_idx"]] and ok.sum() >= 32 and prior["observed"].shape == (64, 120, 2)
    assert np.isfinite(prior["observed"][ok]).all() and np.isnan(prior["observed"][~ok]).all()
    m = ss.model
    th_true = np.array([truth[p_] for p_ in m.param_names])
    full = np.hstack([prior["theta"], np.full((64, m.k), 0.01), np.full((64, 2), 1e-3)])
    ll, st = ss.loglik(full, D)
    i = prior["param_idx"]
    assert np.array_equal(prior["theta"][i], th_true) and st[i] == 0 and np.isfinite(ll[i])
    finite = np.isfinite(ll)
    assert ll[i] >= np.median(ll[finite])  # the data-generating draw explains its own data better than a typical prior draw
