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
