"""GPU parity for SURVEY 8f rank 2: cumulator state augmentation, steady-state observation intercepts, mixed-frequency
missing patterns and per-draw design matrices, against the oracle restatement of statespace.py:260-388,598-723."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import SIGMA_ERR, SIGMA_SHOCK, draws, model
from oracle import solvers as osol
from oracle import statespace as oss

pytestmark = pytest.mark.gpu
TOL_LL = 1e-7  # north-star tolerance on the log-likelihood


@pytest.fixture(scope="module")
def B():
    from geconpy_b200 import batched

    return batched


def _aggregated_data(mod, observed, ta, period, intercept, Tobs, seed, dense=False):
    """Simulate the model at its defaults, aggregate the way the design matrix says, add the intercept and noise, and
    leave NaN where a low-frequency series is not observed (prepare_mixed_frequency_data's "last" placement).
    ``dense``: the aggregate is reported every period (rolling window) instead.  A series with an intercept must be
    dense: the filter does not mask d (y_hat = d + Zm a), so a missing entry with d != 0 costs d^2 / jitter
    on both sides and the absolute tolerance on ll would be meaningless."""
    th = mod.theta_vector()
    A, Bm, C, D = mod.jacobians(th, mode="statespace")
    T, conv, _ = osol.cycle_reduction_core(A, Bm, C, max_iter=1000, tol=1e-12)
    assert conv
    R = osol.selection_matrix(Bm, C, D, T)
    T, R = mod.unpermute_policy(T, R)
    x = oss.simulate(T, R, np.full(mod.k, SIGMA_SHOCK), Tobs, seed=seed)
    xs = mod.steady_state(th)
    rng = np.random.default_rng(seed + 1)
    Y = np.full((Tobs, len(observed)), np.nan)
    for i, name in enumerate(observed):
        j = mod.var_names.index(name)
        m = ta.get(name)
        base = np.log(xs[j]) if name in intercept else 0.0
        for t in range(Tobs):
            if m in ("sum", "mean"):
                if t % period != period - 1 and not dense:
                    continue
                w = x[max(0, t - period + 1) : t + 1, j].sum()
                w, b = (w / period, base) if m == "mean" else (w, period * base)
            elif m in ("first", "last"):
                if t % period != (0 if m == "first" else period - 1) and not dense:
                    continue
                w, b = x[t, j], base
            else:
                w, b = x[t, j], base
            Y[t, i] = w + b + SIGMA_ERR * rng.standard_normal()
    return Y


CASES = [
    ("rbc", ["Y", "C"], {"Y": "sum"}, 4, ["Y", "C"], True),
    ("rbc", ["Y", "C"], {"Y": "mean", "C": "last"}, 3, [], False),
    ("rbc", ["C", "Y"], {}, 4, ["Y"], True),
    ("full_nk", ["Y", "pi", "r_G"], {"Y": "sum", "pi": "mean", "r_G": "last"}, 4, [], False),
    ("full_nk", ["Y", "pi", "r_G"], {"Y": "sum"}, 2, ["pi"], False),
    ("full_nk", ["Y", "pi", "r_G"], {"Y": "sum", "r_G": "first"}, 4, ["Y", "r_G"], True),
]


@pytest.mark.parametrize("name,observed,ta,period,intercept,dense", CASES)
@pytest.mark.parametrize("reduce_state", [True, False])
def test_augmented_pipeline_matches_oracle(name, observed, ta, period, intercept, dense, reduce_state):
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model(name)
    cm = CompiledModel(name)
    ss = BatchedStateSpace(cm).configure(
        observed_states=observed, measurement_error=observed, tol=1e-9, max_iter=200, reduce_state=reduce_state,
        temporal_aggregation=ta, aggregation_period=period, ss_obs_intercept=intercept, chunk=16,
    )  # fmt: skip
    n_cum = sum(m in ("sum", "mean") for m in ta.values()) * (period - 1)
    assert ss.n_aug == ss.n_filter + n_cum
    N = 20
    # with an intercept the data pin the steady state: keep the draws close so that ll stays O(1e3) and the absolute
    # tolerance is meaningful
    th = draws(mod, N, seed=51, width=0.002 if intercept else 0.03, valid=True)
    Y = _aggregated_data(mod, observed, ta, period, intercept, 48, seed=8, dense=dense)
    sig = np.full((N, mod.k), SIGMA_SHOCK)
    err = np.full((N, len(observed)), SIGMA_ERR)
    ll, st = ss.loglik(np.hstack([th, sig, err]), Y)
    n_ok = 0
    for i in range(N):
        ref = oss.loglik_augmented(mod, th[i], Y, observed, sig[i], err[i], temporal_aggregation=ta, aggregation_period=period,
                                   ss_obs_intercept=intercept, tol=1e-9, max_iter=200)  # fmt: skip
        if ref["ok"] and np.isfinite(ref["ll"]):
            n_ok += 1
            assert st[i] == 0, (i, st[i])
            assert abs(ll[i] - ref["ll"]) <= TOL_LL, (i, ll[i], ref["ll"])
        else:
            assert np.isneginf(ll[i]) and st[i] != 0
    assert n_ok >= N // 2


@pytest.mark.parametrize("mask_intercept", [True, False])
def test_mixed_frequency_data_with_a_steady_state_intercept(mask_intercept):
    """``ss_obs_intercept`` together with NaN data ("last" aggregation: three of four quarters missing) through
    the whole pipeline, in both conventions for the intercept at missing entries (``configure(mask_intercept=...)``).  In the
    unmasked convention every missing entry is scored as d_i^2 / jitter: ll ~ -1e8, compared relatively."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    name, observed, ta, period, intercept = "rbc", ["Y", "C"], {"Y": "last"}, 4, ["Y", "C"]
    mod = model(name)
    cm = CompiledModel(name)
    ss = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-9, max_iter=200, temporal_aggregation=ta,
                                         aggregation_period=period, ss_obs_intercept=intercept, mask_intercept=mask_intercept)  # fmt: skip
    N = 12
    th = draws(mod, N, seed=52, width=0.002, valid=True)
    Y = _aggregated_data(mod, observed, ta, period, intercept, 48, seed=9, dense=False)
    assert np.isnan(Y[:, 0]).sum() == 36 and not np.isnan(Y[:, 1]).any()
    sig = np.full((N, mod.k), SIGMA_SHOCK)
    err = np.full((N, len(observed)), SIGMA_ERR)
    ll, st = ss.loglik(np.hstack([th, sig, err]), Y)
    ll_g, grad, st_g = ss.loglik_and_grad(np.hstack([th, sig, err]), Y)
    for i in range(N):
        ref = oss.loglik_augmented(mod, th[i], Y, observed, sig[i], err[i], temporal_aggregation=ta, aggregation_period=period,
                                   ss_obs_intercept=intercept, tol=1e-9, max_iter=200, mask_intercept=mask_intercept)  # fmt: skip
        assert ref["ok"] and st[i] == 0 and st_g[i] == 0
        tol = max(TOL_LL, abs(ref["ll"]) * 1e-12)
        assert abs(ll[i] - ref["ll"]) <= tol and abs(ll_g[i] - ref["ll"]) <= tol, (i, ll[i], ll_g[i], ref["ll"])
    assert np.isfinite(grad).all()
    assert (np.abs(ll) < 1e5).all() if mask_intercept else (ll < -1e6).all()


def test_measurement_error_positions_follow_the_reference():
    """statespace.py:800-808: error variances fill positions 0..len(error_states)-1 of diag(H)."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("rbc_extended")
    observed = mod.spec["observed_default"][:3]
    cm = CompiledModel("rbc_extended")
    ss = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=[observed[2]], tol=1e-9, max_iter=200)
    th = draws(mod, 4, seed=3, width=0.01, valid=True)
    Y = _aggregated_data(mod, observed, {}, 1, [], 30, seed=2)
    sig = np.full((4, mod.k), SIGMA_SHOCK)
    ll, st = ss.loglik(np.hstack([th, sig, np.full((4, 1), 5e-3)]), Y)
    for i in range(4):
        ref = oss.loglik(mod, th[i], Y, observed, sig[i], np.array([5e-3, 0.0, 0.0]), tol=1e-9, max_iter=200)
        assert st[i] == 0 and abs(ll[i] - ref["ll"]) <= TOL_LL


def test_strided_solver_outputs_equal_dense_ones(B):
    """gecon_cr_args.t_ld / t_stride / r_stride: T, R land in the top-left block of a larger buffer, rest untouched."""
    import ctypes as C

    import torch

    from geconpy_b200 import _lib as L
    from helpers import jacobian_batch

    mod = model("rbc_extended")
    th = draws(mod, 6, seed=9, width=0.02, valid=True)
    A, Bm, Cm, D = jacobian_batch(mod, th)
    dense = B.cr_solve(A, Bm, Cm, D, max_iter=200, tol=1e-9)
    n, k, N, na = mod.n, mod.k, len(th), mod.n + 5
    dev = "cuda"
    tA, tB, tC, tD = (torch.as_tensor(x, device=dev) for x in (A, Bm, Cm, D))
    Tbig = torch.full((N, na, na), 7.0, dtype=torch.float64, device=dev)
    Rbig = torch.full((N, na, k), 7.0, dtype=torch.float64, device=dev)
    st = torch.zeros(N, dtype=torch.int32, device=dev)
    args = L.CrArgs(struct_size=C.sizeof(L.CrArgs), A=tA.data_ptr(), B=tB.data_ptr(), C=tC.data_ptr(), D=tD.data_ptr(), N=N, n=n, k=k,
                    max_iter=200, accumulate=0, tol=1e-9, resid_tol=0.0, T=Tbig.data_ptr(), R=Rbig.data_ptr(), status=st.data_ptr(),
                    t_stride=na * na, r_stride=na * k, t_ld=na)  # fmt: skip
    L.check(L.load_library().gecon_cr_solve_batched(C.byref(args), None), "gecon_cr_solve_batched")
    torch.cuda.synchronize()
    Tb, Rb = Tbig.cpu().numpy(), Rbig.cpu().numpy()
    assert np.array_equal(Tb[:, :n, :n], dense.T) and np.array_equal(Rb[:, :n], dense.R)
    assert np.all(Tb[:, n:, :] == 7.0) and np.all(Tb[:, :, n:] == 7.0) and np.all(Rb[:, n:] == 7.0)


@pytest.mark.parametrize("n,k,p", [(6, 2, 2), (12, 3, 3), (30, 4, 5)])
def test_per_draw_design_matrix(B, n, k, p):
    """z_stride = p n: one dense Z per draw (parameter-dependent observation equations, statespace.py:299-331)."""
    rng = np.random.default_rng(n)
    N, Tobs = 5, 40
    T = rng.standard_normal((N, n, n))
    for i in range(N):
        T[i] *= 0.85 / np.abs(np.linalg.eigvals(T[i])).max()
    R = rng.standard_normal((N, n, k))
    q = 0.5 + rng.random((N, k))
    h = 0.1 + rng.random((N, p))
    Z = rng.standard_normal((N, p, n)) * (rng.random((N, p, n)) < 0.4)
    d = 0.1 * rng.standard_normal((N, p))
    Y = rng.standard_normal((Tobs, p))
    Y[rng.random(Y.shape) < 0.15] = np.nan
    ll, st = B.kalman_loglik(T, R, q, Y, Z=Z, hdiag=h, d=d)
    ll_shared, _ = B.kalman_loglik(T, R, q, Y, Z=Z[0], hdiag=h, d=d)
    for i in range(N):
        ref = oss.kalman_loglik(Y, T[i], R[i], np.diag(q[i]), Z[i], np.diag(h[i]), d=d[i])
        assert st[i] == 0 and abs(ll[i] - ref) <= TOL_LL, (i, ll[i], ref)
        ref0 = oss.kalman_loglik(Y, T[i], R[i], np.diag(q[i]), Z[0], np.diag(h[i]), d=d[i])
        assert abs(ll_shared[i] - ref0) <= TOL_LL


# ------------------------------------------------------------------------------------------------ observation equations
def _obs_eq_ss(name, observed, meas, reduce_state=True, **kw):
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    return BatchedStateSpace(CompiledModel(name)).configure(
        observed_states=observed, measurement_error=meas, tol=1e-9, max_iter=200, reduce_state=reduce_state, chunk=16, **kw
    )


def test_observation_equation_equals_the_model_variable_formulation():
    """The property the reference tests (test_statespace.py:583-630): observing ``log(Y[])`` through an observation
    equation gives the likelihood of observing the log-linearised Y with a steady-state intercept."""
    mod = model("rbc")
    N = 12
    th = draws(mod, N, seed=13, width=0.002, valid=True)
    Y = _aggregated_data(mod, ["Y"], {}, 1, ["Y"], 40, seed=4, dense=True)
    full = np.hstack([th, np.full((N, mod.k), SIGMA_SHOCK), np.full((N, 1), SIGMA_ERR)])
    ll_eq, st_eq = _obs_eq_ss("rbc", ["Y_obs"], ["Y_obs"], observation_equations={"Y_obs": "log(Y[])"}).loglik(full, Y)
    ll_var, st_var = _obs_eq_ss("rbc", ["Y"], ["Y"], ss_obs_intercept=["Y"]).loglik(full, Y)
    assert (st_eq == 0).all() and (st_var == 0).all()
    assert np.abs(ll_eq - ll_var).max() <= TOL_LL


OBS_CASES = [
    ("rbc", ["dY", "C"], {"dY": "log(Y[]) - log(Y[-1])"}, {}, 4),
    ("rbc", ["C", "mix"], {"mix": "alpha * log(Y[]) + (1 - alpha) * log(K[-2]) - log(A[ss])"}, {}, 4),
    ("rbc", ["dY", "C"], {"dY": "log(Y[]) - log(Y[-1])"}, {"dY": "sum", "C": "mean"}, 4),
    ("full_nk", ["dY", "pi", "r_G"], {"dY": "100 * (log(Y[]) - log(Y[-1]))"}, {"dY": "mean"}, 2),
]


@pytest.mark.parametrize("name,observed,eqs,ta,period", OBS_CASES)
@pytest.mark.parametrize("reduce_state", [True, False])
def test_observation_equations_match_oracle(name, observed, eqs, ta, period, reduce_state):
    mod = model(name)
    ss = _obs_eq_ss(name, observed, observed, reduce_state, observation_equations=eqs, temporal_aggregation=ta, aggregation_period=period)
    N, Tobs = 12, 36
    th = draws(mod, N, seed=17, width=0.002, valid=True)  # intercepts depend on theta: keep ll O(1e2) for the absolute tolerance
    rng = np.random.default_rng(5)
    sig, err = np.full((N, mod.k), SIGMA_SHOCK), np.full((N, len(observed)), 5e-3)
    # data: draw from the ORACLE's state space at the default parameters so that the likelihood is well scaled
    ref0 = oss.loglik_augmented(mod, th[0], np.zeros((1, len(observed))), observed, sig[0], err[0], temporal_aggregation=ta,
                                aggregation_period=period, observation_equations=eqs, tol=1e-9, max_iter=200)  # fmt: skip
    Ta, Ra, Z, d = ref0["T_aug"], ref0["R_aug"], ref0["Z"], ref0["d"]
    x = np.zeros(Ta.shape[0])
    Y = np.zeros((Tobs, len(observed)))
    for t in range(Tobs):
        x = Ta @ x + Ra @ (SIGMA_SHOCK * rng.standard_normal(mod.k))
        Y[t] = d + Z @ x + 5e-3 * rng.standard_normal(len(observed))
    ll, st = ss.loglik(np.hstack([th, sig, err]), Y)
    n_ok = 0
    for i in range(N):
        ref = oss.loglik_augmented(mod, th[i], Y, observed, sig[i], err[i], temporal_aggregation=ta, aggregation_period=period,
                                   observation_equations=eqs, tol=1e-9, max_iter=200)  # fmt: skip
        if ref["ok"] and np.isfinite(ref["ll"]):
            n_ok += 1
            assert st[i] == 0, (i, st[i])
            assert abs(ll[i] - ref["ll"]) <= TOL_LL, (i, ll[i], ref["ll"])
        else:
            assert np.isneginf(ll[i]) and st[i] != 0
    assert n_ok >= N // 2


# ------------------------------------------------------------------------------------------------ full shock covariance
@pytest.mark.parametrize("n,k,p", [(6, 3, 2), (12, 4, 3), (30, 5, 4)])
def test_full_shock_covariance_kernel(B, n, k, p):
    """gecon_kalman_args.qfull: Q = state_cov used directly (statespace.py:245-249), shared and per draw."""
    rng = np.random.default_rng(100 + n)
    N, Tobs = 4, 30
    T = rng.standard_normal((N, n, n))
    for i in range(N):
        T[i] *= 0.85 / np.abs(np.linalg.eigvals(T[i])).max()
    R = rng.standard_normal((N, n, k))
    Lq = rng.standard_normal((N, k, k))
    Q = Lq @ Lq.transpose(0, 2, 1) + 0.1 * np.eye(k)
    h = 0.1 + rng.random((N, p))
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    Y = rng.standard_normal((Tobs, p))
    ll, st = B.kalman_loglik(T, R, None, Y, obs_idx=obs, hdiag=h, Q=Q)
    ll0, _ = B.kalman_loglik(T, R, None, Y, Z=Z, hdiag=h, Q=Q[0])
    ll_diag, _ = B.kalman_loglik(T, R, np.ones((N, k)), Y, obs_idx=obs, hdiag=h)
    ll_eye, _ = B.kalman_loglik(T, R, None, Y, obs_idx=obs, hdiag=h, Q=np.eye(k))
    assert np.abs(ll_diag - ll_eye).max() <= TOL_LL  # Q = I through both code paths (warp kernel vs CTA kernel)
    for i in range(N):
        assert st[i] == 0
        assert abs(ll[i] - oss.kalman_loglik(Y, T[i], R[i], Q[i], Z, np.diag(h[i]))) <= TOL_LL
        assert abs(ll0[i] - oss.kalman_loglik(Y, T[i], R[i], Q[0], Z, np.diag(h[i]))) <= TOL_LL


def test_full_shock_covariance_pipeline():
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("full_nk")
    observed = mod.spec["observed_default"]
    ss = BatchedStateSpace(CompiledModel("full_nk")).configure(
        observed_states=observed, measurement_error=observed, tol=1e-9, max_iter=200, full_shock_covariance=True, chunk=8
    )
    k, N = mod.k, 10
    assert ss.param_names[len(mod.param_names)] == "state_cov[0,0]" and len(ss.param_names) == len(mod.param_names) + k * k + len(observed)
    th = draws(mod, N, seed=23, width=0.02, valid=True)
    rng = np.random.default_rng(3)
    Lq = SIGMA_SHOCK * (np.eye(k) + 0.3 * rng.standard_normal((N, k, k)))
    Q = Lq @ Lq.transpose(0, 2, 1)
    err = np.full((N, len(observed)), SIGMA_ERR)
    from helpers import simulate_obs

    Y = simulate_obs(mod, 50, seed=3, sigma_err=SIGMA_ERR)
    ll, st = ss.loglik(np.hstack([th, Q.reshape(N, -1), err]), Y)
    n_ok = 0
    for i in range(N):
        ref = oss.loglik(mod, th[i], Y, observed, None, err[i], tol=1e-9, max_iter=200, Q_full=Q[i])
        if ref["ok"] and np.isfinite(ref["ll"]):
            n_ok += 1
            assert st[i] == 0 and abs(ll[i] - ref["ll"]) <= TOL_LL, (i, ll[i], ref["ll"])
        else:
            assert np.isneginf(ll[i])
    assert n_ok >= N // 2
    # the gradient path takes the full covariance too (round 2; checked in tests/test_gpu_gradient.py): same ll, finite gradient
    ll_g, grad, st_g = ss.loglik_and_grad(np.hstack([th, Q.reshape(N, -1), err]), Y)
    assert np.array_equal(st_g, st) and np.abs(ll_g[st == 0] - ll[st == 0]).max() <= TOL_LL and np.isfinite(grad).all()
