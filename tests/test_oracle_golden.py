"""CPU tests: the oracle (oracle/) against every golden vector the reference holds for this path, plus independent
known-answer tests for the Kalman filter, whose parity the reference itself does not pin.

Fixtures come from tests/golden/make_goldens.py (run once in the build container against /root/reference).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import scipy.linalg
import scipy.stats

from helpers import SIGMA_SHOCK, model
from oracle import solvers as osol
from oracle import statespace as oss

GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------------------------------- A, B, C, D
@pytest.mark.parametrize("name", ["one_block_1_ss", "rbc_2_block_ss", "full_nk"])
def test_jacobians_match_reference_goldens(name):
    """tests/model/test_model.py:405-421 (test_linearize), atol 1e-8, original equation x variable order."""
    g = np.load(GOLD / "ref_linearization.npz")
    mod = model(name)
    th = mod.theta_vector(**{k: float(v) for k, v in zip(g[f"{name}/param_names"], g[f"{name}/param_values"]) if k in mod.defaults})
    mats = mod.jacobians(th, permuted=False, mode="model")
    for nm, M in zip("ABCD", mats):
        np.testing.assert_allclose(M, g[f"{name}/{nm}"], atol=1e-8, err_msg=f"{name} {nm}")


def test_permutation_is_consistent():
    mod = model("full_nk")
    th = mod.theta_vector()
    A, B, C, D = mod.jacobians(th, permuted=False)
    Ap, Bp, Cp, Dp = mod.jacobians(th, permuted=True)
    eo, vo = mod.eq_order, mod.var_order
    assert np.array_equal(Ap, A[eo][:, vo]) and np.array_equal(Cp, C[eo][:, vo]) and np.array_equal(Dp, D[eo])
    # lag columns / lead columns are contiguous blocks in solver order (perturbation.py:112-158)
    a_cols = np.flatnonzero(np.abs(Ap).sum(0))
    c_cols = np.flatnonzero(np.abs(Cp).sum(0))
    assert np.array_equal(a_cols, np.arange(a_cols[0], a_cols[-1] + 1))
    assert np.array_equal(c_cols, np.arange(c_cols[0], c_cols[-1] + 1))


# ------------------------------------------------------------------------------------------- cycle reduction
def _cr_cases():
    g = np.load(GOLD / "ref_cycle_reduction.npz")
    keys = sorted({"/".join(k.split("/")[:2]) for k in g.files if k.split("/")[1].isdigit()})
    return g, keys


def test_cycle_reduction_matches_the_reference_function():
    """The oracle's iteration vs the REFERENCE's cycle_reduction_numpy run from /root/reference (fixture)."""
    g, keys = _cr_cases()
    assert len(keys) >= 15
    n_ok = 0
    for key in keys:
        A, B, C, D = (g[f"{key}/{m}"] for m in "ABCD")
        X_ref = g[f"{key}/X"]
        T, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-9)
        X, res, msg, _ = osol.cycle_reduction_numpy(A, B, C, max_iter=1000, tol=1e-9)
        if np.isnan(X_ref).all():
            assert not conv and X is None
            continue
        n_ok += 1
        assert conv and str(g[f"{key}/msg"]) == msg == "Optimization successful"
        np.testing.assert_allclose(T, X_ref, atol=1e-12, rtol=1e-10)
        np.testing.assert_allclose(X, X_ref, atol=1e-12, rtol=1e-10)
        np.testing.assert_allclose(osol.selection_matrix(B, C, D, T), g[f"{key}/R"], atol=1e-12, rtol=1e-10)
        assert np.abs(A + B @ T + C @ T @ T).max() < 1e-8  # tests/model/test_perturbation.py:234-235
    assert n_ok >= 12


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_cycle_reduction_failure_tuple(name):
    g = np.load(GOLD / "ref_cycle_reduction.npz")
    mod = model(name)
    A, B, C, D = mod.jacobians(mod.theta_vector())
    X, res, msg, log_norm = osol.cycle_reduction_numpy(A, B, C, max_iter=3, tol=1e-9)
    assert (X is None) == bool(g[f"{name}/short/X_is_none"])
    assert msg == str(g[f"{name}/short/msg"])
    assert abs(log_norm - float(g[f"{name}/short/log_norm"])) < 1e-9
    T, conv, n_iter = osol.cycle_reduction_core(A, B, C, max_iter=3, tol=1e-9)
    assert not conv and n_iter == 3 and not T.any()


# ------------------------------------------------------------------------------------------- T, R vs Dynare
@pytest.mark.parametrize("name", ["one_block_1_ss", "rbc_2_block_ss", "full_nk"])
@pytest.mark.parametrize("solver", ["gensys", "cycle_reduction"])
def test_policy_matches_dynare(name, solver):
    """tests/model/test_model.py:532-562 (test_solve_matches_dynare): levels, atol = rtol = 1e-5."""
    g = np.load(GOLD / "ref_dynare_policy.npz")
    mod = model(name)
    A, B, C, D = mod.jacobians(mod.theta_vector(), log_linearize=False)
    if solver == "gensys":
        T, R, success, eu = osol.gensys_policy(A, B, C, D)
        assert success and eu[:2] == [1, 1]
    else:
        T, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-12)
        assert conv
        R = osol.selection_matrix(B, C, D, T)
    T, R = mod.unpermute_policy(T, R)
    rows = [mod.var_names.index(v) for v in g[f"{name}/rows"]]
    cols = [mod.var_names.index(v) for v in g[f"{name}/state_cols"]]
    np.testing.assert_allclose(T[rows][:, cols], g[f"{name}/ghx"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(R[rows], g[f"{name}/ghu"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["rbc", "one_block_1_ss", "rbc_extended", "full_nk", "nk_complete_more_shocks"])
def test_gensys_and_cycle_reduction_agree(name):
    """tests/model/test_perturbation.py:166-206: atol = rtol = 1e-8; jumper columns of T are zero."""
    mod = model(name)
    A, B, C, D = mod.jacobians(mod.theta_vector())
    Tg, Rg, success, _ = osol.gensys_policy(A, B, C, D)
    Tc, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=100000, tol=1e-16)
    if not conv:  # tol 1e-16 may be unreachable in floating point; the reference's numpy twin then still returns X
        Tc, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-12)
    assert success and conv
    np.testing.assert_allclose(Tc, Tg, atol=1e-8, rtol=1e-8)
    np.testing.assert_allclose(osol.selection_matrix(B, C, D, Tc), Rg, atol=1e-8, rtol=1e-8)
    assert not Tc[:, np.abs(A).sum(0) == 0].any()


def test_gensys_failure_codes_on_pert_fails():
    """tests/model/test_model.py:501-529: pert_fails.gcn -> eu = [1, 0, 2]."""
    mod = model("pert_fails")
    A, B, C, D = mod.jacobians(mod.theta_vector(), mode="model")
    _T, _R, success, eu = osol.gensys_policy(A, B, C, D)
    assert not success and eu == [1, 0, 2]


# ------------------------------------------------------------------------------------------- Blanchard-Kahn
@pytest.mark.parametrize("name", ["rbc", "rbc_linearized", "full_nk", "nk_complete_more_shocks"])
def test_bk_counts(name):
    """tests/model/test_model.py:593-651: n_forward == n_unstable; the QZ and the solve+eig variants agree on the count."""
    mod = model(name)
    A, B, C, D = mod.jacobians(mod.theta_vector())
    ok_qz, nf_qz, nu_qz = osol.bk_condition_qz(A, B, C, D)
    ok_pt, nf_pt, nu_pt = osol.bk_condition_pt(A, B, C, D, mod.permuted_lead_var_idx)
    # the QZ variant counts the numerically non-zero lead columns, the estimation variant the structural ones
    # (two structurally-lead columns of nk_complete_more_shocks are numerically zero): each is consistent with itself
    assert ok_qz and ok_pt and nu_qz == nf_qz and nu_pt == nf_pt and nf_pt >= nf_qz


def test_bk_violated_on_pert_fails():
    mod = model("pert_fails")
    A, B, C, D = mod.jacobians(mod.theta_vector(), mode="model")
    ok, n_fwd, n_unst = osol.bk_condition_pt(A, B, C, D, mod.permuted_lead_var_idx)
    assert not ok and n_unst != n_fwd


# ------------------------------------------------------------------------------------------- Kalman known answers
def test_kalman_ar1_closed_form():
    """Scalar AR(1) observed without noise: the exact Gaussian likelihood is known in closed form."""
    rho, sig, n_obs = 0.8, 0.3, 60
    rng = np.random.default_rng(0)
    y = np.zeros(n_obs)
    y[0] = rng.standard_normal() * sig / np.sqrt(1 - rho**2)
    for t in range(1, n_obs):
        y[t] = rho * y[t - 1] + sig * rng.standard_normal()
    T, R, Q, Z, H = np.array([[rho]]), np.array([[1.0]]), np.array([[sig**2]]), np.array([[1.0]]), np.zeros((1, 1))
    ll = oss.kalman_loglik(y[:, None], T, R, Q, Z, H, jitter=0.0)
    exact = scipy.stats.norm.logpdf(y[0], scale=sig / np.sqrt(1 - rho**2)) + scipy.stats.norm.logpdf(y[1:], loc=rho * y[:-1], scale=sig).sum()
    assert abs(ll - exact) < 1e-10


@pytest.mark.parametrize("p,with_h", [(1, False), (2, True), (3, True)])
def test_kalman_matches_joint_gaussian_density(p, with_h):
    """log p(y_1..y_T) of a small state-space model = one big multivariate-normal density (computed independently)."""
    rng = np.random.default_rng(p)
    n, k, n_obs = 4, 2, 12
    T = 0.5 * rng.standard_normal((n, n))
    T *= 0.8 / max(abs(np.linalg.eigvals(T)))
    R = rng.standard_normal((n, k))
    Q = np.diag(rng.random(k) + 0.5)
    Z = rng.standard_normal((p, n))
    H = np.diag(rng.random(p) * 0.1 + 0.05) if with_h else np.zeros((p, p))
    Y = rng.standard_normal((n_obs, p))
    ll = oss.kalman_loglik(Y, T, R, Q, Z, H, jitter=0.0)
    P0 = oss.dlyap(T, R @ Q @ R.T)
    # joint covariance of (y_1..y_T): Cov(x_s, x_t) = T^(s-t) P0 for s >= t
    big = np.zeros((n_obs * p, n_obs * p))
    for s in range(n_obs):
        for t in range(n_obs):
            Cx = np.linalg.matrix_power(T, s - t) @ P0 if s >= t else P0 @ np.linalg.matrix_power(T, t - s).T
            big[s * p : (s + 1) * p, t * p : (t + 1) * p] = Z @ Cx @ Z.T + (H if s == t else 0)
    exact = scipy.stats.multivariate_normal.logpdf(Y.ravel(), mean=np.zeros(n_obs * p), cov=big, allow_singular=False)
    assert abs(ll - exact) < 1e-8


def test_kalman_missing_rows_equal_dropping_them_from_the_joint_density():
    rng = np.random.default_rng(5)
    n, k, p, n_obs = 3, 2, 2, 10
    T = 0.4 * rng.standard_normal((n, n))
    R = rng.standard_normal((n, k))
    Q = np.eye(k)
    Z = rng.standard_normal((p, n))
    H = 0.1 * np.eye(p)
    Y = rng.standard_normal((n_obs, p))
    Ym = Y.copy()
    Ym[3, 0] = np.nan
    Ym[6] = np.nan
    Ym[8, 1] = oss.MISSING_FILL
    jit = 1e-10  # the masking scheme NEEDS a positive jitter: a masked row leaves F_aa = jitter (zero would be singular)
    ll = oss.kalman_loglik(Ym, T, R, Q, Z, H, jitter=jit, mvn_const="per_obs")
    P0 = oss.dlyap(T, R @ Q @ R.T)
    big = np.zeros((n_obs * p, n_obs * p))
    for s in range(n_obs):
        for t in range(n_obs):
            Cx = np.linalg.matrix_power(T, s - t) @ P0 if s >= t else P0 @ np.linalg.matrix_power(T, t - s).T
            big[s * p : (s + 1) * p, t * p : (t + 1) * p] = Z @ Cx @ Z.T + (H if s == t else 0)
    keep = ~(np.isnan(Ym) | (Ym == oss.MISSING_FILL)).ravel()
    exact = scipy.stats.multivariate_normal.logpdf(Y.ravel()[keep], mean=np.zeros(keep.sum()), cov=big[np.ix_(keep, keep)])
    # a masked entry of a partially observed row stays in the p x p innovation system as an independent N(0, jitter)
    # coordinate observed at 0: it adds -(log 2 pi + log jitter)/2 (a constant in theta) to the exact density
    n_partial_missing = 2
    assert abs(ll - (exact - 0.5 * (np.log(2 * np.pi) + np.log(jit)) * n_partial_missing)) < 1e-6


def _joint_cov(T, R, Q, Z, H, n_obs, jitter):
    """Covariance of (y_1..y_T) for the model the jittered filter is EXACT for (jitter I added to H inside F and to
    every filtered covariance): observation noise H + j I, and an extra N(0, j I) disturbance of the filtered state before
    each prediction, i.e. state noise R Q R' + j T T'; the first predicted covariance is dlyap(T, R Q R') of the original model."""
    n, p = T.shape[0], Z.shape[0]
    Sig = [oss.dlyap(T, R @ Q @ R.T)]
    for _ in range(n_obs - 1):
        Sig.append(T @ Sig[-1] @ T.T + R @ Q @ R.T + jitter * T @ T.T)
    big = np.zeros((n_obs * p, n_obs * p))
    for s in range(n_obs):
        for t in range(n_obs):
            Cx = np.linalg.matrix_power(T, s - t) @ Sig[t] if s >= t else Sig[s] @ np.linalg.matrix_power(T, t - s).T
            big[s * p : (s + 1) * p, t * p : (t + 1) * p] = Z @ Cx @ Z.T + ((H + jitter * np.eye(p)) if s == t else 0)
    return big


def _gauss_logpdf(y, cov):
    """log N(y; 0, cov) by Cholesky (scipy's multivariate_normal truncates eigenvalues: not accurate enough here)."""
    c = scipy.linalg.cholesky(cov, lower=True)
    z = scipy.linalg.solve_triangular(c, y, lower=True)
    return -0.5 * (len(y) * np.log(2.0 * np.pi) + 2.0 * np.log(np.diag(c)).sum() + z @ z)


@pytest.mark.parametrize("jitter", [1e-8, 1e-3])
def test_kalman_jitter_placement_known_answer(jitter):
    """Known answer at the production jitter (1e-8; 1e-3 makes a misplaced jitter visible).  SURVEY A.5 puts the jitter on H
    inside F and on every filtered covariance, while the Joseph term K H K' uses H WITHOUT it (as upstream does).  Expanding
    the Joseph form with P Z' = K F gives the equivalent statement
        P+ = P - K F K' - j K K' + j I,      F = Z P Z' + H + j I,
    i.e. the textbook update of the model with observation noise H + j I, minus j K K', plus j I.  That form is written out
    here independently (no Joseph form, no symmetrisation, explicit inverse) and must reproduce the oracle; at j = 1e-8 the
    exact joint Gaussian density of the inflated model (noise H + j I, state noise R Q R' + j T T') agrees with both to 1e-7
    on this regular problem, which is NOT true of the benchmark configurations -- there the placement moves ll by 1e-2 ... 1
    (see DESIGN.md section 4), so the recursion above is what parity means."""
    rng = np.random.default_rng(11)
    n, k, p, n_obs = 4, 2, 2, 14
    T = 0.5 * rng.standard_normal((n, n))
    T *= 0.85 / max(abs(np.linalg.eigvals(T)))
    R = rng.standard_normal((n, k))
    Q = np.diag(rng.random(k) + 0.5)
    Z = rng.standard_normal((p, n))
    H = np.diag([0.05, 0.0])  # one observable without measurement error
    Y = rng.standard_normal((n_obs, p))
    ll = oss.kalman_loglik(Y, T, R, Q, Z, H, jitter=jitter)
    a, P, ref = np.zeros(n), oss.dlyap(T, R @ Q @ R.T), 0.0
    for t in range(n_obs):
        F = Z @ P @ Z.T + H + jitter * np.eye(p)
        Fi = np.linalg.inv(F)
        K = P @ Z.T @ Fi
        v = Y[t] - Z @ a
        ref += -0.5 * (p * np.log(2 * np.pi) + np.log(np.linalg.det(F)) + v @ Fi @ v)
        a = T @ (a + K @ v)
        P = T @ (P - K @ F @ K.T - jitter * K @ K.T + jitter * np.eye(n)) @ T.T + R @ Q @ R.T
    assert abs(ll - ref) < 1e-9
    if jitter == 1e-8:
        assert abs(ll - _gauss_logpdf(Y.ravel(), _joint_cov(T, R, Q, Z, H, n_obs, jitter))) < 1e-7


@pytest.mark.parametrize("mask_intercept", [False, True])
def test_kalman_intercept_at_missing_entries(mask_intercept):
    """a non-zero intercept d meets missing data.  mask_intercept=True: missing entries drop out of the density
    altogether; False (SURVEY A.5's restatement of upstream): each missing entry of a partially observed row is scored as
    v_i = -d_i against F_ii = jitter, i.e. adds -(d_i^2 / jitter) / 2 on top."""
    rng = np.random.default_rng(12)
    n, k, p, n_obs, jit = 3, 2, 2, 9, 1e-6
    T = 0.4 * rng.standard_normal((n, n))
    R = rng.standard_normal((n, k))
    Q, Z, H = np.eye(k), rng.standard_normal((p, n)), 0.1 * np.eye(p)
    d = np.array([0.3, -0.2])
    Y = rng.standard_normal((n_obs, p)) + d
    Ym = Y.copy()
    Ym[2, 0] = np.nan
    Ym[5] = np.nan
    Ym[7, 1] = np.nan
    ll = oss.kalman_loglik(Ym, T, R, Q, Z, H, d=d, jitter=jit, mask_intercept=mask_intercept)
    keep = ~np.isnan(Ym).ravel()
    big = _joint_cov(T, R, Q, Z, H, n_obs, jit)
    mean = np.tile(d, n_obs)
    exact = _gauss_logpdf((Y.ravel() - mean)[keep], big[np.ix_(keep, keep)])
    extra = -0.5 * (np.log(2 * np.pi) + np.log(jit)) * 2  # two masked entries in partially observed rows (see the test above)
    if not mask_intercept:
        extra += -0.5 * (d[0] ** 2 + d[1] ** 2) / jit
    assert abs(ll - (exact + extra)) < 1e-5 * max(1.0, abs(extra) * 1e-6)


def test_dlyap_fixed_point():
    mod = model("full_nk")
    A, B, C, D = mod.jacobians(mod.theta_vector())
    T, conv, _ = osol.cycle_reduction_core(A, B, C, tol=1e-12)
    R = osol.selection_matrix(B, C, D, T)
    Q = np.eye(mod.k) * SIGMA_SHOCK**2
    P = oss.dlyap(T, R @ Q @ R.T)
    assert np.abs(P - (T @ P @ T.T + R @ Q @ R.T)).max() < 1e-14 * max(1, np.abs(P).max())


@pytest.mark.parametrize("name,with_err", [("rbc", False), ("full_nk", True)])
def test_compiled_cpu_port_matches_the_numpy_oracle(name, with_err):
    """oracle/fast.py (numba; what bench.py times as the CPU baseline) computes the same gated log-likelihood."""
    from helpers import SIGMA_ERR, draws, simulate_obs
    from oracle import fast

    mod = model(name)
    observed = mod.spec["observed_default"]
    Y = simulate_obs(mod, 60, seed=0, sigma_err=SIGMA_ERR if with_err else 0.0)
    sig = np.full(mod.k, SIGMA_SHOCK)
    err = np.full(len(observed), SIGMA_ERR) if with_err else None
    th = np.vstack([draws(mod, 5, seed=2, width=0.04, valid=True), draws(mod, 3, seed=3, width=0.08, valid=False)])
    for t in th:
        ref = oss.loglik(mod, t, Y, observed, sig, err, tol=1e-8, max_iter=100)
        got = fast.loglik(mod, t, Y, observed, sig, err, tol=1e-8, max_iter=100)
        if ref["ok"] and np.isfinite(ref["ll"]):
            assert abs(got - ref["ll"]) <= 1e-8
        else:
            assert np.isneginf(got)
