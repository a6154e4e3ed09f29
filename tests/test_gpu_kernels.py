"""GPU parity tests: every CUDA kernel, called through the C ABI, against the CPU oracle on identical seeded inputs.

Tolerances are the north star's (BASELINE.json): flags exact, T/R <= 1e-9 relative Frobenius, log-likelihood
<= 1e-7 absolute.
"""

from __future__ import annotations

import numpy as np
import pytest

from helpers import SIGMA_ERR, SIGMA_SHOCK, draws, jacobian_batch, model, observed_idx, rel_fro, simulate_obs
from oracle import solvers as osol
from oracle import statespace as oss

pytestmark = pytest.mark.gpu

TOL_TR = 1e-9
TOL_LL = 1e-7

MODELS = ["rbc", "one_block_1_ss", "rbc_extended", "full_nk", "new_keynesian", "nk_complete_more_shocks", "nk_rbc_composite"]


@pytest.fixture(scope="module")
def B():
    from geconpy_b200 import batched

    return batched


# ------------------------------------------------------------------------------------------- building blocks
@pytest.mark.parametrize("n", [1, 3, 8, 9, 16, 17, 24, 31, 40, 45, 56, 64])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_matches_numpy(B, rng, n, ta, tb):
    A = rng.standard_normal((5, n, n))
    Bm = rng.standard_normal((5, n, n))
    out = B.gemm(A, Bm, trans_a=ta, trans_b=tb, alpha=-0.5)
    ref = -0.5 * np.matmul(A.transpose(0, 2, 1) if ta else A, Bm.transpose(0, 2, 1) if tb else Bm)
    assert np.abs(out - ref).max() <= 1e-13 * n


@pytest.mark.parametrize("n,m", [(1, 1), (5, 2), (9, 9), (24, 24), (24, 4), (31, 9), (33, 33), (45, 13), (56, 56), (64, 64), (66, 5), (72, 72), (88, 88)])
def test_solve_matches_lapack(B, rng, n, m):
    M = rng.standard_normal((7, n, n)) + 0.1 * np.eye(n)
    M[1, [0, -1]] = M[1, [-1, 0]]  # force pivoting
    rhs = rng.standard_normal((7, n, m))
    X, st = B.solve(M, rhs)
    ref = np.linalg.solve(M, rhs)
    assert (st == 0).all()
    for i in range(7):
        assert rel_fro(X[i], ref[i]) <= 1e-11 * max(1.0, np.linalg.cond(M[i]) / 1e3)


def test_solve_singular_is_nan_filled(B, rng):
    n = 6
    M = rng.standard_normal((3, n, n))
    M[1, :, 2] = 0.0
    X, st = B.solve(M, rng.standard_normal((3, n, 2)))
    assert st[0] == 0 and st[2] == 0 and st[1] != 0
    assert np.isnan(X[1]).all() and np.isfinite(X[0]).all()


# ------------------------------------------------------------------------------------------- cycle reduction
@pytest.mark.parametrize("name", MODELS)
def test_cycle_reduction_parity(B, name):
    mod = model(name)
    th = draws(mod, 24, seed=1)
    A, Bm, C, D = jacobian_batch(mod, th)
    res = B.cr_solve(A, Bm, C, D, max_iter=1000, tol=1e-9, resid_tol=1e-8)
    n_conv = 0
    for i in range(len(th)):
        T, conv, n_iter = osol.cycle_reduction_core(A[i], Bm[i], C[i], max_iter=1000, tol=1e-9)
        assert bool(res.converged[i]) == bool(conv), (name, i)
        assert int(res.n_iter[i]) == n_iter, (name, i)
        if conv:
            n_conv += 1
            R = osol.selection_matrix(Bm[i], C[i], D[i], T)
            assert rel_fro(res.T[i], T) <= TOL_TR, (name, i)
            assert rel_fro(res.R[i], R) <= TOL_TR, (name, i)
            # the residual of a converged draw is rounding noise (~1e-16..1e-25); the gate it feeds is 1e-8
            assert abs(res.resid[i] - osol.policy_residual(A[i], Bm[i], C[i], T)) <= 1e-13 + 1e-6 * res.resid[i]
            # jumper columns of T are exactly zero (tests/model/test_perturbation.py:166-206)
            zero_cols = np.abs(A[i]).sum(axis=0) == 0
            assert (res.T[i][:, zero_cols] == 0).all()
    assert n_conv >= len(th) // 2


def test_cycle_reduction_unpermute_and_flags(B):
    mod = model("full_nk")
    th = draws(mod, 4, seed=2)
    A, Bm, C, D = jacobian_batch(mod, th)
    res = B.cr_solve(A, Bm, C, D, unperm=mod.inv_var_order.astype(np.int32))
    raw = B.cr_solve(A, Bm, C, D)
    for i in range(4):
        T, R = mod.unpermute_policy(raw.T[i], raw.R[i])
        assert np.array_equal(res.T[i], T, equal_nan=True) and np.array_equal(res.R[i], R, equal_nan=True)
    # max_iter exhausted -> not converged, T = 0 (cycle_reduction.py:181-183)
    short = B.cr_solve(A, Bm, C, D, max_iter=3)
    assert not short.converged.any()
    assert (short.T == 0).all()
    for i in range(4):
        T, conv, n_iter = osol.cycle_reduction_core(A[i], Bm[i], C[i], max_iter=3, tol=1e-9)
        assert not conv and n_iter == int(short.n_iter[i])  # 3, or 1 for a draw whose Jacobian is NaN
        Rref = osol.selection_matrix(Bm[i], C[i], D[i], T)
        if np.isfinite(Rref).all():
            assert rel_fro(short.R[i], Rref) <= TOL_TR
        else:
            assert np.isnan(short.R[i]).all()


@pytest.mark.parametrize("n,k", [(57, 6), (60, 13), (64, 64)])
def test_cycle_reduction_largest_size(B, rng, n, k):
    """Padded dimension 64 (A1hat in the global workspace): synthetic well-conditioned systems against the oracle."""
    N = 5
    A = 0.3 * rng.standard_normal((N, n, n)) / np.sqrt(n)
    C = 0.3 * rng.standard_normal((N, n, n)) / np.sqrt(n)
    A[:, :, n // 2 :] = 0.0   # lag columns first, lead columns last, as in solver order
    C[:, :, : n // 3] = 0.0
    Bm = np.eye(n) + 0.2 * rng.standard_normal((N, n, n)) / np.sqrt(n)
    D = rng.standard_normal((N, n, k))
    res = B.cr_solve(A, Bm, C, D, max_iter=200, tol=1e-10, resid_tol=1e-8)
    for i in range(N):
        T, conv, n_iter = osol.cycle_reduction_core(A[i], Bm[i], C[i], max_iter=200, tol=1e-10)
        assert conv and bool(res.converged[i]) and int(res.n_iter[i]) == n_iter
        R = osol.selection_matrix(Bm[i], C[i], D[i], T)
        assert rel_fro(res.T[i], T) <= TOL_TR and rel_fro(res.R[i], R) <= TOL_TR
        assert res.status[i] == 0


def test_backward_looking(B, rng):
    n, k = 7, 2
    A = rng.standard_normal((3, n, n)) * 0.3
    Bm = rng.standard_normal((3, n, n)) + 2 * np.eye(n)
    D = rng.standard_normal((3, n, k))
    res = B.cr_solve(A, Bm, None, D)
    for i in range(3):
        T, R = osol.backward_direct(A[i], Bm[i], None, D[i])
        assert rel_fro(res.T[i], T) <= TOL_TR and rel_fro(res.R[i], R) <= TOL_TR
        assert np.abs(A[i] + Bm[i] @ res.T[i]).max() <= 1e-12  # tests/solvers/test_backward_looking.py:38-63


# ------------------------------------------------------------------------------------------- Blanchard-Kahn
@pytest.mark.parametrize("name", MODELS)
def test_bk_count_parity(B, name):
    mod = model(name)
    th = draws(mod, 24, seed=3, width=0.1)
    A, Bm, C, D = jacobian_batch(mod, th)
    fin = np.isfinite(A).all(axis=(1, 2)) & np.isfinite(Bm).all(axis=(1, 2)) & np.isfinite(C).all(axis=(1, 2))
    assert fin.sum() >= 8
    A, Bm, C, D, th = A[fin], Bm[fin], C[fin], D[fin], th[fin]
    lead = mod.permuted_lead_var_idx.astype(np.int32)
    nu, st = B.bk_count(A, Bm, C, lead)
    from geconpy_b200 import _lib as L

    for i in range(len(th)):
        ok, n_fwd, n_unst = osol.bk_condition_pt(A[i], Bm[i], C[i], D[i], mod.permuted_lead_var_idx)
        assert not (st[i] & L.ST_BK_INCONCLUSIVE), (name, i)
        assert int(nu[i]) == n_unst, (name, i)
        assert bool(st[i] & L.ST_BK) == (not ok), (name, i)


@pytest.mark.parametrize("name", MODELS)
def test_bk_certificate_from_the_solver_kernel(B, name):
    """cr_solve(lead_idx=...) proves n_unstable == n_forward for converged draws by spectral-radius certificates;
    whatever it certifies must agree with the oracle's eigenvalue count, and bk_count must skip exactly those draws."""
    from geconpy_b200 import _lib as L

    mod = model(name)
    th = np.vstack([draws(mod, 16, seed=13, width=0.06, valid=True), draws(mod, 16, seed=14, width=0.08, valid=False)])
    A, Bm, C, D = jacobian_batch(mod, th)
    fin = np.isfinite(A).all(axis=(1, 2)) & np.isfinite(Bm).all(axis=(1, 2)) & np.isfinite(C).all(axis=(1, 2))
    A, Bm, C, D = A[fin], Bm[fin], C[fin], D[fin]
    lead = mod.permuted_lead_var_idx.astype(np.int32)
    res = B.cr_solve(A, Bm, C, D, max_iter=200, tol=1e-9, lead_idx=lead)
    cert = (res.status & L.ST_BK_CERTIFIED) != 0
    n_ok = 0
    for i in range(len(A)):
        ok, n_fwd, n_unst = osol.bk_condition_pt(A[i], Bm[i], C[i], D[i], mod.permuted_lead_var_idx)
        n_ok += ok
        if cert[i]:
            assert ok and n_unst == n_fwd == int(res.n_unstable[i]), (name, i)
        else:
            assert int(res.n_unstable[i]) == -1
    assert cert.sum() >= 0.9 * n_ok, (name, int(cert.sum()), n_ok)  # nearly every determinate draw is certified
    # the exact kernel fills in the rest and leaves certified draws alone
    nu, st = B.bk_count(A, Bm, C, lead, status=res.status.copy(), skip_mask=L.ST_BK_CERTIFIED, n_unstable=res.n_unstable.copy())
    for i in range(len(A)):
        ok, n_fwd, n_unst = osol.bk_condition_pt(A[i], Bm[i], C[i], D[i], mod.permuted_lead_var_idx)
        assert int(nu[i]) == n_unst and bool(st[i] & L.ST_BK) == (not ok), (name, i)


def test_bk_pert_fails_model(B):
    """pert_fails.gcn is the reference's broken-model fixture (tests/model/test_model.py:501-529)."""
    mod = model("pert_fails")
    if not mod.analytic_ss:
        pytest.skip("pert_fails has no analytic steady state in the spec")
    th = draws(mod, 2, seed=0)
    A, Bm, C, D = jacobian_batch(mod, th)
    nu, st = B.bk_count(A, Bm, C, mod.permuted_lead_var_idx.astype(np.int32))
    for i in range(2):
        ok, _, n_unst = osol.bk_condition_pt(A[i], Bm[i], C[i], D[i], mod.permuted_lead_var_idx)
        assert int(nu[i]) == n_unst


# ------------------------------------------------------------------------------------------- Lyapunov / Kalman
def _policies(mod, th):
    A, Bm, C, D = jacobian_batch(mod, th)
    T, R = [], []
    for i in range(len(th)):
        t, conv, _ = osol.cycle_reduction_core(A[i], Bm[i], C[i], max_iter=1000, tol=1e-9)
        assert conv
        r = osol.selection_matrix(Bm[i], C[i], D[i], t)
        t, r = mod.unpermute_policy(t, r)
        T.append(t), R.append(r)
    return np.stack(T), np.stack(R)


@pytest.mark.parametrize("name", ["rbc", "full_nk", "nk_complete_more_shocks"])
def test_dlyap_parity(B, name):
    mod = model(name)
    th = draws(mod, 6, seed=4, width=0.02, valid=True)
    T, R = _policies(mod, th)
    q = np.full(mod.k, SIGMA_SHOCK**2)
    P, st, it = B.dlyap(T, R, q)
    assert (st == 0).all()
    for i in range(len(th)):
        ref = oss.dlyap(T[i], R[i] @ np.diag(q) @ R[i].T)
        # scipy's bilinear-transform solve and the doubling iteration both carry ~cond * eps error when rho(T) -> 1
        assert np.abs(P[i] - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
        assert np.abs(P[i] - (T[i] @ P[i] @ T[i].T + R[i] @ np.diag(q) @ R[i].T)).max() <= 1e-13 * max(1.0, np.abs(ref).max())


# (the last two are BASELINE configs 4a / 4b at their own sample length, T_obs = 200)
@pytest.mark.parametrize("name,Tobs", [("rbc", 100), ("rbc", 200), ("full_nk", 200), ("nk_complete_more_shocks", 50), ("nk_rbc_composite", 40),
                                       ("nk_complete_more_shocks", 200), ("nk_rbc_composite", 200)])
@pytest.mark.parametrize("selector", [True, False])
def test_kalman_parity(B, name, Tobs, selector):
    mod = model(name)
    th = draws(mod, 6, seed=5, width=0.02, valid=True)
    T, R = _policies(mod, th)
    Y = simulate_obs(mod, Tobs, seed=0, sigma_err=SIGMA_ERR)
    obs = observed_idx(mod)
    p = len(obs)
    q = np.full((len(th), mod.k), SIGMA_SHOCK**2) * (1.0 + 0.1 * np.arange(len(th)))[:, None]
    h = np.full(p, SIGMA_ERR**2)
    Z = oss.selector_design(mod.var_names, mod.spec["observed_default"])
    kw = dict(obs_idx=obs) if selector else dict(Z=Z)
    ll, st, llt = B.kalman_loglik(T, R, q, Y, hdiag=h, return_per_step=True, **kw)
    for i in range(len(th)):
        ref, ref_t = oss.kalman_loglik(Y, T[i], R[i], np.diag(q[i]), Z, np.diag(h), return_all=True)
        assert st[i] == 0
        assert np.abs(llt[i] - ref_t).max() <= TOL_LL, (name, i, np.abs(llt[i] - ref_t).max())
        assert abs(ll[i] - ref) <= TOL_LL, (name, i, ll[i], ref)


def test_kalman_missing_data_no_measurement_error_and_intercept(B):
    mod = model("full_nk")
    th = draws(mod, 3, seed=6, width=0.02, valid=True)
    T, R = _policies(mod, th)
    Y = simulate_obs(mod, 120, seed=1)
    rng = np.random.default_rng(7)
    Ym = Y.copy()
    Ym[rng.random(Y.shape) < 0.25] = np.nan
    Ym[5] = np.nan          # a fully missing row
    Ym[9, 0] = -9999.0      # the fill value also marks a missing entry
    obs = observed_idx(mod)
    q = np.full(mod.k, SIGMA_SHOCK**2)
    Z = oss.selector_design(mod.var_names, mod.spec["observed_default"])
    d = np.array([0.01, -0.02, 0.005])
    for Yc, dd in ((Y, None), (Ym, None), (Y, d)):
        ll, st = B.kalman_loglik(T, R, q, Yc, obs_idx=obs, d=dd)
        ll2, _ = B.kalman_loglik(T, R, q, Yc, Z=Z, d=dd, mvn_const="bare")
        for i in range(len(th)):
            ref = oss.kalman_loglik(Yc, T[i], R[i], np.diag(q), Z, np.zeros((3, 3)), d=dd)
            ref2 = oss.kalman_loglik(Yc, T[i], R[i], np.diag(q), Z, np.zeros((3, 3)), d=dd, mvn_const="bare")
            assert abs(ll[i] - ref) <= TOL_LL, (i, ll[i], ref)
            assert abs(ll2[i] - ref2) <= TOL_LL, (i, ll2[i], ref2)


@pytest.mark.parametrize("n,p", [(3, 2), (10, 3), (40, 4)])  # one thread per draw / one warp per draw / one CTA per draw
@pytest.mark.parametrize("mask_intercept", [False, True])
def test_kalman_intercept_meets_missing_data(B, rng, n, p, mask_intercept):
    """a non-zero observation intercept together with missing entries, in both conventions
    (gecon_kalman_args.mask_intercept): forward kernels (selector and dense Z) and the gradient kernel against the oracle."""
    from oracle import adjoints as oad

    N, Tobs, k = 3, 40, 3
    T, R = _random_statespace(rng, N, n, k)
    q = 0.5 + rng.random((N, k))
    h = 0.1 + rng.random((N, p))
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    d = 0.5 + 0.1 * rng.standard_normal((N, p))
    Y = rng.standard_normal((Tobs, p)) + d[0]
    Y[rng.random(Y.shape) < 0.25] = np.nan
    Y[4] = np.nan
    jit = 1e-6  # (at 1e-8 the unmasked convention scores d^2 / jitter ~ 1e7 per missing entry: the comparison loses digits)
    for kw in (dict(obs_idx=obs), dict(Z=Z)):
        ll, st = B.kalman_loglik(T, R, q, Y, hdiag=h, d=d, jitter=jit, mask_intercept=mask_intercept, **kw)
        for i in range(N):
            ref = oss.kalman_loglik(Y, T[i], R[i], np.diag(q[i]), Z, np.diag(h[i]), d=d[i], jitter=jit, mask_intercept=mask_intercept)
            assert st[i] == 0 and abs(ll[i] - ref) <= max(TOL_LL, abs(ref) * 1e-12), (n, kw.keys(), i, ll[i], ref)
    if n <= 48:
        out = B.kalman_loglik_grad(T, R, q, Y, obs_idx=obs, hdiag=h, d=d, jitter=jit, mask_intercept=mask_intercept)
        for i in range(N):
            ref = oad.kalman_loglik_adjoints(Y, T[i], R[i], q[i], Z, h[i], d=d[i], jitter=jit, mask_intercept=mask_intercept)
            assert abs(out["ll"][i] - ref["ll"]) <= max(TOL_LL, abs(ref["ll"]) * 1e-12)
            for key in ("T", "R", "q", "h", "d"):
                assert np.abs(out[key][i] - ref[key]).max() <= 1e-7 * max(1.0, np.abs(ref[key]).max()), (key, i)


def _random_statespace(rng, N, n, k, rho=0.9):
    T = rng.standard_normal((N, n, n))
    if n > 1:
        T[:, :, n - max(1, n // 4) :] = 0.0  # jumper columns, as a policy matrix has
    for i in range(N):
        T[i] *= rho / np.abs(np.linalg.eigvals(T[i])).max()
    R = rng.standard_normal((N, n, k))
    return T, R


@pytest.mark.parametrize("n,k,p", [(1, 1, 1), (3, 2, 1), (2, 1, 1), (2, 3, 2), (3, 1, 2), (4, 2, 1), (4, 5, 2),  # one thread per draw (n <= 4, p <= 2)
                                    (5, 3, 2), (7, 2, 4), (12, 4, 3), (15, 6, 5), (15, 16, 8), (20, 7, 7), (23, 9, 6), (23, 3, 8),
                                    (9, 2, 1), (9, 4, 5), (10, 4, 3), (10, 3, 8), (11, 5, 2), (11, 2, 7),  # fringe variant of the warp kernel
                                    (24, 5, 2), (40, 8, 8), (60, 10, 7), (63, 4, 3)])
def test_kalman_synthetic_sizes(B, rng, n, k, p):
    """The three filter kernels over their whole size range (one thread per draw up to n = 4 and p = 2, one warp per draw up to n = 31,
    one CTA per draw above), every
    number of observables, with measurement error, missing data, an intercept and per-step output."""
    N, Tobs = 4, 45
    T, R = _random_statespace(rng, N, n, k)
    q = 0.5 + rng.random((N, k))
    h = 0.1 + rng.random((N, p))
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    d = 0.1 * rng.standard_normal(p)
    x = np.zeros(n)
    Y = np.zeros((Tobs, p))
    for t in range(Tobs):
        x = T[0] @ x + R[0] @ (np.sqrt(q[0]) * rng.standard_normal(k))
        Y[t] = x[obs] + d + np.sqrt(h[0]) * rng.standard_normal(p)
    Ym = Y.copy()
    Ym[rng.random(Y.shape) < 0.2] = np.nan
    Ym[3] = np.nan
    for Yc, dd in ((Y, None), (Ym, d)):
        ll, st, llt = B.kalman_loglik(T, R, q, Yc, obs_idx=obs, hdiag=h, d=dd, return_per_step=True)
        ll1, st1 = B.kalman_loglik(T, R, q, Yc, obs_idx=obs, hdiag=h, d=dd)
        for i in range(N):
            ref, ref_t = oss.kalman_loglik(Yc, T[i], R[i], np.diag(q[i]), Z, np.diag(h[i]), d=dd, return_all=True)
            assert st[i] == 0 and st1[i] == 0
            scale = max(1.0, abs(ref) * 1e-9)  # 1e-7 absolute on likelihoods of order 100
            assert np.abs(llt[i] - ref_t).max() <= TOL_LL * scale, (n, p, i, np.abs(llt[i] - ref_t).max())
            assert abs(ll[i] - ref) <= TOL_LL * scale and abs(ll1[i] - ref) <= TOL_LL * scale, (n, p, i, ll[i], ll1[i], ref)


@pytest.mark.parametrize("n,k,p,tc", [(10, 4, 3, 9), (10, 3, 2, 5), (19, 9, 7, 16), (19, 4, 3, 7), (26, 13, 7, 23), (7, 2, 3, 4), (40, 5, 3, 20)])
def test_kalman_zero_columns_promise(B, rng, n, k, p, tc):
    """gecon_kalman_args.t_cols: with only the first t_cols columns of T non-zero (the structure of a policy matrix whose filter variables
    are ordered [lagged | observed only]), the warp-per-draw kernel skips the k-steps beyond them -- the result is IDENTICAL to the
    dense run (complete and incomplete samples) and both agree with the oracle.  (n = 40: the CTA kernel ignores the promise.)"""
    N, Tobs = 6, 40
    T, R = _random_statespace(rng, N, n, k)
    T[:, :, tc:] = 0.0
    for i in range(N):
        T[i] *= 0.9 / np.abs(np.linalg.eigvals(T[i])).max()
    q = 0.5 + rng.random((N, k))
    h = 0.1 + rng.random((N, p))
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    obs[-1] = n - 1  # an observed variable that is not a state
    obs = np.unique(obs).astype(np.int32)
    p = obs.size
    h = h[:, :p]
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    Y = rng.standard_normal((Tobs, p))
    Y[rng.random(Y.shape) < 0.1] = np.nan
    for Yc in (np.nan_to_num(Y, nan=0.3), Y):
        ll0, st0 = B.kalman_loglik(T, R, q, Yc, obs_idx=obs, hdiag=h)
        ll1, st1 = B.kalman_loglik(T, R, q, Yc, obs_idx=obs, hdiag=h, t_cols=tc)
        assert np.array_equal(ll0, ll1) and np.array_equal(st0, st1) and (st0 == 0).all()
        for i in range(N):
            ref = oss.kalman_loglik(Yc, T[i], R[i], np.diag(q[i]), Z, np.diag(h[i]))
            assert abs(ll1[i] - ref) <= TOL_LL * max(1.0, abs(ref) * 1e-9), (i, ll1[i], ref)
    with pytest.raises(Exception, match="t_cols"):
        B.kalman_loglik(T, R, q, Yc, obs_idx=obs, hdiag=h, t_cols=n + 1)


def test_kalman_gating_and_given_P0(B):
    mod = model("rbc")
    th = draws(mod, 4, seed=8, width=0.02, valid=True)
    T, R = _policies(mod, th)
    Y = simulate_obs(mod, 60, seed=2)
    q = np.full(mod.k, SIGMA_SHOCK**2)
    obs = observed_idx(mod)
    P0 = np.stack([oss.dlyap(T[i], R[i] @ np.diag(q) @ R[i].T) for i in range(4)])
    status_in = np.array([0, 0x10, 0, 0x8], dtype=np.int32)
    ll, st = B.kalman_loglik(T, R, q, Y, obs_idx=obs, P0=P0, status_in=status_in, gate_mask=0x18)
    assert np.isneginf(ll[1]) and np.isneginf(ll[3])
    Z = oss.selector_design(mod.var_names, mod.spec["observed_default"])
    for i in (0, 2):
        ref = oss.kalman_loglik(Y, T[i], R[i], np.diag(q), Z, np.zeros((1, 1)), P0=P0[i])
        assert abs(ll[i] - ref) <= TOL_LL


def test_device_pointer_path_matches_host_path(B):
    import torch

    mod = model("full_nk")
    th = draws(mod, 8, seed=9)
    A, Bm, C, D = jacobian_batch(mod, th)
    host = B.cr_solve(A, Bm, C, D)
    dev = B.cr_solve(*(torch.as_tensor(x, device="cuda") for x in (A, Bm, C, D)))
    assert np.array_equal(dev.T.cpu().numpy(), host.T, equal_nan=True)
    assert np.array_equal(dev.R.cpu().numpy(), host.R, equal_nan=True)
    assert np.array_equal(dev.status.cpu().numpy(), host.status)
