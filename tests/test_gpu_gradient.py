"""GPU parity of the gradient path through the C ABI: Kalman reverse sweep, policy / selection
adjoints (o1_policy_function_adjoints), generated vector-Jacobian product and the whole theta -> (ll, dll/dtheta) chain,
against oracle/adjoints.py (itself pinned against finite differences in tests/test_adjoints_cpu.py)."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import SIGMA_ERR, SIGMA_SHOCK, draws, jacobian_batch, model, simulate_obs
from oracle import adjoints as oad
from oracle import solvers as osol
from oracle import statespace as oss

pytestmark = pytest.mark.gpu
RTOL = 1e-8  # relative to the largest entry of each gradient block


@pytest.fixture(scope="module")
def B():
    from geconpy_b200 import batched

    return batched


def _close(got, ref, rtol=RTOL):
    return np.abs(np.asarray(got) - ref).max() <= rtol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n,k,p,Tobs,missing,selector", [(1, 1, 1, 10, False, True), (5, 2, 2, 25, False, False), (5, 2, 2, 25, True, False),
                                                          (10, 4, 3, 60, True, True), (13, 3, 8, 20, True, False), (19, 9, 7, 30, False, True),
                                                          (33, 5, 4, 15, True, True), (48, 6, 3, 8, False, False)])
@pytest.mark.parametrize("device_path", [False, True])
def test_kalman_grad_matches_oracle(B, n, k, p, Tobs, missing, selector, device_path):
    import torch

    rng = np.random.default_rng(n * 7 + p)
    N = 3
    T = rng.standard_normal((N, n, n))
    for i in range(N):
        T[i] *= 0.8 / np.abs(np.linalg.eigvals(T[i])).max()
    R = rng.standard_normal((N, n, k))
    q, h = 0.5 + rng.random((N, k)), 0.1 + rng.random((N, p))
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    Z = rng.standard_normal((p, n))
    if selector:
        Z = np.zeros((p, n))
        Z[np.arange(p), obs] = 1.0
    Y = rng.standard_normal((Tobs, p))
    d = 0.1 * rng.standard_normal((N, p))
    if missing:
        Y[3, 0], Y[5, p - 1] = np.nan, -9999.0
        Y[7] = np.nan
        d[:] = 0.0
    kw = dict(obs_idx=obs) if selector else dict(Z=Z)
    args = [T, R, np.sqrt(q), Y]
    if device_path:
        args = [torch.as_tensor(x, device="cuda") for x in args]
    out = B.kalman_loglik_grad(*args, hdiag=np.sqrt(h), d=d, sigma_inputs=True, **kw)
    out = {key: (v.cpu().numpy() if hasattr(v, "cpu") else v) for key, v in out.items()}
    ll_fwd, _ = B.kalman_loglik(T, R, q, Y, hdiag=h, d=d, **kw)
    for i in range(N):
        g = oad.kalman_loglik_adjoints(Y, T[i], R[i], q[i], Z, h[i], d[i])
        assert out["status"][i] == 0
        assert abs(out["ll"][i] - g["ll"]) <= 1e-7 and abs(out["ll"][i] - ll_fwd[i]) <= 1e-7
        assert _close(out["T"][i], g["T"]) and _close(out["R"][i], g["R"]) and _close(out["d"][i], g["d"])
        assert _close(out["q"][i], 2 * np.sqrt(q[i]) * g["q"]) and _close(out["h"][i], 2 * np.sqrt(h[i]) * g["h"])


@pytest.mark.parametrize("name", ["rbc", "rbc_extended", "full_nk", "nk_complete_more_shocks", "nk_rbc_composite"])
@pytest.mark.parametrize("with_R", [False, True])
def test_policy_adjoints_match_the_kronecker_restatement(B, name, with_R):
    mod = model(name)
    th = draws(mod, 4, seed=2, width=0.02, valid=True)
    A, Bm, Cm, D = jacobian_batch(mod, th)
    N, n, k = len(th), mod.n, mod.k
    res = B.cr_solve(A, Bm, Cm, D, max_iter=1000, tol=1e-13)
    T, R = res.T, res.R
    rng = np.random.default_rng(1)
    Tb, Rb = rng.standard_normal((N, n, n)), rng.standard_normal((N, n, k))
    Ab, Bb, Cb, Db, st = B.policy_adjoints(A, Bm, Cm, T, Tb, **(dict(D=D, R=R, R_bar=Rb) if with_R else {}))
    for i in range(N):
        tb, b0, c0 = Tb[i], 0.0, 0.0
        if with_R:
            b0, c0, d0, tadd = oad.selection_adjoints(Bm[i], Cm[i], D[i], T[i], R[i], Rb[i])
            tb = tb + tadd
            assert _close(Db[i], d0)
        S, SB, SC = oad.policy_adjoints_kron(A[i], Bm[i], Cm[i], T[i], tb)
        assert st[i] == 0
        assert _close(Ab[i], S) and _close(Bb[i], SB + b0) and _close(Cb[i], SC + c0)


def test_o1_policy_function_adjoints_keeps_the_reference_signature(B):
    from geconpy_b200.solvers.shared import o1_policy_function_adjoints

    mod = model("rbc")
    A, Bm, Cm, D = mod.jacobians(mod.theta_vector(), mode="statespace")
    T = osol.cycle_reduction_core(A, Bm, Cm, max_iter=1000, tol=1e-13)[0]
    Tb = np.random.default_rng(0).standard_normal(T.shape)
    out = o1_policy_function_adjoints(A, Bm, Cm, T, Tb)
    assert isinstance(out, list) and len(out) == 3
    for got, ref in zip(out, oad.policy_adjoints_kron(A, Bm, Cm, T, Tb)):
        assert got.shape == T.shape and _close(got, ref)
    # singular C T + B -> NaN + status, never an exception
    Ab, _, _, _, st = B.policy_adjoints(np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 4)), np.ones((4, 4)))
    assert st != 0 and np.isnan(Ab).all()


@pytest.mark.parametrize("name,Tobs", [("rbc", 60), ("full_nk", 50), ("nk_complete_more_shocks", 30)])
def test_pipeline_gradient_matches_oracle(name, Tobs):
    """theta -> (ll, dll/d[theta, sigma_shock, error_sigma]) through all seven launches."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model(name)
    observed = mod.spec["observed_default"]
    cm = CompiledModel(name)
    ss = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-13, max_iter=1000, chunk=8)
    N = 10
    th = draws(mod, N, seed=31, width=0.02, valid=True)
    Y = simulate_obs(mod, Tobs, seed=3, sigma_err=SIGMA_ERR)
    sig = np.full((N, mod.k), SIGMA_SHOCK) * (1.0 + 0.1 * np.arange(N))[:, None]
    err = np.full((N, len(observed)), SIGMA_ERR)
    full = np.hstack([th, sig, err])
    ll, grad, st = ss.loglik_and_grad(full, Y)
    ll_fwd, st_fwd = ss.loglik(full, Y)
    assert np.array_equal(st, st_fwd)
    n_ok = 0
    for i in range(N):
        ref = oss.loglik(mod, th[i], Y, observed, sig[i], err[i], tol=1e-13, max_iter=1000)
        if not (ref["ok"] and np.isfinite(ref["ll"])):
            assert st[i] != 0 and np.isneginf(ll[i]) and not grad[i].any()
            continue
        n_ok += 1
        g = oad.loglik_grad(mod, th[i], Y, observed, sig[i], err[i])
        assert st[i] == 0 and abs(ll[i] - g["ll"]) <= 1e-7 and abs(ll[i] - ll_fwd[i]) <= 1e-7
        nt, k = mod.theta_vector().size, mod.k
        # the oracle's last stage is a central difference (1e-9 relative): compare at 1e-6 of the largest component
        assert _close(grad[i, :nt], g["theta"], rtol=1e-6), (i, np.abs(grad[i, :nt] - g["theta"]).max())
        assert _close(grad[i, nt : nt + k], g["sigma_shock"], rtol=1e-7) and _close(grad[i, nt + k :], g["sigma_err"], rtol=1e-7)
    assert n_ok >= N // 2


def test_pipeline_gradient_with_aggregation_and_intercept():
    """Augmented states and a steady-state intercept flow through the gradient: checked against central differences of
    the GPU log-likelihood itself (loose tolerance: ll ~ 1e3, step 1e-6)."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    from test_gpu_augmentation import _aggregated_data

    mod = model("rbc")
    observed, ta = ["Y", "C"], {"Y": "sum"}
    ss = BatchedStateSpace(CompiledModel("rbc")).configure(
        observed_states=observed, measurement_error=observed, tol=1e-13, max_iter=1000, temporal_aggregation=ta, aggregation_period=4,
        ss_obs_intercept=["Y", "C"],
    )  # fmt: skip
    Y = _aggregated_data(mod, observed, ta, 4, ["Y", "C"], 40, seed=8, dense=True)
    th = draws(mod, 3, seed=5, width=0.002, valid=True)
    full = np.hstack([th, np.full((3, mod.k), SIGMA_SHOCK), np.full((3, 2), SIGMA_ERR)])
    ll, grad, st = ss.loglik_and_grad(full, Y)
    assert (st == 0).all()
    for j in range(full.shape[1]):
        h = 1e-6 * np.maximum(1.0, np.abs(full[:, j])) if j < th.shape[1] else 1e-8
        fp, fm = full.copy(), full.copy()
        fp[:, j] += h
        fm[:, j] -= h
        num = (ss.loglik(fp, Y)[0] - ss.loglik(fm, Y)[0]) / (2 * h)
        assert np.abs(num - grad[:, j]).max() <= 2e-4 * max(1.0, np.abs(num).max()), (j, num, grad[:, j])


def test_batched_hmc_on_the_statespace_posterior():
    """Many HMC chains in lock-step, each leapfrog step one batched loglik_and_grad call: the integrator conserves energy
    with the GPU gradient (i.e. gradient and likelihood are consistent along whole trajectories), chains started at prior
    draws climb towards the data-generating parameters, and everything stays inside the prior box."""
    import torch

    from geconpy_b200.hmc import BatchedHMC, statespace_target
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("rbc")
    ss = BatchedStateSpace(CompiledModel("rbc")).configure(observed_states=["Y"], measurement_error=["Y"], tol=1e-10, max_iter=200)
    Y = simulate_obs(mod, 80, seed=3, sigma_err=SIGMA_ERR)
    th_true = torch.as_tensor(mod.theta_vector(), device="cuda")
    N = 512
    gen = torch.Generator(device="cuda").manual_seed(0)
    lo, hi = th_true * 0.97, th_true * 1.03
    lo[mod.param_names.index("beta")], hi[mod.param_names.index("beta")] = 0.985, 0.995
    th0 = lo + (hi - lo) * torch.rand((N, th_true.numel()), generator=gen, dtype=torch.float64, device="cuda")
    tail = torch.tensor([[SIGMA_SHOCK, SIGMA_ERR]], dtype=torch.float64, device="cuda")
    target = statespace_target(ss, torch.as_tensor(Y, device="cuda"), tail)
    hmc = BatchedHMC(target, lo, hi, step_scale=0.004, n_leapfrog=4, seed=5).initialise(th0)
    lp_start = float(hmc.logp[torch.isfinite(hmc.logp)].mean())
    stats = hmc.run(12)
    assert all(s.max_energy_error < 0.5 for s in stats), [s.max_energy_error for s in stats]
    assert all(s.accept_rate > 0.7 for s in stats), [s.accept_rate for s in stats]
    assert stats[-1].mean_logp > lp_start and stats[-1].n_failed == 0
    assert bool(((hmc.theta >= lo) & (hmc.theta <= hi)).all())


def test_pipeline_gradient_through_observation_equations():
    """Parameter-dependent design matrix and intercept: dll/dZ from the reverse sweep times the generated VJP of the
    observation kernel, checked against central differences of the GPU log-likelihood itself."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("rbc")
    observed = ["C", "mix"]
    eqs = {"mix": "alpha * log(Y[]) + (1 - alpha) * log(K[-2]) - log(A[ss])"}
    ss = BatchedStateSpace(CompiledModel("rbc")).configure(
        observed_states=observed, measurement_error=observed, tol=1e-13, max_iter=1000, observation_equations=eqs,
        temporal_aggregation={"mix": "mean"}, aggregation_period=2,
    )  # fmt: skip
    th = draws(mod, 3, seed=5, width=0.002, valid=True)
    rng = np.random.default_rng(2)
    ref0 = oss.loglik_augmented(mod, th[0], np.zeros((1, 2)), observed, [SIGMA_SHOCK], [5e-3, 5e-3], temporal_aggregation={"mix": "mean"},
                                aggregation_period=2, observation_equations=eqs, tol=1e-9, max_iter=200)  # fmt: skip
    x, Y = np.zeros(ref0["T_aug"].shape[0]), np.zeros((40, 2))
    for t_ in range(40):
        x = ref0["T_aug"] @ x + ref0["R_aug"] @ (SIGMA_SHOCK * rng.standard_normal(mod.k))
        Y[t_] = ref0["d"] + ref0["Z"] @ x + 5e-3 * rng.standard_normal(2)
    full = np.hstack([th, np.full((3, mod.k), SIGMA_SHOCK), np.full((3, 2), 5e-3)])
    ll, grad, st = ss.loglik_and_grad(full, Y)
    assert (st == 0).all()
    ll_fwd, _ = ss.loglik(full, Y)
    assert np.abs(ll - ll_fwd).max() <= 1e-7
    for j in range(full.shape[1]):
        h = 1e-6 * np.maximum(1.0, np.abs(full[:, j])) if j < th.shape[1] else 1e-8
        fp, fm = full.copy(), full.copy()
        fp[:, j] += h
        fm[:, j] -= h
        num = (ss.loglik(fp, Y)[0] - ss.loglik(fm, Y)[0]) / (2 * h)
        assert np.abs(num - grad[:, j]).max() <= 2e-4 * max(1.0, np.abs(num).max()), (j, num, grad[:, j])


def test_pipeline_gradient_with_a_full_shock_covariance():
    """full_shock_covariance=True (statespace.py:245-249): the parameter vector carries state_cov[i, j], the
    gradient its symmetrised derivative.  Checked against central differences of the GPU log-likelihood along SYMMETRIC directions
    of Q (E_ab + E_ba: the only directions a covariance can move in), and the diagonal-covariance path as a special case."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("full_nk")
    observed = mod.spec["observed_default"]
    cm = CompiledModel("full_nk")
    kw = dict(observed_states=observed, measurement_error=observed, tol=1e-13, max_iter=1000)
    ss = BatchedStateSpace(cm).configure(full_shock_covariance=True, **kw)
    ss_d = BatchedStateSpace(cm).configure(**kw)
    N, k, nt = 3, mod.k, mod.theta_vector().size
    th = draws(mod, N, seed=12, width=0.02, valid=True)
    Y = simulate_obs(mod, 40, seed=3, sigma_err=SIGMA_ERR)
    rng = np.random.default_rng(2)
    Lq = SIGMA_SHOCK * (np.eye(k) + 0.3 * rng.standard_normal((N, k, k)))
    Q = np.einsum("nij,nkj->nik", Lq, Lq)
    err = np.full((N, len(observed)), SIGMA_ERR)
    full = np.hstack([th, Q.reshape(N, -1), err])
    ll, grad, st = ss.loglik_and_grad(full, Y)
    assert (st == 0).all() and grad.shape == (N, nt + k * k + len(observed))
    gQ = grad[:, nt : nt + k * k].reshape(N, k, k)
    assert np.abs(gQ - gQ.transpose(0, 2, 1)).max() <= 1e-12 * np.abs(gQ).max()
    for a, b in ((0, 0), (1, 1), (0, 1), (2, 3), (1, 3)):
        E = np.zeros((k, k))
        E[a, b] = E[b, a] = 1.0
        h = 1e-6 * Q[:, a, a].mean()
        fp, fm = full.copy(), full.copy()
        fp[:, nt : nt + k * k] += h * E.ravel()
        fm[:, nt : nt + k * k] -= h * E.ravel()
        num = (ss.loglik(fp, Y)[0] - ss.loglik(fm, Y)[0]) / (2 * h)
        want = (gQ * E[None]).sum(axis=(1, 2))
        assert np.abs(num - want).max() <= 1e-4 * np.abs(want).max(), (a, b, num, want)
    # a diagonal Q through the full-covariance path == the sigma path: same ll, d/d sigma = 2 sigma d/d Q_cc, same theta gradient
    sig = np.full((N, k), SIGMA_SHOCK) * (1.0 + 0.1 * np.arange(N))[:, None]
    Qd = np.stack([np.diag(s_**2) for s_ in sig])
    ll_f, g_f, _ = ss.loglik_and_grad(np.hstack([th, Qd.reshape(N, -1), err]), Y)
    ll_s, g_s, _ = ss_d.loglik_and_grad(np.hstack([th, sig, err]), Y)
    assert np.abs(ll_f - ll_s).max() <= 1e-7
    assert _close(g_f[:, :nt], g_s[:, :nt], rtol=1e-8)
    dQ = g_f[:, nt : nt + k * k].reshape(N, k, k)
    assert _close(2 * sig * np.diagonal(dQ, axis1=1, axis2=2), g_s[:, nt : nt + k], rtol=1e-8)
