"""Gradient path, CPU side: the oracle's adjoints are pinned against finite differences, and the SAME
source the GPU runs (csrc/grad.cuh, the generated vector-Jacobian product) is compiled with g++ -DGECON_HOST_CHECK and
checked against the oracle -- so the arithmetic of the kernels is verified without a GPU."""

from __future__ import annotations

import ctypes as C
import subprocess

from pathlib import Path

import numpy as np
import pytest

from helpers import SIGMA_ERR, SIGMA_SHOCK, draws, jacobian_batch, model, simulate_obs
from oracle import adjoints as oad
from oracle import solvers as osol
from oracle import statespace as oss

from geconpy_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent


def _fd(x, fun, eps=1e-6):
    out = np.zeros_like(x)
    for i in np.ndindex(*x.shape):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps
        xm[i] -= eps
        out[i] = (fun(xp) - fun(xm)) / (2 * eps)
    return out


def _random_filter(rng, N, n, k, p, Tobs, missing):
    T = rng.standard_normal((N, n, n))
    for i in range(N):
        T[i] *= 0.8 / np.abs(np.linalg.eigvals(T[i])).max()
    R = rng.standard_normal((N, n, k))
    q, h = 0.5 + rng.random((N, k)), 0.1 + rng.random((N, p))
    Z = rng.standard_normal((p, n))
    Y = rng.standard_normal((Tobs, p))
    d = 0.1 * rng.standard_normal((N, p))
    if missing:  # the filter does not mask d, so keep d = 0 where entries are missing
        Y[3, 0], Y[5, p - 1] = np.nan, -9999.0
        Y[7] = np.nan
        d[:] = 0.0
    return T, R, q, h, Z, Y, d


@pytest.mark.parametrize("missing", [False, True])
def test_oracle_kalman_adjoints_match_finite_differences(missing):
    rng = np.random.default_rng(0)
    T, R, q, h, Z, Y, d = (x[0] if x.ndim == 3 or (x.ndim == 2 and x.shape[0] == 1) else x for x in _random_filter(rng, 1, 5, 2, 2, 25, missing))
    g = oad.kalman_loglik_adjoints(Y, T, R, q, Z, h, d)

    def f(T=T, R=R, q=q, h=h, d=d):
        return oss.kalman_loglik(Y, T, R, np.diag(q), Z, np.diag(h), d=d)

    assert abs(g["ll"] - f()) < 1e-10
    numZ = _fd(Z, lambda z: oss.kalman_loglik(Y, T, R, np.diag(q), z, np.diag(h), d=d))
    assert np.abs(numZ - g["Z"]).max() <= 2e-7 * max(1.0, np.abs(numZ).max())
    for name, x in (("T", T), ("R", R), ("q", q), ("h", h), ("d", d)):
        num = _fd(x, lambda v, name=name: f(**{name: v}))
        assert np.abs(num - g[name]).max() <= 2e-7 * max(1.0, np.abs(num).max()), name


def test_oracle_policy_adjoints_match_finite_differences_and_the_stein_form():
    """The restated Kronecker formula of o1_policy_function_adjoints (shared.py:53-71) IS d<T_bar, T>/d(A, B, C)."""
    mod = model("rbc_extended")
    A, B, Cm, _D = mod.jacobians(mod.theta_vector(), mode="statespace")
    T = osol.cycle_reduction_core(A, B, Cm, max_iter=1000, tol=1e-13)[0]
    rng = np.random.default_rng(1)
    Tb = rng.standard_normal(T.shape)
    kron = oad.policy_adjoints_kron(A, B, Cm, T, Tb)
    stein = oad.policy_adjoints_stein(A, B, Cm, T, Tb)
    for a, b in zip(kron, stein):
        assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()
    mats = [A, B, Cm]
    for which in range(3):
        for i, j in np.argwhere(mats[which] != 0)[:5]:
            def val(x):
                m = [M.copy() for M in mats]
                m[which][i, j] = x
                return (osol.cycle_reduction_core(*m, max_iter=1000, tol=1e-13)[0] * Tb).sum()

            x0, eps = mats[which][i, j], 1e-6
            num = (val(x0 + eps) - val(x0 - eps)) / (2 * eps)
            assert abs(num - kron[which][i, j]) <= 1e-6 * max(1.0, abs(num)), (which, i, j)


def test_oracle_full_gradient_matches_brute_force_finite_differences():
    mod = model("rbc")
    observed = mod.spec["observed_default"]
    th = draws(mod, 2, seed=3, width=0.01, valid=True)[1]
    Y = simulate_obs(mod, 40, seed=3, sigma_err=SIGMA_ERR)
    sig, err = np.full(mod.k, SIGMA_SHOCK), np.full(len(observed), SIGMA_ERR)
    g = oad.loglik_grad(mod, th, Y, observed, sig, err)

    def f(t=th, s=sig, e=err):
        return oss.loglik(mod, t, Y, observed, s, e, tol=1e-13, max_iter=1000)["ll"]

    assert abs(g["ll"] - f()) < 1e-9
    num = _fd(th, lambda t: f(t=t), eps=1e-6)
    assert np.abs(num - g["theta"]).max() <= 1e-6 * max(1.0, np.abs(num).max())
    assert np.abs(_fd(sig, lambda s: f(s=s), eps=1e-8) - g["sigma_shock"]).max() <= 1e-5 * np.abs(g["sigma_shock"]).max()
    assert np.abs(_fd(err, lambda e: f(e=e), eps=1e-9) - g["sigma_err"]).max() <= 1e-4 * np.abs(g["sigma_err"]).max()


# ---------------------------------------------------------------------------------------- the kernels' source on the CPU
@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    so = tmp_path_factory.mktemp("grad_hc") / "libgecon_grad_hostcheck.so"
    src = ROOT / "geconpy_b200" / "csrc" / "grad_host_check.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DGECON_HOST_CHECK", "-o", str(so), str(src)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.gecon_kalman_grad_hostcheck.argtypes = [C.POINTER(L.KalmanGradArgs)]
    lib.gecon_policy_adjoint_hostcheck.argtypes = [C.POINTER(L.PolicyAdjointArgs)]
    return lib


@pytest.mark.parametrize("n,k,p,Tobs,missing,selector", [(5, 2, 2, 25, False, False), (5, 2, 2, 25, True, False), (10, 4, 3, 40, True, True),
                                                          (1, 1, 1, 10, False, True), (17, 3, 8, 12, True, False)])
def test_kalman_grad_source_matches_oracle(hostcheck, n, k, p, Tobs, missing, selector):
    rng = np.random.default_rng(n + p)
    N = 2
    T, R, q, h, Z, Y, d = _random_filter(rng, N, n, k, p, Tobs, missing)
    obs = np.sort(rng.choice(n, size=p, replace=False)).astype(np.int32)
    if selector:
        Z = np.zeros((p, n))
        Z[np.arange(p), obs] = 1.0
    sig, serr = np.sqrt(q), np.sqrt(h)
    ll, st = np.zeros(N), np.zeros(N, np.int32)
    Tb, Rb, qb, hb, db = np.zeros((N, n, n)), np.zeros((N, n, k)), np.zeros((N, k)), np.zeros((N, p)), np.zeros((N, p))
    Zb = np.zeros((N, p, n))
    a = L.KalmanGradArgs(
        struct_size=C.sizeof(L.KalmanGradArgs), T=T.ctypes.data, R=R.ctypes.data, qdiag=sig.ctypes.data, q_stride=k, hdiag=serr.ctypes.data,
        h_stride=p, Z=None if selector else Z.ctypes.data, obs_idx=obs.ctypes.data if selector else None, d=d.ctypes.data, d_stride=p,
        Y=Y.ctypes.data, N=N, n=n, k=k, p=p, Tobs=Tobs, jitter=1e-8, missing_fill=-9999.0, mvn_const_mode=0, lyap_max_iter=0,
        status_in=None, gate_mask=0, sigma_inputs=1, ll=ll.ctypes.data, status=st.ctypes.data, T_bar=Tb.ctypes.data, R_bar=Rb.ctypes.data,
        q_bar=qb.ctypes.data, h_bar=hb.ctypes.data, d_bar=db.ctypes.data, z_stride=0, Z_bar=None if selector else Zb.ctypes.data,
    )  # fmt: skip
    assert hostcheck.gecon_kalman_grad_hostcheck(C.byref(a)) == 0
    for i in range(N):
        g = oad.kalman_loglik_adjoints(Y, T[i], R[i], q[i], Z, h[i], d[i])
        assert st[i] == 0 and abs(ll[i] - g["ll"]) <= 1e-9
        for got, ref in ((Tb[i], g["T"]), (Rb[i], g["R"]), (qb[i], 2 * sig[i] * g["q"]), (hb[i], 2 * serr[i] * g["h"]), (db[i], g["d"])):
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
        if not selector:
            assert np.abs(Zb[i] - g["Z"]).max() <= 1e-10 * max(1.0, np.abs(g["Z"]).max())


def test_kalman_grad_source_with_a_full_shock_covariance(hostcheck):
    """full_shock_covariance=True (statespace.py:245-249): Q is a k x k input, C0 = R Q R'; dll/dQ_ab for every entry, against the
    oracle's adjoint, itself against central differences of the oracle's filter."""
    rng = np.random.default_rng(5)
    N, n, k, p, Tobs = 2, 6, 3, 2, 30
    T, R, _q, h, Z, Y, d = _random_filter(rng, N, n, k, p, Tobs, False)
    Lq = rng.standard_normal((N, k, k))
    Q = np.einsum("nij,nkj->nik", Lq, Lq) + 0.3 * np.eye(k)
    serr = np.sqrt(h)
    ll, st = np.zeros(N), np.zeros(N, np.int32)
    Tb, Rb, Qb, hb, db = np.zeros((N, n, n)), np.zeros((N, n, k)), np.zeros((N, k, k)), np.zeros((N, p)), np.zeros((N, p))
    Zb = np.zeros((N, p, n))
    a = L.KalmanGradArgs(
        struct_size=C.sizeof(L.KalmanGradArgs), T=T.ctypes.data, R=R.ctypes.data, qdiag=None, q_stride=0, hdiag=serr.ctypes.data, h_stride=p,
        Z=Z.ctypes.data, obs_idx=None, d=d.ctypes.data, d_stride=p, Y=Y.ctypes.data, N=N, n=n, k=k, p=p, Tobs=Tobs, jitter=1e-8,
        missing_fill=-9999.0, mvn_const_mode=0, lyap_max_iter=0, status_in=None, gate_mask=0, sigma_inputs=1, ll=ll.ctypes.data,
        status=st.ctypes.data, T_bar=Tb.ctypes.data, R_bar=Rb.ctypes.data, q_bar=None, h_bar=hb.ctypes.data, d_bar=db.ctypes.data, z_stride=0,
        Z_bar=Zb.ctypes.data, qfull=Q.ctypes.data, qfull_stride=k * k, qfull_bar=Qb.ctypes.data,
    )  # fmt: skip
    assert hostcheck.gecon_kalman_grad_hostcheck(C.byref(a)) == 0
    for i in range(N):
        g = oad.kalman_loglik_adjoints(Y, T[i], R[i], None, Z, h[i], d[i], Q=Q[i])
        assert st[i] == 0 and abs(ll[i] - g["ll"]) <= 1e-9
        # (Q is a covariance: only the symmetric part of dll/dQ is defined -- how it splits between Q_ab and Q_ba depends on how a
        # filter is extended to non-symmetric covariances -- so the kernel returns it symmetrised)
        q_ref = 0.5 * (g["Q"] + g["Q"].T)
        assert np.abs(Qb[i] - Qb[i].T).max() <= 1e-13 * np.abs(Qb[i]).max()
        for got, ref in ((Tb[i], g["T"]), (Rb[i], g["R"]), (Qb[i], q_ref), (hb[i], 2 * serr[i] * g["h"]), (db[i], g["d"]), (Zb[i], g["Z"])):
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    g = oad.kalman_loglik_adjoints(Y, T[0], R[0], None, Z, h[0], d[0], Q=Q[0])
    num = _fd(Q[0], lambda q_: oss.kalman_loglik(Y, T[0], R[0], q_, Z, np.diag(h[0]), d=d[0]))
    assert np.abs(0.5 * (num + num.T) - 0.5 * (g["Q"] + g["Q"].T)).max() <= 2e-7 * max(1.0, np.abs(num).max())  # (symmetric parts: see above)


@pytest.mark.parametrize("name,Tobs", [("full_nk", 50), ("nk_complete_more_shocks", 30)])
def test_kalman_grad_source_on_model_matrices_with_tiny_error_variances(hostcheck, name, Tobs):
    """The regime the estimation runs in -- T, R of a solved model, selector Z, error variances 1e-6 next to state variances 1e-4 --
    is where an un-symmetrised rank-p covariance update loses the gradient completely (its antisymmetric mode is not damped); random
    well-conditioned filters do not show it.  Same source as the GPU kernel, against the oracle's Joseph-form adjoint."""
    mod = model(name)
    observed = mod.spec["observed_default"]
    th = draws(mod, 2, seed=31, width=0.02, valid=True)
    Y = simulate_obs(mod, Tobs, seed=3, sigma_err=SIGMA_ERR)
    N, n, k, p = len(th), mod.n, mod.k, len(observed)
    T, R = np.zeros((N, n, n)), np.zeros((N, n, k))
    for i in range(N):
        A, B, Cm, D = mod.jacobians(th[i], mode="statespace")
        Ti = osol.cycle_reduction_core(A, B, Cm, max_iter=1000, tol=1e-13)[0]
        T[i], R[i] = mod.unpermute_policy(Ti, osol.selection_matrix(B, Cm, D, Ti))
    obs = np.array([mod.var_names.index(v) for v in observed], dtype=np.int32)
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    sig, serr, d = np.full((N, k), SIGMA_SHOCK), np.full((N, p), SIGMA_ERR), np.zeros((N, p))
    ll, st = np.zeros(N), np.zeros(N, np.int32)
    Tb, Rb, qb, hb, db = np.zeros((N, n, n)), np.zeros((N, n, k)), np.zeros((N, k)), np.zeros((N, p)), np.zeros((N, p))
    a = L.KalmanGradArgs(
        struct_size=C.sizeof(L.KalmanGradArgs), T=T.ctypes.data, R=R.ctypes.data, qdiag=sig.ctypes.data, q_stride=k, hdiag=serr.ctypes.data,
        h_stride=p, Z=None, obs_idx=obs.ctypes.data, d=d.ctypes.data, d_stride=p, Y=Y.ctypes.data, N=N, n=n, k=k, p=p, Tobs=Tobs, jitter=1e-8,
        missing_fill=-9999.0, mvn_const_mode=0, lyap_max_iter=0, status_in=None, gate_mask=0, sigma_inputs=1, ll=ll.ctypes.data,
        status=st.ctypes.data, T_bar=Tb.ctypes.data, R_bar=Rb.ctypes.data, q_bar=qb.ctypes.data, h_bar=hb.ctypes.data, d_bar=db.ctypes.data,
        z_stride=0, Z_bar=None,
    )  # fmt: skip
    assert hostcheck.gecon_kalman_grad_hostcheck(C.byref(a)) == 0
    for i in range(N):
        g = oad.kalman_loglik_adjoints(Y, T[i], R[i], sig[i] ** 2, Z, serr[i] ** 2, d[i])
        assert st[i] == 0 and abs(ll[i] - g["ll"]) <= 1e-9
        for got, ref in ((Tb[i], g["T"]), (Rb[i], g["R"]), (qb[i], 2 * sig[i] * g["q"]), (hb[i], 2 * serr[i] * g["h"]), (db[i], g["d"])):
            assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", ["rbc", "rbc_extended", "full_nk", "nk_complete_more_shocks"])
@pytest.mark.parametrize("with_R", [False, True])
def test_policy_adjoint_source_matches_the_kronecker_restatement(hostcheck, name, with_R):
    mod = model(name)
    th = draws(mod, 3, seed=2, width=0.02, valid=True)
    A, B, Cm, D = jacobian_batch(mod, th)
    N, n, k = len(th), mod.n, mod.k
    T, R = np.zeros((N, n, n)), np.zeros((N, n, k))
    for i in range(N):
        T[i] = osol.cycle_reduction_core(A[i], B[i], Cm[i], max_iter=1000, tol=1e-13)[0]
        R[i] = osol.selection_matrix(B[i], Cm[i], D[i], T[i])
    rng = np.random.default_rng(1)
    Tb, Rb = rng.standard_normal((N, n, n)), rng.standard_normal((N, n, k))
    Ab, Bb, Cb, Db, st = np.zeros((N, n, n)), np.zeros((N, n, n)), np.zeros((N, n, n)), np.zeros((N, n, k)), np.zeros(N, np.int32)
    a = L.PolicyAdjointArgs(
        struct_size=C.sizeof(L.PolicyAdjointArgs), A=A.ctypes.data, B=B.ctypes.data, C=Cm.ctypes.data, D=D.ctypes.data if with_R else None,
        T=T.ctypes.data, R=R.ctypes.data if with_R else None, T_bar=Tb.ctypes.data, R_bar=Rb.ctypes.data if with_R else None, N=N, n=n, k=k,
        max_iter=0, A_bar=Ab.ctypes.data, B_bar=Bb.ctypes.data, C_bar=Cb.ctypes.data, D_bar=Db.ctypes.data, status=st.ctypes.data,
    )  # fmt: skip
    assert hostcheck.gecon_policy_adjoint_hostcheck(C.byref(a)) == 0
    for i in range(N):
        tb, b0, c0, d0 = Tb[i], 0.0, 0.0, np.zeros((n, k))
        if with_R:
            b0, c0, d0, tadd = oad.selection_adjoints(B[i], Cm[i], D[i], T[i], R[i], Rb[i])
            tb = tb + tadd
        S, SB, SC = oad.policy_adjoints_kron(A[i], B[i], Cm[i], T[i], tb)
        assert st[i] == 0
        for got, ref in ((Ab[i], S), (Bb[i], SB + b0), (Cb[i], SC + c0), (Db[i], d0)):
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_policy_adjoint_source_flags_a_singular_system(hostcheck):
    n = 4
    A, B, Cm, T, Tb = (np.zeros((1, n, n)) for _ in range(5))
    Ab, Bb, Cb, st = np.zeros((1, n, n)), np.zeros((1, n, n)), np.zeros((1, n, n)), np.zeros(1, np.int32)
    a = L.PolicyAdjointArgs(struct_size=C.sizeof(L.PolicyAdjointArgs), A=A.ctypes.data, B=B.ctypes.data, C=Cm.ctypes.data, T=T.ctypes.data,
                            T_bar=Tb.ctypes.data, N=1, n=n, k=0, A_bar=Ab.ctypes.data, B_bar=Bb.ctypes.data, C_bar=Cb.ctypes.data, status=st.ctypes.data)  # fmt: skip
    hostcheck.gecon_policy_adjoint_hostcheck(C.byref(a))
    assert st[0] == L.ST_SINGULAR and np.isnan(Ab).all() and np.isnan(Cb).all()


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_generated_vjp_matches_finite_differences_of_the_generated_jacobian(tmp_path, name):
    """codegen.vjp_body: reverse mode over the straight-line program, against central differences of the forward code."""
    from geconpy_b200.model.codegen import LinearizedModel, load_spec

    lin = LinearizedModel(load_spec(ROOT / "geconpy_b200" / "model" / "specs" / f"{name}.json"))
    src = tmp_path / f"{name}.cpp"
    src.write_text(lin.cuda_source())
    so = tmp_path / f"{name}.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-DGECON_HOST_CHECK", "-o", str(so), str(src)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.gecon_model_vjp_host_check.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    lib.gecon_model_eval_host_check.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    mod = model(name)
    th = draws(mod, 2, seed=1, width=0.02, valid=True)
    N, n, k, nt = len(th), lin.n, lin.k, lin.n_theta
    rng = np.random.default_rng(0)
    Ab, Bb, Cb = (rng.standard_normal((N, n, n)) for _ in range(3))
    Db, xb = rng.standard_normal((N, n, k)), rng.standard_normal((N, n))
    thb = np.zeros((N, nt))
    lib.gecon_model_vjp_host_check(th.ctypes.data, N, Ab.ctypes.data, Bb.ctypes.data, Cb.ctypes.data, Db.ctypes.data, xb.ctypes.data, thb.ctypes.data)

    def f(t):
        t = np.ascontiguousarray(t[None])
        A, B, Cm, D, x, st = np.zeros((1, n, n)), np.zeros((1, n, n)), np.zeros((1, n, n)), np.zeros((1, n, k)), np.zeros((1, n)), np.zeros(1, np.int32)
        lib.gecon_model_eval_host_check(t.ctypes.data, 1, A.ctypes.data, B.ctypes.data, Cm.ctypes.data, D.ctypes.data, x.ctypes.data, st.ctypes.data)
        return A[0], B[0], Cm[0], D[0], x[0]

    for i in range(N):
        for j in range(nt):
            eps = 1e-6 * max(1.0, abs(th[i, j]))
            tp, tm = th[i].copy(), th[i].copy()
            tp[j] += eps
            tm[j] -= eps
            num = sum(((a - b) * w).sum() for a, b, w in zip(f(tp), f(tm), (Ab[i], Bb[i], Cb[i], Db[i], xb[i]))) / (2 * eps)
            assert abs(num - thb[i, j]) <= 1e-6 * max(1.0, abs(num)), (i, j)
