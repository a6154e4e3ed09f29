"""Oracle for the gradient path: adjoints of the policy function and of the Kalman log-likelihood.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

* ``policy_adjoints_kron`` is a line-by-line numpy restatement of ``o1_policy_function_adjoints``
  (gEconpy/solvers/shared.py:12-71): the n^2 x n^2 Kronecker system for the Lagrange multipliers S, then
  A_bar = S, B_bar = S T', C_bar = S T' T'.  The reference function is a pytensor graph and cannot be executed here
  (pytensor is not installed): PARITY UNPINNED by the reference; pinned instead against central finite differences of
  the oracle's own cycle reduction (tests/test_adjoints_cpu.py), which is an independent check of the formula.
* ``selection_adjoints``: reverse mode of ``R = -(C T + B)^-1 D`` (gEconpy/solvers/shared.py:74-75); in the reference
  pytensor differentiates ``pt.linalg.solve`` itself.
* ``kalman_loglik_adjoints``: reverse mode of ``oracle.statespace.kalman_loglik`` with P0 = dlyap(T, R Q R') (the
  graph pytensor differentiates for NUTS, gEconpy/model/statespace.py:812-820,1151-1157).  Pinned against central
  finite differences of ``kalman_loglik``.
"""

from __future__ import annotations

import numpy as np

from . import statespace as oss


def policy_adjoints_kron(A, B, C, T, T_bar, jitter=1e-16):
    """gEconpy/solvers/shared.py:53-71, numpy for pytensor."""
    n = A.shape[0]
    vec_T_bar = T_bar.T.ravel()
    eye = np.eye(n)
    M1 = np.kron(T, C.T)
    M2 = np.kron(eye, T.T @ C.T)
    M3 = np.kron(eye, B.T)
    M = M1 + M2 + M3
    M[np.arange(n * n), np.arange(n * n)] += jitter  # stabilize(), shared.py:6-9
    vec_S = np.linalg.solve(M, -vec_T_bar)
    S = vec_S.reshape((n, n)).T
    return S, S @ T.T, S @ T.T @ T.T


def policy_adjoints_stein(A, B, C, T, T_bar, max_iter=64):
    """The same multipliers without the n^2 x n^2 system (what the GPU kernel does): the Kronecker system is
    W' S + C' S T' = -T_bar with W = C T + B, i.e. the Stein equation S = Q + G S T', G = -W^-T C', Q = -W^-T T_bar,
    solved by doubling S <- S + G_k S T_k', G_{k+1} = G_k^2, T_{k+1} = T_k^2 (rho(G) rho(T) < 1 under Blanchard-Kahn)."""
    W = C @ T + B
    G = -np.linalg.solve(W.T, C.T)
    S = -np.linalg.solve(W.T, T_bar)
    Tk = T.T.copy()
    for _ in range(max_iter):
        inc = G @ S @ Tk
        S = S + inc
        if np.abs(inc).max() <= 1e-17 * max(np.abs(S).max(), 1e-300):
            break
        G = G @ G
        Tk = Tk @ Tk
    return S, S @ T.T, S @ T.T @ T.T


def selection_adjoints(B, C, D, T, R, R_bar):
    """Reverse mode of R = -(C T + B)^-1 D: returns (B_bar, C_bar, D_bar, T_bar contribution)."""
    W = C @ T + B
    D_bar = -np.linalg.solve(W.T, R_bar)
    W_bar = D_bar @ R.T
    return W_bar, W_bar @ T.T, D_bar, C.T @ W_bar


def dlyap_adjoint(T, P0, P0_bar, max_iter=64):
    """Adjoint of P0 = T P0 T' + C0: S = T' S T + P0_bar (doubling); returns (C0_bar, T_bar)."""
    S = P0_bar.copy()
    A = T.copy()
    for _ in range(max_iter):
        inc = A.T @ S @ A
        S = S + inc
        if np.abs(inc).max() <= 1e-17 * max(np.abs(S).max(), 1e-300):
            break
        A = A @ A
    return S, S @ T @ P0.T + S.T @ T @ P0


def kalman_loglik_adjoints(Y, T, R, q, Z, h, d=None, jitter=oss.JITTER_DEFAULT, missing_fill=oss.MISSING_FILL, mvn_const="per_obs",
                           mask_intercept=False, Q=None):
    """ll and its gradient with respect to T (n,n), R (n,k), q (k, shock VARIANCES; or Q (k,k), the full shock covariance, when
    ``Q`` is given), h (p, error VARIANCES), d (p).

    Forward pass as ``oracle.statespace.kalman_loglik`` (a0 = 0, P0 = dlyap(T, R diag(q) R'), d NOT masked), storing the
    predicted moments; then the reverse sweep.  Matrices are treated as unconstrained in the sweep (the forward map
    keeps P symmetric for every input, so the chain rule through the un-symmetrised formulas is exact)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, k, p, Tobs = T.shape[0], R.shape[1], Z.shape[0], Y.shape[0]
    d = np.zeros(p) if d is None else np.asarray(d, dtype=np.float64)
    Qm = np.diag(q) if Q is None else np.asarray(Q, dtype=np.float64)  # Q: full shock covariance (statespace.py:245-249); then q is unused
    C0 = R @ Qm @ R.T
    P0 = oss.dlyap(T, C0)
    I_n, I_p = np.eye(n), np.eye(p)
    log2pi = np.log(2.0 * np.pi)
    a, P = np.zeros(n), P0.copy()
    As, Ps = [], []
    ll = 0.0
    for t in range(Tobs):
        As.append(a.copy())
        Ps.append(P.copy())
        y = Y[t]
        mask = np.isnan(y) | (y == missing_fill)
        w = (~mask).astype(np.float64)
        Zm, Hm, ym = w[:, None] * Z, np.diag(w * h), np.where(mask, 0.0, y)
        v = ym - ((w * d if mask_intercept else d) + Zm @ a)
        PZ = P @ Zm.T
        F = Zm @ PZ + Hm + jitter * I_p
        Finv = np.linalg.inv(F)
        K = PZ @ Finv
        L = I_n - K @ Zm
        if not mask.all():
            const = p * log2pi if mvn_const == "per_obs" else log2pi
            ll += -0.5 * (const + np.linalg.slogdet(F)[1] + v @ Finv @ v)
        a_f = a + K @ v
        P_f = L @ P @ L.T + K @ Hm @ K.T + jitter * I_n
        a = T @ a_f
        P = T @ P_f @ T.T + C0
    # ---- reverse sweep
    T_bar, C0_bar = np.zeros((n, n)), np.zeros((n, n))
    h_bar, d_bar = np.zeros(p), np.zeros(p)
    Z_bar = np.zeros((p, n))
    a_bar, P_bar = np.zeros(n), np.zeros((n, n))  # adjoints of the predicted moments of step t + 1
    for t in range(Tobs - 1, -1, -1):
        a, P = As[t], Ps[t]
        y = Y[t]
        mask = np.isnan(y) | (y == missing_fill)
        w = (~mask).astype(np.float64)
        Zm, Hm, ym = w[:, None] * Z, np.diag(w * h), np.where(mask, 0.0, y)
        v = ym - ((w * d if mask_intercept else d) + Zm @ a)
        PZ = P @ Zm.T
        F = Zm @ PZ + Hm + jitter * I_p
        Finv = np.linalg.inv(F)
        K = PZ @ Finv
        L = I_n - K @ Zm
        a_f = a + K @ v
        P_f = L @ P @ L.T + K @ Hm @ K.T + jitter * I_n
        # predict
        T_bar += P_bar @ T @ P_f.T + P_bar.T @ T @ P_f + np.outer(a_bar, a_f)
        C0_bar += P_bar
        Pf_bar = T.T @ P_bar @ T
        af_bar = T.T @ a_bar
        # log-likelihood term
        e = Finv @ v
        if mask.all():
            F_bar, v_bar = np.zeros((p, p)), np.zeros(p)
        else:
            F_bar, v_bar = -0.5 * (Finv - np.outer(e, e)), -e
        # update
        L_bar = Pf_bar @ L @ P.T + Pf_bar.T @ L @ P
        P_bar = L.T @ Pf_bar @ L
        K_bar = Pf_bar @ K @ Hm.T + Pf_bar.T @ K @ Hm - L_bar @ Zm.T + np.outer(af_bar, v)
        Hm_bar = K.T @ Pf_bar @ K
        a_bar = af_bar.copy()
        v_bar = v_bar + K.T @ af_bar
        PZ_bar = K_bar @ Finv.T
        F_bar = F_bar - K.T @ K_bar @ Finv.T
        PZ_bar = PZ_bar + Zm.T @ F_bar
        Hm_bar = Hm_bar + F_bar
        P_bar = P_bar + PZ_bar @ Zm
        a_bar = a_bar - Zm.T @ v_bar
        d_bar -= (w * v_bar) if mask_intercept else v_bar
        h_bar += w * np.diag(Hm_bar)
        # design matrix: L = I - K Zm, v = ym - d - Zm a, PZ = P Zm', G = Zm PZ   (Zm = diag(w) Z)
        Z_bar += w[:, None] * (-K.T @ L_bar - np.outer(v_bar, a) + PZ_bar.T @ P + F_bar @ PZ.T)
    # ---- P0 = dlyap(T, C0), a0 = 0
    S, T_lyap = dlyap_adjoint(T, P0, P_bar)
    C0_bar += S
    T_bar += T_lyap
    R_bar = C0_bar @ R @ Qm.T + C0_bar.T @ R @ Qm
    Q_bar = R.T @ C0_bar @ R  # every entry of Q treated as an independent input
    return dict(ll=ll, T=T_bar, R=R_bar, q=np.diag(Q_bar).copy(), Q=Q_bar, h=h_bar, d=d_bar, Z=Z_bar)


def loglik_grad(model, theta, Y, observed, sigma_shock, sigma_err=None, tol=1e-13, max_iter=1000, jitter=oss.JITTER_DEFAULT, fd_eps=1e-6):
    """theta -> (ll, dll/dtheta, dll/dsigma_shock, dll/dsigma_err) by chaining the adjoints above; only the last stage
    (dA, dB, dC, dD / dtheta) is taken by central differences of ``OracleModel.jacobians`` (smooth, well scaled: ~1e-9
    relative).  Everything runs in solver order; the filter sees the permuted positions of the observed variables."""
    from . import solvers

    theta = np.asarray(theta, dtype=np.float64)
    A, B, C, D = model.jacobians(theta, mode="statespace")
    T, conv, _ = solvers.cycle_reduction_core(A, B, C, max_iter=max_iter, tol=tol)
    R = solvers.selection_matrix(B, C, D, T)
    n, p = A.shape[0], len(observed)
    obs = model.inv_var_order[[model.var_names.index(v) for v in observed]]
    Z = np.zeros((p, n))
    Z[np.arange(p), obs] = 1.0
    sig = np.asarray(sigma_shock, dtype=np.float64)
    err = np.zeros(p) if sigma_err is None else np.asarray(sigma_err, dtype=np.float64)
    g = kalman_loglik_adjoints(Y, T, R, sig**2, Z, err**2, None, jitter=jitter)
    B1, C1, D_bar, T_add = selection_adjoints(B, C, D, T, R, g["R"])
    S, SB, SC = policy_adjoints_kron(A, B, C, T, g["T"] + T_add)
    bars = (S, SB + B1, SC + C1, D_bar)
    th_bar = np.zeros_like(theta)
    for j in range(theta.size):
        h = fd_eps * max(1.0, abs(theta[j]))
        tp, tm = theta.copy(), theta.copy()
        tp[j] += h
        tm[j] -= h
        Mp, Mm = model.jacobians(tp, mode="statespace"), model.jacobians(tm, mode="statespace")
        th_bar[j] = sum(((a - b) * w).sum() for a, b, w in zip(Mp, Mm, bars)) / (2 * h)
    return dict(ll=g["ll"], theta=th_bar, sigma_shock=2.0 * sig * g["q"], sigma_err=2.0 * err * g["h"], converged=bool(conv))
