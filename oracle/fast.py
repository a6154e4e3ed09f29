"""numba-compiled twin of the oracle's per-draw path -- TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``).

``bench.py``'s CPU arm (``--impl reference`` and the ``cpu_baseline`` leg) times THIS, not the interpreter-bound numpy
oracle: the reference's estimation graph is compiled (pytensor's numba / C linkers calling LAPACK,
gEconpy/solvers/cycle_reduction.py:127-183 is itself ``numba_njit``), so a compiled restatement is the fair CPU
baseline.  Same arithmetic as ``oracle.solvers`` / ``oracle.statespace`` (checked against them in
tests/test_oracle_golden.py); P0 uses the doubling iteration instead of scipy's Schur solve (numba cannot call scipy).
"""

from __future__ import annotations

import numpy as np

from numba import njit


@njit(cache=True)
def _l1(M):
    best = 0.0
    for j in range(M.shape[1]):
        s = 0.0
        for i in range(M.shape[0]):
            s += abs(M[i, j])
        if s > best or s != s:
            best = s
    return best


@njit(cache=True)
def cycle_reduction(A, B, C, max_iter, tol):
    """_cycle_reduction_core (gEconpy/solvers/cycle_reduction.py:127-183): (T, converged, n_iter)."""
    n = A.shape[0]
    A0, A1, A2, A1h = A.copy(), B.copy(), C.copy(), B.copy()
    converged = False
    n_iter = 0
    for _ in range(max_iter):
        n_iter += 1
        X = np.linalg.solve(A1, np.hstack((A0, A2)))
        X0, X2 = np.ascontiguousarray(X[:, :n]), np.ascontiguousarray(X[:, n:])
        m00, m02, m20, m22 = A0 @ X0, A0 @ X2, A2 @ X0, A2 @ X2
        A1 = A1 - m02 - m20
        A1h = A1h - m20
        A0 = -m00
        A2 = -m22
        a0 = _l1(A0)
        if a0 < tol:
            if _l1(A2) < tol:
                converged = True
                break
        elif a0 != a0:
            break
    T = -np.linalg.solve(A1h, A) if converged else np.zeros_like(A)
    return T, converged, n_iter


@njit(cache=True)
def bk_count(A, B, C, lead):
    """check_bk_condition_pt (gEconpy/model/perturbation.py:586-625): number of |eig| > 1 of the regularised pencil."""
    n = A.shape[0]
    m = n + lead.size
    G0 = np.zeros((m, m))
    G1 = np.zeros((m, m))
    G0[:n, :n] = -B
    G1[:n, :n] = A
    for j in range(lead.size):
        G0[:n, n + j] = -C[:, lead[j]]
        G0[n + j, lead[j]] = 1.0
        G1[n + j, n + j] = 1.0
    for i in range(m):
        G0[i, i] += 1e-8
    M = np.linalg.solve(G0, G1)
    w = np.linalg.eigvals(M.astype(np.complex128))
    cnt = 0
    for i in range(m):
        if abs(w[i]) > 1.0:
            cnt += 1
    return cnt


@njit(cache=True)
def dlyap(T, RQR, max_iter=64):
    P = RQR.copy()
    Ak = T.copy()
    for _ in range(max_iter):
        D = Ak @ P @ Ak.T
        P = P + D
        if np.abs(D).max() <= 1e-16 * np.abs(P).max():
            break
        Ak = Ak @ Ak
    return 0.5 * (P + P.T)


@njit(cache=True)
def kalman(Y, T, R, q, obs_idx, h, jitter, missing_fill):
    """Standard filter of SURVEY.md Appendix A.5 with a selector design matrix and diagonal Q, H."""
    n, p = T.shape[0], obs_idx.size
    Q = np.diag(q)
    RQR = R @ Q @ R.T
    RQR = 0.5 * (RQR + RQR.T)
    P = dlyap(T, R @ Q @ R.T)
    a = np.zeros(n)
    I_n = np.eye(n)
    ll = 0.0
    log2pi = np.log(2.0 * np.pi)
    Z = np.zeros((p, n))
    for i in range(p):
        Z[i, obs_idx[i]] = 1.0
    H = np.diag(h)
    for t in range(Y.shape[0]):
        W = np.zeros((p, p))
        y = np.zeros(p)
        n_missing = 0
        for i in range(p):
            v = Y[t, i]
            if v != v or v == missing_fill:
                n_missing += 1
            else:
                W[i, i] = 1.0
                y[i] = v
        Zm = W @ Z
        Hm = W @ H
        v = y - Zm @ a
        PZt = P @ Zm.T
        F = Zm @ PZt + Hm + jitter * np.eye(p)
        L = np.linalg.cholesky(F)
        K = np.linalg.solve(F, PZt.T).T
        Fv = np.linalg.solve(F, v)
        logdet = 0.0
        for i in range(p):
            logdet += 2.0 * np.log(L[i, i])
        IKZ = I_n - K @ Zm
        a_f = a + K @ v
        Pf = IKZ @ P @ IKZ.T
        KHK = K @ Hm @ K.T
        Pf = 0.5 * (Pf + Pf.T) + 0.5 * (KHK + KHK.T) + jitter * I_n
        if n_missing < p:
            ll += -0.5 * (p * log2pi + logdet + v @ Fv)
        a = T @ a_f
        TP = T @ Pf @ T.T
        P = 0.5 * (TP + TP.T) + RQR
    return ll


@njit(cache=True)
def loglik_from_matrices(A, B, C, D, Y, obs_idx, sigma, sigma_err, inv_var_order, lead, max_iter, tol, solver_tol, jitter):
    """One evaluation from (permuted) Jacobians: solve, R, residual gate, BK gate, un-permute, P0, filter.
    Returns (ll or -inf, n_iter)."""
    n = A.shape[0]
    T, conv, n_iter = cycle_reduction(A, B, C, max_iter, tol)
    R = -np.linalg.solve(C @ T + B, D)
    E = A + B @ T + C @ T @ T
    resid = (E * E).sum()
    nu = bk_count(A, B, C, lead)
    if nu != lead.size or not (resid < solver_tol):
        return -np.inf, n_iter
    Tu = np.empty((n, n))
    Ru = np.empty((n, R.shape[1]))
    for i in range(n):
        Ru[i, :] = R[inv_var_order[i], :]
        for j in range(n):
            Tu[i, j] = T[inv_var_order[i], inv_var_order[j]]
    return kalman(Y, Tu, Ru, sigma * sigma, obs_idx, sigma_err * sigma_err, jitter, -9999.0), n_iter


def loglik(model, theta, Y, observed, sigma_shock, sigma_err=None, tol=1e-8, max_iter=100, solver_tol=1e-8, jitter=1e-8):
    """theta -> gated log-likelihood with the compiled kernels above (Jacobian evaluation stays sympy-lambdified)."""
    A, B, C, D = model.jacobians(theta, mode="statespace")
    if not (np.isfinite(A).all() and np.isfinite(B).all() and np.isfinite(C).all() and np.isfinite(D).all()):
        return -np.inf
    obs_idx = np.array([model.var_names.index(v) for v in observed], dtype=np.int64)
    herr = np.zeros(len(observed)) if sigma_err is None else np.asarray(sigma_err, dtype=np.float64)
    try:
        ll, _ = loglik_from_matrices(
            A, B, C, D, np.ascontiguousarray(Y, dtype=np.float64), obs_idx, np.asarray(sigma_shock, dtype=np.float64), herr,
            model.inv_var_order.astype(np.int64), model.permuted_lead_var_idx.astype(np.int64), int(max_iter), float(tol),
            float(solver_tol), float(jitter),
        )  # fmt: skip
    except Exception:  # LAPACK failure inside numba (singular / not PD): the reference gates such draws to -inf
        return -np.inf
    return float(ll) if np.isfinite(ll) else -np.inf
