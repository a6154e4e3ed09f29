"""Oracle stage 2: A,B,C,D -> T, R, flags.  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Each function cites the reference lines it restates (paths relative to ``/root/reference``).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla

_FLOAT_ZERO_TOL = 1e-8  # gEconpy/model/perturbation.py:26


# --------------------------------------------------------------------------------------------- helpers
def _l1(M):
    """Induced 1-norm = max absolute column sum (``np.linalg.norm(M, ord=1)``); NaN-propagating."""
    return np.abs(M).sum(axis=0).max() if M.size else 0.0


def _solve_nanfill(M, rhs):
    """``_solve_gen`` of the numba path NaN-fills instead of raising (cycle_reduction.py:179-181)."""
    with np.errstate(all="ignore"):
        try:
            lu, piv = sla.lu_factor(M, check_finite=False)
            if not np.all(np.isfinite(lu)) or np.any(np.diag(lu) == 0.0):
                return np.full_like(rhs, np.nan, dtype=np.float64)
            return sla.lu_solve((lu, piv), rhs, check_finite=False)
        except Exception:
            return np.full_like(rhs, np.nan, dtype=np.float64)


# --------------------------------------------------------------------------------------------- cycle reduction
def cycle_reduction_core(A0, A1, A2, max_iter=1000, tol=1e-9):
    """gEconpy/solvers/cycle_reduction.py:127-183 ``_cycle_reduction_core`` (the numba core; authoritative flags).

    Returns (T, converged, n_iter) where n_iter is the number of iterations executed.
    T = 0 when not converged; NaN when the last solve hits a singular matrix.
    """
    A0 = np.asarray(A0, dtype=np.float64)
    A1 = np.asarray(A1, dtype=np.float64)
    A2 = np.asarray(A2, dtype=np.float64)
    A0_initial = A0
    A1_hat = A1
    converged = False
    n_iter = 0
    with np.errstate(all="ignore"):
        for _ in range(int(max_iter)):
            n_iter += 1
            lu, piv = sla.lu_factor(A1, check_finite=False)
            X0 = sla.lu_solve((lu, piv), A0, check_finite=False)
            X2 = sla.lu_solve((lu, piv), A2, check_finite=False)
            m00 = A0 @ X0
            m02 = A0 @ X2
            m20 = A2 @ X0
            m22 = A2 @ X2
            A1 = A1 - m02 - m20
            A1_hat = A1_hat - m20
            A0 = -m00
            A2 = -m22
            a0 = _l1(A0)
            if a0 < tol:
                if _l1(A2) < tol:
                    converged = True
                    break
            elif np.isnan(a0):
                break
        T = -_solve_nanfill(A1_hat, A0_initial) if converged else np.zeros_like(A0_initial)
    return T, converged, n_iter


def cycle_reduction_scan(A, B, C, max_iter=50, tol=1e-7):
    """gEconpy/solvers/cycle_reduction.py:246-294 ``_scan_cycle_reduction`` (the scan twin) as a plain loop.

    A fixed ``max_iter``-step scan whose step is ``ifelse(norm < tol, noop, cycle_step)``: only ||A0||_1 is tested, the
    step counter advances on real steps only, 1e-16 is added to the diagonal before every solve (shared.py:6-9), and
    ``T = -solve(stabilize(A1_hat), A)`` is ALWAYS computed (there is no convergence flag).  Returns (T, n_steps)."""
    A0, A1, A2 = (np.asarray(x, dtype=np.float64) for x in (A, B, C))
    A1_hat = A1
    n = A0.shape[0]
    norm, n_steps = 1e9, 0
    eye = np.eye(n) * 1e-16
    with np.errstate(all="ignore"):
        for _ in range(int(max_iter)):
            if norm < tol:
                continue
            tmp = np.vstack([A0, A2]) @ _solve_nanfill(A1 + eye, np.hstack([A0, A2]))
            A1 = A1 - tmp[:n, n:] - tmp[n:, :n]
            A1_hat = A1_hat - tmp[n:, :n]
            A0, A2 = -tmp[:n, :n], -tmp[n:, n:]
            norm = _l1(A0)
            n_steps += 1
        T = -_solve_nanfill(A1_hat + eye, np.asarray(A, dtype=np.float64))
    return T, n_steps


def cycle_reduction_numpy(A0, A1, A2, max_iter=1000, tol=1e-7):
    """gEconpy/solvers/cycle_reduction.py:23-114 ``cycle_reduction_numpy`` (result-string twin).

    Same iteration; differs from the core at the edge described in SURVEY.md fact 6: if the loop runs out
    with ||A0|| < tol but ||A2|| >= tol on the last pass, it returns the failure tuple; if A0 never passed
    it fails too; a NaN norm fails.  Returns (X|None, res|None, msg, log_norm).
    """
    result = "Optimization successful"
    log_norm = 0
    A0 = np.asarray(A0, dtype=np.float64)
    A1 = np.asarray(A1, dtype=np.float64)
    A2 = np.asarray(A2, dtype=np.float64)
    A0_i, A1_i, A2_i, A1_hat = A0, A1, A2, A1
    with np.errstate(all="ignore"):
        for i in range(int(max_iter)):
            X = np.linalg.solve(A1, np.hstack((A0, A2)))
            n = A0.shape[0]
            X0, X2 = X[:, :n], X[:, n:]
            m00, m02, m20, m22 = A0 @ X0, A0 @ X2, A2 @ X0, A2 @ X2
            A1 = A1 - m02 - m20
            A0 = -m00
            A2 = -m22
            A1_hat = A1_hat - m20
            a0 = _l1(A0)
            if a0 < tol:
                if _l1(A2) < tol:
                    break
            elif np.isnan(a0) or i == (max_iter - 1):
                if a0 < tol:
                    result = "Iteration on matrix A0 and A1 converged towards a solution, but A2 did not."
                    log_norm = np.log(_l1(A2))
                else:
                    result = "Iteration on all matrices failed to converged"
                    log_norm = np.log(_l1(A1))
                return None, None, result, log_norm
        Xs = -np.linalg.solve(A1_hat, A0_i)
        res = A0_i + A1_i @ Xs + A2_i @ Xs @ Xs
    return Xs, res, result, log_norm


def selection_matrix(B, C, D, T):
    """gEconpy/solvers/shared.py:74-75: R = -(C T + B)^{-1} D."""
    return -_solve_nanfill(C @ T + B, D)


def policy_residual(A, B, C, T):
    """gEconpy/model/statespace.py:213: sum of squares of A + B T + C T T (solver/permuted order)."""
    with np.errstate(all="ignore"):
        return float(np.square(A + B @ T + C @ T @ T).sum())


def backward_direct(A, B, C, D):
    """gEconpy/solvers/backward_looking.py:8-133: T = solve(-B, A), R = -solve(B, D) for models with C == 0."""
    T = _solve_nanfill(-B, A)
    R = -_solve_nanfill(B, D)
    return T, R


# --------------------------------------------------------------------------------------------- gensys
def gensys_setup(A, B, C, D, tol=1e-8):
    """gEconpy/solvers/gensys.py:568-614 ``_gensys_setup``: (G0 = -Gamma0_sel, Gamma1_sel, c, Psi, Pi)."""
    n = A.shape[0]
    k = D.shape[1]
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > tol)
    sel = np.concatenate((np.arange(n), lead + n))
    Z = np.zeros((n, n))
    I = np.eye(n)
    G0 = np.block([[B, C], [-I, Z]])
    G1 = np.block([[A, Z], [Z, I]])
    Pi = np.vstack((Z, I))
    Psi = np.vstack((D, np.zeros((n, k))))
    G0 = G0[sel][:, sel]
    G1 = G1[sel][:, sel]
    return -G0, G1, np.zeros((n + lead.size, 1)), Psi[sel], Pi[sel][:, lead]


def _svd_keep(M, realsmall):
    u, s, vh = sla.svd(M, full_matrices=False, check_finite=False)
    keep = np.flatnonzero(s > realsmall)
    return u[:, keep], s[keep], vh.conj().T[:, keep]


def gensys(g0, g1, c, psi, pi, tol=1e-8):
    """gEconpy/solvers/gensys.py:190-395 ``_gensys_core`` with scipy's ``ordqz`` standing in for the numba ``gges``.

    Returns (G1, impact, eu, gev) -- the outputs the hot path and its tests consume.  ``eu`` is a list of 3 ints;
    on coincident zeros eu = [-2, -2, 0] and G1, impact are None.
    """
    n = g1.shape[0]
    realsmall = tol if tol > 0 else np.spacing(1)
    AA, BB, alpha, beta, Q_raw, Z = sla.ordqz(
        g0.astype(np.complex128), g1.astype(np.complex128), sort="ouc", output="complex", check_finite=False
    )
    Q = Q_raw.conj().T
    abs_a, abs_b = np.abs(alpha), np.abs(beta)
    zxz = bool(np.any((abs_a < realsmall) & (abs_b < realsmall)))
    stable = ((abs_b < realsmall) & (abs_a >= realsmall)) | ((abs_b >= realsmall) & (abs_a > abs_b))
    n_unstable = int((~stable).sum())
    n_stable = n - n_unstable
    gev = np.column_stack((alpha, beta))
    eu = [0, 0, 0]
    if zxz:
        return None, None, [-2, -2, 0], gev

    Q1, Q2 = Q[:n_stable], Q[n_stable:]
    pi_c = pi.astype(np.complex128)
    n_eta = pi.shape[1]
    if n_unstable == 0:
        u_eta, d_eta, v_eta = np.zeros((0, 0), complex), np.zeros(0), np.zeros((n_eta, 0), complex)
    else:
        u_eta, d_eta, v_eta = _svd_keep(Q2 @ pi_c, realsmall)
    if d_eta.size >= n_unstable:
        eu[0] = 1
    if n_unstable == n:
        u_eta_1, d_eta_1, v_eta_1 = np.zeros((0, 0), complex), np.zeros(0), np.zeros((n_eta, 0), complex)
    else:
        u_eta_1, d_eta_1, v_eta_1 = _svd_keep(Q1 @ pi_c, realsmall)

    if v_eta_1.shape[0] == 0 or v_eta_1.shape[1] == 0:
        unique = True
    else:
        loose = v_eta_1 - v_eta @ (v_eta.conj().T @ v_eta_1)
        s = sla.svd(loose, compute_uv=False, check_finite=False) if loose.size else np.zeros(0)
        n_loose = int((s > realsmall * n).sum())
        eu[2] = n_loose
        unique = n_loose == 0
    if unique:
        eu[1] = 1

    vh = v_eta.conj().T
    inner = u_eta @ (vh / d_eta.reshape(-1, 1) if d_eta.size else vh) @ v_eta_1 @ (
        d_eta_1.reshape(-1, 1) * u_eta_1.conj().T if d_eta_1.size else u_eta_1.conj().T
    )
    T_mat = np.column_stack((np.eye(n_stable, dtype=complex), -inner.conj().T))
    G_0 = np.vstack((T_mat @ AA, np.column_stack((np.zeros((n_unstable, n_stable)), np.eye(n_unstable)))))
    rhs = np.vstack((T_mat @ BB, np.zeros((n_unstable, n))))
    with np.errstate(all="ignore"):
        G_1c = _solve_nanfill_complex(G_0, rhs)
        G_1 = (Z @ G_1c @ Z.conj().T).real
        imp_rhs = np.vstack((T_mat @ Q @ psi.astype(complex), np.zeros((n_unstable, psi.shape[1]))))
        impact = (Z @ _solve_nanfill_complex(G_0, imp_rhs)).real
    return G_1, impact, eu, gev


def _solve_nanfill_complex(M, rhs):
    try:
        return sla.solve(M, rhs, check_finite=False)
    except Exception:
        return np.full(rhs.shape, np.nan, dtype=complex)


def gensys_policy(A, B, C, D, tol=1e-8):
    """gEconpy/solvers/gensys.py:617-631,657-666: T = G1[:n,:n], success = eu[0]==1 and eu[1]==1, R via shared.py:74."""
    n = A.shape[0]
    g0, g1, c, psi, pi = gensys_setup(A, B, C, D, tol)
    G1, impact, eu, _gev = gensys(g0, g1, c, psi, pi, tol)
    if G1 is None:
        return None, None, False, eu
    T = np.ascontiguousarray(G1[:n, :n])
    success = eu[0] == 1 and eu[1] == 1
    R = selection_matrix(B, C, D, T)
    return T, R, success, eu


# --------------------------------------------------------------------------------------------- Blanchard-Kahn
def bk_eigenvalues_qz(A, B, C, D, tol=1e-8):
    """gEconpy/model/perturbation.py:412-445 ``compute_bk_eigenvalues`` (numpy / ordqz variant)."""
    # The reference binds the FIRST output of _gensys_setup (which is already -Gamma_0_sel, gensys.py:611)
    # to a variable called Gamma_0 and negates it again, so the pencil it decomposes is (+Gamma_0_sel, Gamma_1_sel):
    # its eigenvalues are the negatives of gensys', with identical moduli.  Restated literally.
    first, G1, *_ = gensys_setup(A, B, C, D, tol)
    AA, BB, *_ = sla.ordqz(-first, G1, sort="ouc", output="complex", check_finite=False)
    lam = np.diag(BB) / (np.diag(AA) + tol)
    lam = lam[np.argsort(np.abs(lam))]
    n_forward = int((np.abs(C).sum(axis=0) > tol).sum())
    return lam.real, lam.imag, n_forward


def bk_condition_qz(A, B, C, D, tol=1e-8):
    """gEconpy/model/perturbation.py:553-560: (satisfied, n_forward, n_unstable)."""
    re, im, n_forward = bk_eigenvalues_qz(A, B, C, D, tol)
    n_unstable = int((np.sqrt(re**2 + im**2) > 1).sum())
    return n_forward == n_unstable, n_forward, n_unstable


def bk_matrix_pt(A, B, C, lead_var_idx):
    """gEconpy/model/perturbation.py:472-504: M = solve(-Gamma0_sel + 1e-8 I, Gamma1_sel)."""
    n = A.shape[0]
    lead = np.asarray(lead_var_idx, dtype=int)
    Z = np.zeros((n, n))
    I = np.eye(n)
    G0 = np.block([[B, C], [-I, Z]])
    G1 = np.block([[A, Z], [Z, I]])
    sel = np.concatenate((np.arange(n), lead + n))
    G0, G1 = G0[sel][:, sel], G1[sel][:, sel]
    G0_reg = -G0 + np.eye(sel.size) * _FLOAT_ZERO_TOL
    return G0_reg, G1


def real_eig(M):
    """gEconpy/pytensorf/real_eig.py:31-36: eigenvalues sorted by ascending modulus, as (re, im)."""
    w = np.linalg.eigvals(M)
    w = w[np.argsort(np.abs(w))]
    return w.real.copy(), w.imag.copy()


def bk_condition_pt(A, B, C, D, lead_var_idx):
    """gEconpy/model/perturbation.py:586-625 ``check_bk_condition_pt``: (bk_ok, n_forward, n_unstable)."""
    G0_reg, G1 = bk_matrix_pt(A, B, C, lead_var_idx)
    with np.errstate(all="ignore"):
        try:
            M = np.linalg.solve(G0_reg, G1)
            re, im = real_eig(M)
        except Exception:
            return False, len(lead_var_idx), -1
    n_unstable = int((np.sqrt(re**2 + im**2) > 1).sum())
    n_forward = len(lead_var_idx)
    return n_forward == n_unstable, n_forward, n_unstable


# --------------------------------------------------------------------------------------------- diagnostics
def residual_norms_statespace(A, B, C, D, T, R, state_var_mask):
    """gEconpy/model/statespace.py:1179-1204: deterministic / stochastic recursion residual norms
    (all matrices in ONE consistent variable order)."""
    P = T[state_var_mask][:, state_var_mask]
    Q = R[state_var_mask]
    A_p = A[:, state_var_mask]
    R_p = T[:, state_var_mask]
    nd = np.linalg.norm(A_p + B @ R_p + C @ R_p @ P)
    ns = np.linalg.norm(B @ R + C @ R_p @ Q + D)
    return float(nd), float(ns)


def gecon_representation_norms(A, B, C, D, T, R, tol=1e-8):
    """gEconpy/model/perturbation.py:321-380 (statespace_to_gEcon_representation) + :287-318 (residual_norms), the
    quantities ``solvability_check`` reports (statistics/perturbation_diagnostics.py:141-153)."""
    n = T.shape[1]
    state_idx = np.where(np.abs(T[np.argmax(np.abs(T), axis=0), np.arange(n)]) >= tol)[0]
    mask = np.isin(np.arange(n), state_idx)
    PP = T.copy()
    PP[np.abs(PP) < tol] = 0
    QQ = R[:n].copy()
    QQ[np.abs(QQ) < tol] = 0
    P, Q = PP[mask][:, mask], QQ[mask]
    A_prime, R_prime, S_prime = A[:, mask], PP[:, mask], QQ
    nd = np.linalg.norm(A_prime + B @ R_prime + C @ R_prime @ P)
    ns = np.linalg.norm(B @ S_prime + C @ R_prime @ Q + D)
    return float(nd), float(ns)


def solvability_one(model, theta, tol=1e-8, max_iter=100, norm_tol=1e-8):
    """statistics/perturbation_diagnostics.py:105-161 ``_check_one_draw`` on an OracleModel:
    (failure_step | None, norm_deterministic, norm_stochastic)."""
    nd = ns = np.nan
    A, B, C, D = model.jacobians(theta, mode="statespace")
    if not all(np.isfinite(M).all() for M in (A, B, C, D)):
        return "steady_state", nd, ns
    T, conv, _ = cycle_reduction_core(A, B, C, max_iter=max_iter, tol=tol)
    if not conv or not np.isfinite(T).all():
        return "perturbation", nd, ns
    R = selection_matrix(B, C, D, T)
    ok, _, _ = bk_condition_pt(A, B, C, D, model.permuted_lead_var_idx)
    if not ok:
        return "blanchard-kahn", nd, ns
    nd, ns = gecon_representation_norms(A, B, C, D, T, R, tol)
    if nd > norm_tol:
        return "deterministic_norm", nd, ns
    if ns > norm_tol:
        return "stochastic_norm", nd, ns
    return None, nd, ns
