"""Oracle stage 3: T,R -> P0 -> Kalman log-likelihood, and the full theta -> logp path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

PARITY UNPINNED for the Kalman filter: the reference delegates it to ``pymc_extras`` (third party, >= 0.12.0,
``pyproject.toml:45``; call site ``gEconpy/model/statespace.py:1151-1157``), whose source is not under
``/root/reference`` and which is not installed here.  ``kalman_loglik`` restates the published ``StandardFilter``
algorithm as recorded in SURVEY.md Appendix A.5 (update -> jitter -> predict; Joseph form; log det F; jitter on
both H and the filtered covariance; missing entries masked out of Z and H).  The two details that upstream has
changed across versions are options: ``mvn_const`` ("per_obs": p * log(2 pi) [default]; "bare": log(2 pi)).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import solvers

JITTER_DEFAULT = 1e-8  # pymc_extras.statespace.utils.constants.JITTER_DEFAULT for float64
MISSING_FILL = -9999.0  # pymc_extras.statespace.utils.constants.MISSING_FILL


def dlyap(T, RQR, method="bilinear"):
    """``pt.linalg.solve_discrete_lyapunov(T, R Q R^T, method)`` (gEconpy/model/statespace.py:814-815):
    X = T X T^T + RQR.  pytensor's "bilinear" method is the same bilinear transform + continuous Lyapunov solve
    that scipy implements; "direct" is the Kronecker solve."""
    with np.errstate(all="ignore"):
        try:
            return sla.solve_discrete_lyapunov(T, RQR, method=method)
        except Exception:
            return np.full_like(RQR, np.nan)


def _sym(X):
    return 0.5 * (X + X.T)


def kalman_loglik(
    Y,
    T,
    R,
    Q,
    Z,
    H,
    d=None,
    c=None,
    a0=None,
    P0=None,
    jitter=JITTER_DEFAULT,
    missing_fill=MISSING_FILL,
    mvn_const="per_obs",
    return_all=False,
    mask_intercept=False,
):
    """Standard (covariance-form) Kalman filter log-likelihood, SURVEY.md Appendix A.5.

    ``mask_intercept``: False (A.5 as recorded: y_hat = d + Z_masked a, so a missing entry has v_i = -d_i against
    F_ii = jitter); True: d is masked like Z and H (v_i = 0).  Which one upstream follows is recorded by
    tests/golden/make_kalman_goldens.py wherever pymc_extras is installed.

    Y: (T_obs, p); T: (n,n); R: (n,k); Q: (k,k); Z: (p,n); H: (p,p); d: (p,) or None; c: (n,) or None.
    a0/P0 are the PREDICTED moments of the first observation (a0 = 0, P0 = dlyap by default).
    """
    Y = np.asarray(Y, dtype=np.float64)
    n = T.shape[0]
    p = Z.shape[0]
    d = np.zeros(p) if d is None else np.asarray(d, dtype=np.float64)
    c = np.zeros(n) if c is None else np.asarray(c, dtype=np.float64)
    RQR = _sym(R @ Q @ R.T)
    a = np.zeros(n) if a0 is None else np.asarray(a0, dtype=np.float64).copy()
    P = dlyap(T, R @ Q @ R.T) if P0 is None else np.asarray(P0, dtype=np.float64).copy()
    I_n = np.eye(n)
    I_p = np.eye(p)
    lls = np.zeros(Y.shape[0])
    log2pi = np.log(2.0 * np.pi)
    with np.errstate(all="ignore"):
        for t in range(Y.shape[0]):
            y = Y[t]
            mask = np.isnan(y) | (y == missing_fill)
            all_missing = bool(mask.all())
            W = np.diag((~mask).astype(np.float64))
            Zm = W @ Z
            Hm = W @ H
            ym = np.where(mask, 0.0, y)
            # update
            v = ym - ((W @ d if mask_intercept else d) + Zm @ a)
            PZt = P @ Zm.T
            F = Zm @ PZt + Hm + jitter * I_p
            try:
                cF = sla.cho_factor(F, lower=True, check_finite=False)
                K = sla.cho_solve(cF, PZt.T, check_finite=False).T
                Finv_v = sla.cho_solve(cF, v, check_finite=False)
                logdet = 2.0 * np.log(np.diag(cF[0])).sum()
            except Exception:
                K = np.full((n, p), np.nan)
                Finv_v = np.full(p, np.nan)
                logdet = np.nan
            IKZ = I_n - K @ Zm
            a_f = a + K @ v
            P_f = _sym(IKZ @ P @ IKZ.T) + _sym(K @ Hm @ K.T)
            const = p * log2pi if mvn_const == "per_obs" else log2pi
            lls[t] = 0.0 if all_missing else -0.5 * (const + logdet + v @ Finv_v)
            P_f = P_f + jitter * I_n
            # predict
            a = T @ a_f + c
            P = _sym(T @ P_f @ T.T) + RQR
    if return_all:
        return float(lls.sum()), lls
    return float(lls.sum())


def selector_design(var_names, observed):
    """gEconpy/model/statespace.py:282-296: one-hot selector rows (no aggregation, no obs equations)."""
    Z = np.zeros((len(observed), len(var_names)))
    for i, name in enumerate(observed):
        Z[i, var_names.index(name)] = 1.0
    return Z


def loglik_from_matrices(
    A,
    B,
    C,
    D,
    Y,
    obs_idx,
    sigma_shock,
    sigma_err=None,
    inv_var_order=None,
    permuted_lead_idx=None,
    solver="cycle_reduction",
    tol=1e-8,
    max_iter=1000,
    solver_tol=1e-8,
    jitter=JITTER_DEFAULT,
    check_bk=True,
    check_resid=True,
    mvn_const="per_obs",
    lyapunov_method="bilinear",
    Q_full=None,
):
    """One likelihood evaluation from already-evaluated (permuted) Jacobians.

    Follows gEconpy/model/statespace.py:197-222 (solve, residual in solver order, un-permute),
    :769-770 (BK on permuted lead positions), :800-820 (Q, H, a0 = 0, P0) and :1206-1215 (-inf gating).
    Returns a dict with ll (gated), ll_raw, flags and intermediates.
    """
    n = A.shape[0]
    out = {}
    if solver == "cycle_reduction":
        T, conv, n_iter = solvers.cycle_reduction_core(A, B, C, max_iter=max_iter, tol=tol)
        out["converged"], out["n_iter"] = bool(conv), n_iter
    elif solver == "gensys":
        T, _R, success, eu = solvers.gensys_policy(A, B, C, D, tol=tol)
        if T is None:
            T = np.full((n, n), np.nan)
        out["converged"], out["n_iter"], out["eu"] = bool(success), 0, eu
    else:
        raise ValueError(solver)
    R = solvers.selection_matrix(B, C, D, T)
    resid = solvers.policy_residual(A, B, C, T)
    out["resid"] = resid
    if permuted_lead_idx is not None:
        bk_ok, n_fwd, n_unst = solvers.bk_condition_pt(A, B, C, D, permuted_lead_idx)
    else:
        bk_ok, n_fwd, n_unst = True, 0, 0
    out["bk_ok"], out["n_forward"], out["n_unstable"] = bool(bk_ok), n_fwd, n_unst
    out["T_solver"], out["R_solver"] = T, R
    if inv_var_order is not None:
        T = T[inv_var_order][:, inv_var_order]
        R = R[inv_var_order]
    out["T"], out["R"] = T, R
    # full_shock_covariance: state_cov is used as Q directly (gEconpy/model/statespace.py:245-249)
    Q = np.diag(np.asarray(sigma_shock, dtype=np.float64) ** 2) if Q_full is None else np.asarray(Q_full, dtype=np.float64)
    p = len(obs_idx)
    Z = np.zeros((p, n))
    Z[np.arange(p), obs_idx] = 1.0
    H = np.zeros((p, p))
    if sigma_err is not None:
        H = np.diag(np.asarray(sigma_err, dtype=np.float64) ** 2)
    with np.errstate(all="ignore"):
        P0 = dlyap(T, R @ Q @ R.T, method=lyapunov_method)
        ll_raw = kalman_loglik(Y, T, R, Q, Z, H, P0=P0, jitter=jitter, mvn_const=mvn_const) if np.all(np.isfinite(P0)) else np.nan
    out["P0"] = P0
    out["ll_raw"] = ll_raw
    ok = True
    if check_bk and not bk_ok:
        ok = False
    if check_resid and not (resid < solver_tol):
        ok = False
    out["ok"] = ok
    out["ll"] = ll_raw if ok else -np.inf
    return out


CUMULATOR_AGGREGATIONS = ("sum", "mean")  # gEconpy/model/statespace.py:48


def cumulator_variables(temporal_aggregation):
    """gEconpy/model/statespace.py:561-571 (no observation equations): insertion order of the dict."""
    return [v for v, m in (temporal_aggregation or {}).items() if m in CUMULATOR_AGGREGATIONS]


def augment_transition(T, var_names, temporal_aggregation, aggregation_period):
    """gEconpy/model/statespace.py:598-650: T_aug = [[T, 0], [F, kron(I, shift)]]; F copies each aggregated variable
    into the first slot of its chain, ``shift`` moves the chain down by one slot per period."""
    cum = cumulator_variables(temporal_aggregation)
    if not cum:
        return T
    k_orig = T.shape[0]
    n_lags = aggregation_period - 1
    n_cum = len(cum) * n_lags
    shift = np.zeros((n_lags, n_lags))
    if n_lags > 1:
        shift[np.arange(1, n_lags), np.arange(n_lags - 1)] = 1.0
    Cc = np.kron(np.eye(len(cum)), shift)
    F = np.zeros((n_cum, k_orig))
    for pos, name in enumerate(cum):
        F[pos * n_lags, var_names.index(name)] = 1.0
    return np.block([[T, np.zeros((k_orig, n_cum))], [F, Cc]])


def augment_selection(R, n_extra):
    """gEconpy/model/statespace.py:696-723: zero rows for the deterministic lag copies."""
    return R if n_extra == 0 else np.vstack([R, np.zeros((n_extra, R.shape[1]))])


def design_matrix(var_names, observed, temporal_aggregation, aggregation_period):
    """gEconpy/model/statespace.py:279-296 (selector path): weight 1 (or 1/s for "mean") on the variable's column and
    on its cumulator slots; "first"/"last"/unaggregated variables get a plain selector row."""
    ta = temporal_aggregation or {}
    cum = cumulator_variables(ta)
    n_lags = aggregation_period - 1
    k_orig = len(var_names)
    Z = np.zeros((len(observed), k_orig + len(cum) * n_lags))
    for i, name in enumerate(observed):
        j = var_names.index(name)
        m = ta.get(name)
        if m in CUMULATOR_AGGREGATIONS:
            w = 1.0 / aggregation_period if m == "mean" else 1.0
            start = k_orig + cum.index(name) * n_lags
            Z[i, j] = w
            Z[i, start : start + n_lags] = w
        else:
            Z[i, j] = 1.0
    return Z


def obs_intercept(x_ss, var_names, observed, ss_obs_intercept, log_linearized, temporal_aggregation, aggregation_period):
    """gEconpy/model/statespace.py:363-388: log x_ss (log-linearised) or x_ss (level) for the listed states, zero for the
    others; "sum" aggregation multiplies the per-period intercept by the aggregation period."""
    ta = temporal_aggregation or {}
    d = np.zeros(len(observed))
    for i, name in enumerate(observed):
        if name not in (ss_obs_intercept or []):
            continue
        v = x_ss[var_names.index(name)]
        base = np.log(v) if name in log_linearized else v
        d[i] = aggregation_period * base if ta.get(name) == "sum" else base
    return d


def observation_equation_terms(expr_str, var_names, x_ss, params, log_linearized):
    """Linearisation of a GCN-syntax observation equation (gEconpy/model/statespace.py:446-507) WITHOUT symbolic
    differentiation: every reference v[-k] is replaced by v_ss exp(v~) (log-linearised) or v_ss + v~, the intercept is
    the value at v~ = 0 and each coefficient is d/dv~ there, taken by a complex step (exact to rounding).  Independent of
    the product's sympy path on purpose.  Returns (intercept, {(variable, lag): coefficient})."""
    import re

    refs = {}

    def ref(m):
        v, idx = m.group(1), m.group(2).replace(" ", "")
        if idx == "ss":
            return f"_ss[{v!r}]"
        lag = 0 if idx == "" else int(idx)
        if lag > 0:
            raise ValueError(f"lead reference {v}[{idx}] in an observation equation")
        refs[(v, lag)] = True
        return f"_x[({v!r}, {lag})]"

    code = re.sub(r"([A-Za-z_][A-Za-z_0-9]*)\[([^\]]*)\]", ref, expr_str).replace("^", "**")
    ss = {v: float(x_ss[var_names.index(v)]) for v in var_names}
    env = {"log": np.log, "exp": np.exp, "sqrt": np.sqrt, "_ss": ss, **{k: float(v) for k, v in params.items()}}

    def g(tildes):
        x = {}
        for (v, lag) in refs:
            t = tildes.get((v, lag), 0.0)
            x[(v, lag)] = ss[v] * np.exp(t) if v in log_linearized else ss[v] + t
        return eval(code, {"__builtins__": {}}, {**env, "_x": x})  # noqa: S307 - test infrastructure, fixed expressions

    intercept = float(np.real(g({})))
    h = 1e-30
    coeffs = {key: float(np.imag(g({key: 1j * h})) / h) for key in refs}
    return intercept, coeffs


def obs_lag_layout(coeff_keys, temporal_aggregation, aggregation_period, k_prev):
    """``_obs_lag_depths`` and ``_obs_lag_starts`` (gEconpy/model/statespace.py:1040-1076)."""
    ta = temporal_aggregation or {}
    depths = {}
    for obs_name, keys in coeff_keys.items():
        broadcast = aggregation_period - 1 if ta.get(obs_name) in CUMULATOR_AGGREGATIONS else 0
        for v, lag in keys:
            need = -lag + broadcast
            if need > 0:
                depths[v] = max(depths.get(v, 0), need)
    starts, off = {}, k_prev
    for v, dep in depths.items():
        starts[v] = off
        off += dep
    return depths, starts


def append_obs_lag_block(T_aug, var_names, depths, starts):
    """gEconpy/model/statespace.py:652-694: shift-companion chains for the variables an observation equation lags."""
    n_lag = sum(depths.values())
    if n_lag == 0:
        return T_aug
    k_prev = T_aug.shape[0]
    F, Cc = np.zeros((n_lag, k_prev)), np.zeros((n_lag, n_lag))
    for v, dep in depths.items():
        b0 = starts[v] - k_prev
        F[b0, var_names.index(v)] = 1.0
        for j in range(1, dep):
            Cc[b0 + j, b0 + j - 1] = 1.0
    return np.block([[T_aug, np.zeros((k_prev, n_lag))], [F, Cc]])


def loglik_augmented(
    model, theta, Y, observed, sigma_shock, sigma_err=None, temporal_aggregation=None, aggregation_period=4,
    ss_obs_intercept=None, log_linearized=None, tol=1e-8, max_iter=1000, solver_tol=1e-8, jitter=JITTER_DEFAULT,
    mvn_const="per_obs", observation_equations=None, mask_intercept=False,
):  # fmt: skip
    """theta -> logp with temporal aggregation and steady-state intercepts: make_symbolic_graph's sequence
    (gEconpy/model/statespace.py:769-820) -- solve, un-permute, augment T and R, build Z and d, P0 of the AUGMENTED
    system, filter, gate."""
    names = model.var_names
    eqs = dict(observation_equations or {})
    model_observed = [v for v in observed if v not in eqs]
    base = loglik(model, theta, np.zeros((1, max(1, len(model_observed)))), model_observed or [names[0]], sigma_shock, None, tol=tol,
                  max_iter=max_iter, solver_tol=solver_tol)
    T, R = base["T"], base["R"]
    ta = temporal_aggregation or {}
    loglin = set(names) if log_linearized is None else set(log_linearized)
    x_ss = model.steady_state(theta)
    ta_model = {k: v for k, v in ta.items() if k not in eqs}  # aggregated equations live in the lag block (statespace.py:562-571)
    T_aug = augment_transition(T, names, ta_model, aggregation_period)
    params = dict(zip(model.param_names, np.asarray(theta, dtype=np.float64)))
    params.update(dict(zip(model.spec.get("deterministic_params", {}), model._det_values(theta))))
    terms = {k: observation_equation_terms(e, names, x_ss, params, loglin) for k, e in eqs.items()}
    depths, starts = obs_lag_layout({k: t[1].keys() for k, t in terms.items()}, ta, aggregation_period, T_aug.shape[0])
    T_aug = append_obs_lag_block(T_aug, names, depths, starts)
    R_aug = augment_selection(R, T_aug.shape[0] - T.shape[0])
    Zm = design_matrix(names, model_observed, ta_model, aggregation_period)
    Z = np.zeros((len(observed), T_aug.shape[0]))
    d = np.zeros(len(observed))
    dm = obs_intercept(x_ss, names, model_observed, ss_obs_intercept, loglin, ta_model, aggregation_period)
    for i, name in enumerate(observed):
        if name not in eqs:
            j = model_observed.index(name)
            Z[i, : Zm.shape[1]] = Zm[j]
            d[i] = dm[j]
            continue
        icpt, coeffs = terms[name]
        agg = ta.get(name)
        n_per = aggregation_period if agg in CUMULATOR_AGGREGATIONS else 1
        w = 1.0 / n_per if agg == "mean" else 1.0
        for (v, lag), c in coeffs.items():  # statespace.py:308-322
            for dd in range(n_per):
                eff = lag - dd
                col = names.index(v) if eff == 0 else starts[v] + (-eff - 1)
                Z[i, col] += w * c
        d[i] = aggregation_period * icpt if agg == "sum" else icpt
    Q = np.diag(np.asarray(sigma_shock, dtype=np.float64) ** 2)
    p = len(observed)
    H = np.zeros((p, p)) if sigma_err is None else np.diag(np.asarray(sigma_err, dtype=np.float64) ** 2)
    with np.errstate(all="ignore"):
        P0 = dlyap(T_aug, R_aug @ Q @ R_aug.T)
        ll_raw = kalman_loglik(Y, T_aug, R_aug, Q, Z, H, d=d, P0=P0, jitter=jitter, mvn_const=mvn_const, mask_intercept=mask_intercept) if np.all(np.isfinite(P0)) else np.nan
    out = dict(base)
    out.update(T_aug=T_aug, R_aug=R_aug, Z=Z, d=d, P0=P0, ll_raw=ll_raw, ll=ll_raw if base["ok"] else -np.inf)
    return out


def loglik(model, theta, Y, observed, sigma_shock, sigma_err=None, **kwargs):
    """theta -> logp for an ``oracle.model.OracleModel`` (the path of SURVEY.md section 3.3, without priors)."""
    A, B, C, D = model.jacobians(theta, mode="statespace")
    obs_idx = [model.var_names.index(v) for v in observed]
    return loglik_from_matrices(
        A,
        B,
        C,
        D,
        Y,
        obs_idx,
        sigma_shock,
        sigma_err,
        inv_var_order=model.inv_var_order,
        permuted_lead_idx=model.permuted_lead_var_idx,
        **kwargs,
    )


def simulate(T, R, sigma_shock, n_steps, seed=0, x0=None):
    """x_t = T x_{t-1} + R eps_t (gEconpy/model/simulate.py:171-183), eps ~ N(0, diag(sigma^2))."""
    rng = np.random.default_rng(seed)
    n, k = R.shape
    x = np.zeros(n) if x0 is None else np.asarray(x0, dtype=np.float64)
    out = np.zeros((n_steps, n))
    eps = rng.standard_normal((n_steps, k)) * np.asarray(sigma_shock)
    for t in range(n_steps):
        x = T @ x + R @ eps[t]
        out[t] = x
    return out
