"""Batched-over-draws entry points of the hot path (thin layer over the C ABI).

Every function takes either numpy arrays (HOST path: the library copies in, launches, copies out) or torch CUDA
tensors (DEVICE path: pointers + the current torch stream are handed to the ``*_batched`` entry points, nothing is
copied).  Arrays carry a leading draw axis: ``A[N, n, n]``, ``D[N, n, k]``...; 2-D inputs are treated as N = 1.
There is no CPU implementation behind these functions.
"""

from __future__ import annotations

import ctypes as C

from dataclasses import dataclass

import numpy as np

from . import _lib as L

try:  # torch is plumbing (device memory, streams); the host path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Marshal:
    """Keeps converted arrays alive and yields raw pointers for either path."""

    def __init__(self, device: bool):
        self.device = device
        self.keep = []

    def inp(self, x, dtype=np.float64):
        if x is None:
            return None, None
        if self.device:
            tdt = torch.float64 if dtype == np.float64 else torch.int32
            if not _is_torch(x):
                x = torch.as_tensor(np.ascontiguousarray(x, dtype=dtype), device=self.dev)
            if x.dtype != tdt or not x.is_contiguous():
                x = x.to(tdt).contiguous()
            self.keep.append(x)
            return x, x.data_ptr()
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        a = np.ascontiguousarray(x, dtype=dtype)
        self.keep.append(a)
        return a, a.ctypes.data

    def out(self, shape, dtype=np.float64):
        if self.device:
            tdt = torch.float64 if dtype == np.float64 else torch.int32
            t = torch.empty(shape, dtype=tdt, device=self.dev)
            self.keep.append(t)
            return t, t.data_ptr()
        a = np.empty(shape, dtype=dtype)
        self.keep.append(a)
        return a, a.ctypes.data

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream) if self.device else None


def _marshal_for(*arrays) -> _Marshal:
    dev = None
    for a in arrays:
        if _is_torch(a) and a.is_cuda:
            dev = a.device
            break
    m = _Marshal(dev is not None)
    m.dev = dev
    if dev is None:
        L.require_device()
    return m


def _batch3(x):
    """(n,n) -> (1,n,n); returns (array, was_2d)."""
    if x is None:
        return None, False
    if x.ndim == 2:
        return x[None], True
    if x.ndim != 3:
        raise ValueError(f"expected a 2-D or 3-D array, got shape {tuple(x.shape)}")
    return x, False


@dataclass
class CycleReductionResult:
    T: object
    R: object
    status: object
    n_iter: object
    resid: object
    norms: object
    n_unstable: object = None
    solv_norms: object = None

    @property
    def converged(self):
        return (self.status & L.ST_CR_NOT_CONVERGED) == 0


def cr_solve(A, B, C_, D=None, max_iter=1000, tol=1e-9, resid_tol=0.0, unperm=None, subset=None, lead_idx=None,
             solvability_norms=False, trunc_tol=1e-8, col_ranges=None, scan_semantics=False) -> CycleReductionResult:
    """Batched cycle reduction + R + residual (``gecon_cr_solve_*``).

    Reference: ``_cycle_reduction_core`` (gEconpy/solvers/cycle_reduction.py:127-183), ``pt_compute_selection_matrix``
    (solvers/shared.py:74-75), residual (model/statespace.py:213).  ``C_ = None`` solves the backward-looking system
    (solvers/backward_looking.py).  ``unperm``: full permutation applied to the outputs (statespace.py:217-220);
    ``subset``: index list, outputs are the sub-blocks ``T[subset][:, subset]``, ``R[subset]``.
    ``lead_idx``: lead-variable columns; converged draws then carry ``ST_BK_CERTIFIED`` in ``status`` when the kernel
    could prove n_unstable == n_forward (``result.n_unstable`` = n_lead for those draws, -1 otherwise).
    ``col_ranges`` = (lag_lo, lag_hi, lead_lo, lead_hi): the caller's promise that A is zero outside columns
    [lag_lo, lag_hi) and C outside [lead_lo, lead_hi) (``gecon_cr_args.lag_lo`` ...); entries outside are not read.
    ``scan_semantics``: the conventions of the reference's scan twin (cycle_reduction.py:246-294) instead of the numba
    core's: stop as soon as ||A0||_1 < tol, always solve for T, ``n_iter`` = the steps actually taken.
    """
    if unperm is not None and subset is not None:
        raise ValueError("give at most one of unperm and subset")
    m = _marshal_for(A, B, C_, D)
    A, pA = m.inp(A)
    B, pB = m.inp(B)
    C_, pC = m.inp(C_)
    D, pD = m.inp(D)
    A, squeeze = _batch3(A)
    N, n = A.shape[0], A.shape[1]
    k = 0
    if D is not None:
        k = D.shape[-1]
    n_out = 0 if subset is None else len(subset)
    no = n_out or n
    T, pT = m.out((N, no, no))
    R, pR = m.out((N, no, k)) if D is not None else (None, None)
    status, pS = m.out((N,), np.int32)
    n_iter, pI = m.out((N,), np.int32)
    resid, pRes = m.out((N,))
    norms, pNo = m.out((N, 3))
    gather = unperm if subset is None else subset
    _, pU = m.inp(None if gather is None else np.ascontiguousarray(gather, dtype=np.int32), np.int32)
    lead = None if lead_idx is None else np.ascontiguousarray(lead_idx, dtype=np.int32)
    _, pL = m.inp(lead, np.int32)
    nu, pNu = m.out((N,), np.int32) if lead is not None else (None, None)
    sn, pSn = m.out((N, 2)) if solvability_norms else (None, None)
    args = L.CrArgs(
        struct_size=C.sizeof(L.CrArgs), A=pA, B=pB, C=pC, D=pD, N=N, n=n, k=k, max_iter=int(max_iter), tol=float(tol),
        resid_tol=float(resid_tol), unperm=pU, T=pT, R=pR, status=pS, n_iter=pI, resid=pRes, norms=pNo, n_out=n_out,
        n_lead=(0 if lead is None else int(lead.size)), lead_idx=pL, n_unstable=pNu,
        solv_norms=pSn, trunc_tol=float(trunc_tol),
    )  # fmt: skip
    if col_ranges is not None:
        args.lag_lo, args.lag_hi, args.lead_lo, args.lead_hi = (int(v) for v in col_ranges)
    args.scan_semantics = int(bool(scan_semantics))
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_cr_solve_batched(C.byref(args), m.stream()), "gecon_cr_solve_batched")
    else:
        L.check(lib.gecon_cr_solve_host(C.byref(args)), "gecon_cr_solve_host")
    if squeeze:
        return CycleReductionResult(T[0], None if R is None else R[0], status[0], n_iter[0], resid[0], norms[0], None if nu is None else nu[0],
                                    None if sn is None else sn[0])
    return CycleReductionResult(T, R, status, n_iter, resid, norms, nu, sn)


def bk_count(A, B, C_, lead_idx, status=None, max_iter=0, skip_mask=0, n_unstable=None):
    """Batched Blanchard-Kahn count (``gecon_bk_count_*``): returns (n_unstable[N], status[N]).

    Reference: ``check_bk_condition_pt`` (gEconpy/model/perturbation.py:586-625).
    """
    m = _marshal_for(A, B, C_)
    A, pA = m.inp(A)
    B, pB = m.inp(B)
    C_, pC = m.inp(C_)
    A, squeeze = _batch3(A)
    N, n = A.shape[0], A.shape[1]
    lead = np.ascontiguousarray(lead_idx, dtype=np.int32)
    _, pL = m.inp(lead, np.int32)
    nu, pNu = m.out((N,), np.int32) if n_unstable is None else m.inp(n_unstable, np.int32)
    if status is None:
        st, pS = m.out((N,), np.int32)
        acc = 0
    else:
        st, pS = m.inp(status, np.int32)
        acc = 1
    args = L.BkArgs(
        struct_size=C.sizeof(L.BkArgs), A=pA, B=pB, C=pC, N=N, n=n, n_lead=int(lead.size), lead_idx=pL, accumulate=acc,
        max_iter=int(max_iter), n_unstable=pNu, status=pS, skip_mask=int(skip_mask),
    )  # fmt: skip
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_bk_count_batched(C.byref(args), m.stream()), "gecon_bk_count_batched")
    else:
        L.check(lib.gecon_bk_count_host(C.byref(args)), "gecon_bk_count_host")
    if squeeze:
        return nu[0], st[0]
    return nu, st


def dlyap(T, R, qdiag, max_iter=0):
    """Batched discrete Lyapunov solve P = T P T' + R diag(q) R' (``gecon_dlyap_*``): returns (P, status, n_iter).

    Reference call site: gEconpy/model/statespace.py:814-815.
    """
    m = _marshal_for(T, R)
    T, pT = m.inp(T)
    R, pR = m.inp(R)
    T, squeeze = _batch3(T)
    N, n = T.shape[0], T.shape[1]
    k = R.shape[-1]
    q, pq = m.inp(qdiag)
    q_stride = k if q.ndim == 2 else 0
    P, pP = m.out((N, n, n))
    st, pS = m.out((N,), np.int32)
    it, pI = m.out((N,), np.int32)
    args = L.DlyapArgs(
        struct_size=C.sizeof(L.DlyapArgs), T=pT, R=pR, qdiag=pq, q_stride=q_stride, N=N, n=n, k=k, max_iter=int(max_iter),
        accumulate=0, P=pP, status=pS, n_iter=pI,
    )  # fmt: skip
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_dlyap_batched(C.byref(args), m.stream()), "gecon_dlyap_batched")
    else:
        L.check(lib.gecon_dlyap_host(C.byref(args)), "gecon_dlyap_host")
    if squeeze:
        return P[0], st[0], it[0]
    return P, st, it


def propagate(T, R=None, E=None, X0=None, n_steps=None, start_at_x0=False):
    """Batched linear propagation ``X_t = T X_{t-1} + R E_t`` (``gecon_propagate_*``): returns ``out[N, L, n, m]``.

    ``E``: shock panels ``[L, k, m]`` shared by all draws or ``[N, L, k, m]``; ``X0``: ``[N, n, m]`` state before t = 0.
    Reference recursions: ``_simulate_linear_system`` (gEconpy/model/simulate.py:171-183), impulse responses
    (simulate.py:201-318), ``_compute_autocovariance_matrix`` (model/statistics/covariance.py:133-161).
    """
    m_ = _marshal_for(T, R, E, X0)
    T, pT = m_.inp(T)
    T, squeeze = _batch3(T)
    N, n = T.shape[0], T.shape[1]
    pT = T.data_ptr() if m_.device else T.ctypes.data
    R, pR = m_.inp(R)
    k = 0 if R is None else R.shape[-1]
    E, pE = m_.inp(E)
    X0, pX = m_.inp(X0)
    if E is not None:
        if E.ndim not in (3, 4) or E.shape[-2] != k:
            raise ValueError(f"E must be [L, k, m] or [N, L, k, m] with k = {k}; got {tuple(E.shape)}")
        Lsteps, mm = E.shape[-3], E.shape[-1]
        e_stride = Lsteps * k * mm if E.ndim == 4 else 0
        if E.ndim == 4 and E.shape[0] != N:
            raise ValueError("per-draw shocks need one panel per draw")
    else:
        if X0 is None or n_steps is None:
            raise ValueError("without shocks give X0 and n_steps")
        Lsteps, mm, e_stride = int(n_steps), X0.shape[-1], 0
    if X0 is not None and tuple(X0.shape) != (N, n, mm):
        raise ValueError(f"X0 must be [N, n, m] = {(N, n, mm)}; got {tuple(X0.shape)}")
    out, pO = m_.out((N, Lsteps, n, mm))
    args = L.PropagateArgs(struct_size=C.sizeof(L.PropagateArgs), T=pT, R=pR, X0=pX, E=pE, e_stride=e_stride, N=N, n=n, k=k, m=mm,
                           L=Lsteps, start_at_x0=int(bool(start_at_x0)), reserved0=0, out=pO)  # fmt: skip
    lib = L.load_library()
    if m_.device:
        L.check(lib.gecon_propagate_batched(C.byref(args), m_.stream()), "gecon_propagate_batched")
    else:
        L.check(lib.gecon_propagate_host(C.byref(args)), "gecon_propagate_host")
    return out[0] if squeeze else out


def kalman_loglik(
    T,
    R,
    qdiag,
    Y,
    Z=None,
    obs_idx=None,
    hdiag=None,
    d=None,
    P0=None,
    jitter=1e-8,
    missing_fill=-9999.0,
    mvn_const="per_obs",
    status_in=None,
    gate_mask=0,
    return_per_step=False,
    lyap_max_iter=0,
    Q=None,
    mask_intercept=False,
    t_cols=0,
):
    """Batched Kalman-filter log-likelihood (``gecon_kalman_ll_*``): returns (ll[N], status[N][, ll_t[N, Tobs]]).

    Reference call site: gEconpy/model/statespace.py:1151-1157 (pymc_extras StandardFilter); semantics in
    SURVEY.md Appendix A.5.  ``qdiag`` / ``hdiag`` are VARIANCES, per draw (N, k) / (N, p) or shared (k,) / (p,).
    ``Z`` is (p, n) shared or (N, p, n), one design matrix per draw (parameter-dependent observation equations).
    ``Q``: full shock covariance, (k, k) shared or (N, k, k) (``full_shock_covariance``); ``qdiag`` is then ignored.
    ``t_cols`` > 0: the caller's promise that only the first ``t_cols`` columns of T can be non-zero (``gecon_kalman_args.t_cols``).
    """
    if (Z is None) == (obs_idx is None):
        raise ValueError("give exactly one of Z (dense design matrix) and obs_idx (selector)")
    m = _marshal_for(T, R)
    T, pT = m.inp(T)
    R, pR = m.inp(R)
    T, squeeze = _batch3(T)
    N, n = T.shape[0], T.shape[1]
    k = R.shape[-1]
    Ya, pY = m.inp(Y)
    if Ya.ndim == 1:
        Tobs, p = Ya.shape[0], 1
    else:
        Tobs, p = Ya.shape
    q, pq = m.inp(qdiag if Q is None else None)
    Qa, pQ = m.inp(Q)
    h, ph = m.inp(hdiag)
    dd, pd_ = m.inp(d)
    Za, pZ = m.inp(Z)
    _, pO = m.inp(None if obs_idx is None else np.ascontiguousarray(obs_idx, dtype=np.int32), np.int32)
    _, pP0 = m.inp(P0)
    _, pSin = m.inp(status_in, np.int32)
    ll, pll = m.out((N,))
    st, pS = m.out((N,), np.int32)
    llt, pllt = m.out((N, Tobs)) if return_per_step else (None, None)
    args = L.KalmanArgs(
        struct_size=C.sizeof(L.KalmanArgs), T=pT, R=pR, qdiag=pq, q_stride=(k if (q is not None and q.ndim == 2) else 0), hdiag=ph,
        h_stride=(p if (h is not None and h.ndim == 2) else 0), Z=pZ, obs_idx=pO, d=pd_,
        d_stride=(p if (dd is not None and dd.ndim == 2) else 0), Y=pY, P0=pP0, qfull=pQ,
        qfull_stride=(k * k if (Qa is not None and Qa.ndim == 3) else 0), N=N, n=n, k=k, p=p, Tobs=Tobs,
        jitter=float(jitter), missing_fill=float(missing_fill), mvn_const_mode=(0 if mvn_const == "per_obs" else 1),
        lyap_max_iter=int(lyap_max_iter), status_in=pSin, gate_mask=int(gate_mask), ll=pll, status=pS, ll_t=pllt,
        z_stride=(p * n if (Za is not None and Za.ndim == 3) else 0), mask_intercept=int(bool(mask_intercept)), t_cols=int(t_cols),
    )  # fmt: skip
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_kalman_ll_batched(C.byref(args), m.stream()), "gecon_kalman_ll_batched")
    else:
        L.check(lib.gecon_kalman_ll_host(C.byref(args)), "gecon_kalman_ll_host")
    if squeeze:
        ll, st = ll[0], st[0]
        llt = None if llt is None else llt[0]
    return (ll, st, llt) if return_per_step else (ll, st)


def kalman_loglik_grad(T, R, qdiag, Y, Z=None, obs_idx=None, hdiag=None, d=None, jitter=1e-8, missing_fill=-9999.0,
                       mvn_const="per_obs", status_in=None, gate_mask=0, sigma_inputs=False, lyap_max_iter=0, mask_intercept=False,
                       Q=None):
    """Log-likelihood AND its gradient (``gecon_kalman_grad_*``): returns a dict with ``ll`` [N],
    ``status`` [N], ``T`` [N,n,n], ``R`` [N,n,k], ``q`` [N,k], ``h`` [N,p], ``d`` [N,p] -- the derivatives of ll with
    respect to the arguments of the same name (``q``/``h`` w.r.t. the standard deviations when ``sigma_inputs``) -- and,
    for a dense design matrix (shared ``(p, n)`` or one per draw ``(N, p, n)``), ``Z`` [N,p,n].
    ``Q`` ([k, k] or [N, k, k]): full shock covariance instead of ``qdiag`` (then pass ``qdiag=None``); the result carries
    ``Q`` [N,k,k] (symmetrised dll/dQ) instead of ``q``.
    What pytensor differentiates behind ``build_statespace_graph`` (gEconpy/model/statespace.py:812-820,1151-1157)."""
    if (Q is None) == (qdiag is None):
        raise ValueError("give exactly one of qdiag (shock variances / standard deviations) and Q (full shock covariance)")
    if (Z is None) == (obs_idx is None):
        raise ValueError("give exactly one of Z (dense design matrix) and obs_idx (selector)")
    m = _marshal_for(T, R)
    T, pT = m.inp(T)
    R, pR = m.inp(R)
    T, squeeze = _batch3(T)
    N, n = T.shape[0], T.shape[1]
    k = R.shape[-1]
    Ya, pY = m.inp(Y)
    Tobs, p = (Ya.shape[0], 1) if Ya.ndim == 1 else Ya.shape
    q, pq = m.inp(qdiag)
    Qa, pQ = m.inp(Q)
    h, ph = m.inp(hdiag)
    dd, pd_ = m.inp(d)
    Za, pZ = m.inp(Z)
    _, pO = m.inp(None if obs_idx is None else np.ascontiguousarray(obs_idx, dtype=np.int32), np.int32)
    _, pSin = m.inp(status_in, np.int32)
    ll, pll = m.out((N,))
    st, pS = m.out((N,), np.int32)
    Zb, pZb = m.out((N, p, n)) if Za is not None else (None, None)
    Tb, pTb = m.out((N, n, n))
    Rb, pRb = m.out((N, n, k))
    qb, pqb = m.out((N, k)) if Qa is None else (None, None)
    Qb, pQb = m.out((N, k, k)) if Qa is not None else (None, None)
    hb, phb = m.out((N, p))
    db, pdb = m.out((N, p))
    args = L.KalmanGradArgs(
        struct_size=C.sizeof(L.KalmanGradArgs), T=pT, R=pR, qdiag=pq, q_stride=(k if (q is not None and q.ndim == 2) else 0), hdiag=ph,
        h_stride=(p if (h is not None and h.ndim == 2) else 0), Z=pZ, obs_idx=pO, d=pd_,
        d_stride=(p if (dd is not None and dd.ndim == 2) else 0), Y=pY, N=N, n=n, k=k, p=p, Tobs=Tobs, jitter=float(jitter),
        missing_fill=float(missing_fill), mvn_const_mode=(0 if mvn_const == "per_obs" else 1), lyap_max_iter=int(lyap_max_iter),
        status_in=pSin, gate_mask=int(gate_mask), sigma_inputs=int(bool(sigma_inputs)), ll=pll, status=pS, T_bar=pTb, R_bar=pRb,
        q_bar=pqb, h_bar=phb, d_bar=pdb, z_stride=(p * n if (Za is not None and Za.ndim == 3) else 0), Z_bar=pZb,
        mask_intercept=int(bool(mask_intercept)), qfull=pQ, qfull_stride=(k * k if (Qa is not None and Qa.ndim == 3) else 0), qfull_bar=pQb,
    )  # fmt: skip
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_kalman_grad_batched(C.byref(args), m.stream()), "gecon_kalman_grad_batched")
    else:
        L.check(lib.gecon_kalman_grad_host(C.byref(args)), "gecon_kalman_grad_host")
    out = dict(ll=ll, status=st, T=Tb, R=Rb, h=hb, d=db)
    out.update(dict(q=qb) if Qa is None else dict(Q=Qb))
    if Zb is not None:
        out["Z"] = Zb
    if squeeze:
        out = {key: val[0] for key, val in out.items()}
    return out


def policy_adjoints(A, B, C_, T, T_bar, D=None, R=None, R_bar=None, max_iter=0):
    """Reverse mode of the perturbation solution (``gecon_policy_adjoint_*``): returns (A_bar, B_bar, C_bar, D_bar, status).
    ``o1_policy_function_adjoints`` (gEconpy/solvers/shared.py:12-71) when only ``T_bar`` is given; with ``D, R, R_bar``
    the selection matrix ``R = -(C T + B)^-1 D`` (shared.py:74-75) is differentiated too (``D_bar`` is None otherwise)."""
    with_r = R_bar is not None
    if with_r and (D is None or R is None):
        raise ValueError("R_bar needs D and R")
    m = _marshal_for(A, B, C_, T, T_bar)
    A, pA = m.inp(A)
    B, pB = m.inp(B)
    Cc, pC = m.inp(C_)
    T, pT = m.inp(T)
    Tb, pTb = m.inp(T_bar)
    A, squeeze = _batch3(A)
    N, n = A.shape[0], A.shape[1]
    Dd, pD = m.inp(D if with_r else None)
    Rr, pR = m.inp(R if with_r else None)
    Rb, pRb = m.inp(R_bar)
    k = Dd.shape[-1] if with_r else 0
    Ab, pAb = m.out((N, n, n))
    Bb, pBb = m.out((N, n, n))
    Cb, pCb = m.out((N, n, n))
    Db, pDb = m.out((N, n, k)) if with_r else (None, None)
    st, pS = m.out((N,), np.int32)
    args = L.PolicyAdjointArgs(
        struct_size=C.sizeof(L.PolicyAdjointArgs), A=pA, B=pB, C=pC, D=pD, T=pT, R=pR, T_bar=pTb, R_bar=pRb, N=N, n=n, k=k,
        max_iter=int(max_iter), A_bar=pAb, B_bar=pBb, C_bar=pCb, D_bar=pDb, status=pS,
    )  # fmt: skip
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_policy_adjoint_batched(C.byref(args), m.stream()), "gecon_policy_adjoint_batched")
    else:
        L.check(lib.gecon_policy_adjoint_host(C.byref(args)), "gecon_policy_adjoint_host")
    if squeeze:
        Ab, Bb, Cb, st = Ab[0], Bb[0], Cb[0], st[0]
        Db = None if Db is None else Db[0]
    return Ab, Bb, Cb, Db, st


def solve(M, RHS):
    """Batched general solve with partial pivoting (``gecon_solve_*``): returns (X, status)."""
    m = _marshal_for(M, RHS)
    M, pM = m.inp(M)
    RHS, pR = m.inp(RHS)
    M, squeeze = _batch3(M)
    N, n = M.shape[0], M.shape[1]
    mm = RHS.shape[-1]
    X, pX = m.out((N, n, mm))
    st, pS = m.out((N,), np.int32)
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_solve_batched(pM, pR, N, n, mm, pX, pS, m.stream()), "gecon_solve_batched")
    else:
        L.check(lib.gecon_solve_host(pM, pR, N, n, mm, pX, pS), "gecon_solve_host")
    if squeeze:
        return X[0], st[0]
    return X, st


def gemm(A, B, trans_a=False, trans_b=False, alpha=1.0):
    """Batched n x n product alpha * op(A) op(B) on the in-CTA DMMA path (``gecon_gemm_*``)."""
    m = _marshal_for(A, B)
    A, pA = m.inp(A)
    B, pB = m.inp(B)
    A, squeeze = _batch3(A)
    N, n = A.shape[0], A.shape[1]
    out, pC = m.out((N, n, n))
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_gemm_batched(pA, pB, N, n, int(trans_a), int(trans_b), float(alpha), pC, m.stream()), "gecon_gemm_batched")
    else:
        L.check(lib.gecon_gemm_host(pA, pB, N, n, int(trans_a), int(trans_b), float(alpha), pC), "gecon_gemm_host")
    return out[0] if squeeze else out


def real_eig(M, balance=True, sort=True):
    """Eigenvalues of real general matrices ``M`` ((m, m) or (N, m, m); numpy or torch CUDA): ``(re, im, status)``.

    ``gecon_real_eig_*``: balancing, Householder Hessenberg reduction and Francis double-shift QR, one warp per matrix.
    ``sort``: ascending modulus, as ``RealEig.perform`` returns them (gEconpy/pytensorf/real_eig.py:31-36); ties (complex
    pairs) keep the kernel's order, positive imaginary part first."""
    m = _marshal_for(M)
    M, pM = m.inp(M)
    M, squeeze = _batch3(M)
    N, n = M.shape[0], M.shape[1]
    re, pre = m.out((N, n))
    im, pim = m.out((N, n))
    st, pS = m.out((N,), np.int32)
    lib = L.load_library()
    if m.device:
        L.check(lib.gecon_real_eig_batched(pM, N, n, int(bool(balance)), pre, pim, pS, m.stream()), "gecon_real_eig_batched")
    else:
        L.check(lib.gecon_real_eig_host(pM, N, n, int(bool(balance)), pre, pim, pS), "gecon_real_eig_host")
    if sort:
        if m.device:
            idx = torch.argsort(torch.hypot(re, im), dim=1, stable=True)
            re, im = torch.gather(re, 1, idx), torch.gather(im, 1, idx)
        else:
            idx = np.argsort(np.hypot(re, im), axis=1, kind="stable")
            re, im = np.take_along_axis(re, idx, 1), np.take_along_axis(im, idx, 1)
    if squeeze:
        return re[0], im[0], st[0]
    return re, im, st


def kernel_info(which: str, n: int, p: int = 1, Tobs: int = 1) -> dict:
    """Resident CTAs per SM, dynamic shared memory and threads per CTA of a kernel (needs a device)."""
    idx = {"cr_solve": 0, "kalman_ll": 1, "bk_count": 2, "dlyap": 3}[which]
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    L.check(L.load_library().gecon_kernel_info(idx, n, p, Tobs, C.byref(a), C.byref(b), C.byref(c)), "gecon_kernel_info")
    return {"ctas_per_sm": a.value, "smem_bytes": b.value, "threads": c.value}


def fp64_peak() -> dict:
    """Measured fp64 peaks of the current device (``gecon_fp64_peak``): {"dfma_tflops", "dmma_tflops"}."""
    L.require_device()
    a, b = C.c_double(), C.c_double()
    L.check(L.load_library().gecon_fp64_peak(C.byref(a), C.byref(b)), "gecon_fp64_peak")
    return {"dfma_tflops": a.value, "dmma_tflops": b.value}


def launch_count() -> int:
    return int(L.load_library().gecon_launch_count())
