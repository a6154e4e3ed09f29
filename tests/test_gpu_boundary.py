"""GPU tests of the drop-in boundary added in round 2: eigenvalue kernel and the functions built on it, the Op
classes executed through the pytensor shim, scan / backward-direct / constant-parameter options of ``configure``, the
structured ``gensys`` entry point, and the two solver kernels against each other."""

from __future__ import annotations

import os

import numpy as np
import pytest

import pt_shim

from helpers import SIGMA_ERR, SIGMA_SHOCK, draws, jacobian_batch, model, rel_fro, simulate_obs
from oracle import solvers as osol
from oracle import statespace as oss

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from geconpy_b200 import batched

    return batched


def _match_spectra(got, ref, tol):
    """Greedy one-to-one matching of two eigenvalue multisets; returns the largest distance."""
    ref = list(ref)
    worst = 0.0
    for z in got:
        j = int(np.argmin([abs(z - r) for r in ref]))
        worst = max(worst, abs(z - ref.pop(j)))
    return worst


# ------------------------------------------------------------------------------------------- eigenvalues
@pytest.mark.parametrize("m", [1, 2, 3, 5, 8, 24, 38, 64, 88, 120])
def test_real_eig_matches_lapack(B, rng, m):
    """gecon_real_eig_* against numpy.linalg.eigvals (dgeev): random, badly scaled, block-triangular and defective matrices."""
    N = 6
    M = rng.standard_normal((N, m, m))
    M[1] *= np.exp(3.0 * rng.standard_normal((m, 1)))           # badly row-scaled (balancing)
    M[2] = np.triu(M[2])                                         # already triangular: real spectrum on the diagonal
    if m >= 4:
        M[3][m // 2 :, : m // 2] = 0.0                           # block triangular: deflation
        M[4] = np.diag(np.ones(m - 1), 1) * 2.0 + np.eye(m) * 0.5  # one Jordan block (defective)
    re, im, st = B.real_eig(M)
    assert (st == 0).all() and re.shape == (N, m)
    for i in range(N):
        ref = np.linalg.eigvals(M[i])
        got = re[i] + 1j * im[i]
        scale = max(1.0, np.abs(ref).max())
        tol = (1e-6 if i == 4 else 1e-9) * scale  # a Jordan block of size m is only determined to eps^(1/m) in theory; ours is exact
        assert _match_spectra(got, ref, tol) <= tol, (m, i)
        assert (np.diff(np.hypot(re[i], im[i])) >= -1e-12 * scale).all()  # ascending modulus
        assert abs(im[i].sum()) <= 1e-9 * scale  # conjugate pairs
    import torch

    re_d, im_d, st_d = B.real_eig(torch.as_tensor(M, device="cuda"))
    assert np.array_equal(re_d.cpu().numpy(), re) and np.array_equal(im_d.cpu().numpy(), im)


def test_real_eig_flags_nonfinite_input(B):
    from geconpy_b200 import _lib as L

    M = np.eye(4)[None].repeat(2, 0)
    M[1, 2, 1] = np.nan
    re, im, st = B.real_eig(M)
    assert st[0] == 0 and st[1] == L.ST_LL_NONFINITE and np.isnan(re[1]).all() and np.allclose(re[0], 1.0)


@pytest.mark.parametrize("name", ["rbc", "full_nk", "nk_complete_more_shocks", "open_rbc"])
def test_bk_eigenvalues_and_table(B, name):
    """compute_bk_eigenvalues(_pt) and the per-eigenvalue DataFrame of check_bk_condition (perturbation.py:412-583): the count of
    |lambda| > 1 from the eigenvalue kernel equals the eigenvalue-free count kernel and the oracle's; the finite eigenvalues
    equal numpy's on the same regularised matrix."""
    from geconpy_b200.model import perturbation as P

    mod = model(name)
    th = draws(mod, 6, seed=3, width=0.03, valid=True)
    A, Bm, C, D = jacobian_batch(mod, th)
    lead = mod.permuted_lead_var_idx
    re, im = P.compute_bk_eigenvalues_pt(A, Bm, C, D, lead)
    ok, n_fwd, nu = P.check_bk_condition_pt(A, Bm, C, D, lead)
    for i in range(len(th)):
        G0_reg, G1 = osol.bk_matrix_pt(A[i], Bm[i], C[i], lead)
        ref = np.linalg.eigvals(np.linalg.solve(G0_reg, G1))
        mod_got = np.hypot(re[i], im[i])
        assert int((mod_got > 1).sum()) == int((np.abs(ref) > 1).sum()) == int(nu[i])
        # finite eigenvalues against the QZ eigenvalues of the PENCIL (G1, G0_reg), which never forms the 1e8-sized entries of
        # M = G0_reg^-1 G1: a backward-stable eigen-solver applied to M itself (dgeev included: ``lapack`` below) is only good to
        # eps * ||M|| * cond ~ 1e-4 there.  compute_bk_eigenvalues_pt runs the eigenvalue kernel on the Cayley transform of the
        # pencil instead, so it must be at least as close to QZ as dgeev-on-M is, and close in absolute terms.  The regularised
        # infinite eigenvalues (~1e8) are excluded from the comparison (they are counted above).
        import scipy.linalg

        qz = scipy.linalg.eigvals(G1, G0_reg)
        got_fin = (re[i] + 1j * im[i])[mod_got < 1e3]
        assert got_fin.size == (np.abs(qz) < 1e3).sum() == (np.abs(ref) < 1e3).sum()
        ours, lapack = _match_spectra(got_fin, qz[np.abs(qz) < 1e3], 0), _match_spectra(ref[np.abs(ref) < 1e3], qz[np.abs(qz) < 1e3], 0)
        assert ours <= 1e-6 and ours <= max(1e-9, lapack), (name, i, ours, lapack)
    df = P.check_bk_condition(A[0], Bm[0], C[0], D[0], verbose=False)
    n_lead_numeric = int((np.abs(C[0]).sum(axis=0) > 1e-8).sum())  # the numpy variant's rule (perturbation.py:441)
    assert list(df.columns) == ["Modulus", "Real", "Imaginary"] and len(df) == mod.n + n_lead_numeric
    assert (np.diff(df["Modulus"].values) >= 0).all()
    assert P.check_bk_condition(A[0], Bm[0], C[0], D[0], verbose=False, return_value="bool") == bool(ok[0])
    re1, im1, n_forward = P.compute_bk_eigenvalues(A[0], Bm[0], C[0], D[0])
    assert n_forward == n_lead_numeric and int((np.hypot(re1, im1) > 1).sum()) - n_forward == int(nu[0]) - len(lead)


# ------------------------------------------------------------------------------------------- Op layer, executed
def test_ops_perform_through_the_shim(B, monkeypatch, rng):
    mods = pt_shim.install(monkeypatch)
    try:
        cr, gs, re_ = mods["cycle_reduction"], mods["gensys"], mods["real_eig"]
        mod = model("full_nk")
        th = draws(mod, 5, seed=4, width=0.02, valid=True)
        A, Bm, C, D = jacobian_batch(mod, th)
        ref = B.cr_solve(A, Bm, C, D, max_iter=1000, tol=1e-9)
        # single matrix and Blockwise-style leading axis
        _node, (T0,) = pt_shim.run(cr.CycleReductionWrapper(), A[0], Bm[0], C[0])
        assert np.array_equal(T0, ref.T[0])
        _node, (Tb,) = pt_shim.run(cr.CycleReductionWrapper(), A, Bm, C)
        assert Tb.shape == A.shape and np.array_equal(Tb, ref.T)
        _node, (T32,) = pt_shim.run(cr.CycleReductionWrapper(), *(x[0].astype(np.float32) for x in (A, Bm, C)))
        assert T32.dtype == np.float32
        # scan twin: T and the step count the reference publishes as n_cycle_steps
        _node, (Ts, n_steps) = pt_shim.run(cr.ScanCycleReduction(max_iter=50, tol=1e-7), A[0], Bm[0], C[0])
        To, no = osol.cycle_reduction_scan(A[0], Bm[0], C[0], 50, 1e-7)
        assert n_steps == no and n_steps.dtype == np.int32 and rel_fro(Ts, To) <= 1e-9
        _node, (_Tsb, nb) = pt_shim.run(cr.ScanCycleReduction(max_iter=50, tol=1e-7), A, Bm, C)
        assert nb.shape == (5,) and nb[0] == no
        # gensys: T + success
        _node, (Tg, okg) = pt_shim.run(gs.GensysWrapper(), A[0], Bm[0], C[0], D[0])
        assert bool(okg) and rel_fro(Tg, ref.T[0]) <= 1e-9
        # eigenvalues
        M = rng.standard_normal((7, 7))
        _node, (er, ei) = pt_shim.run(re_.RealEig(), M)
        assert _match_spectra(er + 1j * ei, np.linalg.eigvals(M), 1e-10) <= 1e-10
        # the adjoint Op = the adjoint kernel
        Tbar = rng.standard_normal(A.shape)
        _node, (Ab, Bb, Cb) = pt_shim.run(cr.PolicyAdjoint(), A, Bm, C, ref.T, Tbar)
        Ab2, Bb2, Cb2, _Db, _st = B.policy_adjoints(A, Bm, C, ref.T, Tbar)
        assert np.array_equal(Ab, Ab2) and np.array_equal(Bb, Bb2) and np.array_equal(Cb, Cb2)
    finally:
        monkeypatch.undo()
        pt_shim.uninstall()


# ------------------------------------------------------------------------------------------- solver kernels
@pytest.mark.parametrize("name", ["rbc", "rbc_extended", "open_rbc", "full_nk", "new_keynesian", "nk_complete_more_shocks"])
def test_warp_and_cta_solver_kernels_agree(B, name):
    """The one-warp-per-draw kernel (with and without the packed column ranges) against the CTA-per-draw kernel and the oracle,
    on a population that contains failing draws: status, iteration counts and Blanchard-Kahn certificates exact, T and R to 1e-9."""
    mod = model(name)
    th = np.vstack([draws(mod, 40, seed=11, width=0.04, valid=True), draws(mod, 24, seed=12, width=0.10, valid=False)])
    A, Bm, C, D = jacobian_batch(mod, th)
    fin = np.array([all(np.isfinite(M[i]).all() for M in (A, Bm, C, D)) for i in range(len(th))])
    A, Bm, C, D = A[fin], Bm[fin], C[fin], D[fin]
    lead = mod.permuted_lead_var_idx
    n_static = int((~mod.var_has_lag & ~mod.var_has_lead).sum())
    n_lag, n_mixed = int((mod.var_has_lag & ~mod.var_has_lead).sum()), int((mod.var_has_lag & mod.var_has_lead).sum())
    rng_ = (n_static, n_static + n_lag + n_mixed, n_static + n_lag, mod.n)
    kw = dict(max_iter=60, tol=1e-8, resid_tol=1e-8, lead_idx=lead)
    os.environ["GECON_CR_KERNEL"] = "cta"
    try:
        cta = B.cr_solve(A, Bm, C, D, **kw)
    finally:
        os.environ["GECON_CR_KERNEL"] = "warp"
    try:
        dense = B.cr_solve(A, Bm, C, D, **kw)
        packed = B.cr_solve(A, Bm, C, D, col_ranges=rng_, **kw)
        sub = np.sort(np.unique(np.concatenate([np.arange(rng_[0], rng_[1]), lead[:2]]))).astype(np.int32)
        packed_sub = B.cr_solve(A, Bm, C, D, col_ranges=rng_, subset=sub, **kw)
    finally:
        del os.environ["GECON_CR_KERNEL"]
    n_bad = 0
    for i in range(len(A)):
        To, conv, it = osol.cycle_reduction_core(A[i], Bm[i], C[i], max_iter=60, tol=1e-8)
        n_bad += not conv
        CERT = 0x800  # GECON_ST_BK_CERTIFIED is an optimisation hint: a borderline power bound may be found by one kernel and
        # left to the exact count by the other (the dense variant bounds powers of the whole T, not of its lag block)
        bk_ok = osol.bk_condition_pt(A[i], Bm[i], C[i], D[i], lead)[0]
        for res in (cta, dense, packed):
            assert bool(res.converged[i]) == conv and res.n_iter[i] == it, (name, i)
            assert (res.status[i] & ~CERT) == (cta.status[i] & ~CERT), (name, i, res.status[i], cta.status[i])
            if res.status[i] & CERT:
                assert bk_ok and res.n_unstable[i] == len(lead)
            else:
                assert res.n_unstable[i] == -1
            if conv and np.isfinite(To).all():
                assert rel_fro(res.T[i], To) <= 1e-9 and rel_fro(res.R[i], osol.selection_matrix(Bm[i], C[i], D[i], To)) <= 1e-9
                assert abs(res.resid[i] - cta.resid[i]) <= 1e-12 + 1e-6 * abs(cta.resid[i])
            elif not conv:
                assert not res.T[i].any()
        assert np.array_equal(np.isnan(dense.norms[i]), np.isnan(cta.norms[i]))
        if np.isfinite(cta.norms[i]).all():
            np.testing.assert_allclose(packed.norms[i], cta.norms[i], rtol=1e-6, atol=1e-300)
        if conv and np.isfinite(To).all():
            assert rel_fro(packed_sub.T[i], To[np.ix_(sub, sub)]) <= 1e-9
    assert np.array_equal(packed_sub.status & ~0x800, packed.status & ~0x800)
    if name in ("full_nk", "new_keynesian"):
        assert n_bad >= 1  # the wide box does produce draws the iteration rejects


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_scan_semantics(B, kernel):
    """gecon_cr_args.scan_semantics against the oracle's restatement of the scan twin (cycle_reduction.py:246-294): T, the step
    count, and -- with too few steps -- a T that is still solved for (no zeroing) while the flag says not converged."""
    mod = model("full_nk")
    th = draws(mod, 12, seed=13, width=0.03, valid=True)
    A, Bm, C, D = jacobian_batch(mod, th)
    os.environ["GECON_CR_KERNEL"] = kernel
    try:
        full = B.cr_solve(A, Bm, C, D, max_iter=50, tol=1e-7, scan_semantics=True)
        short = B.cr_solve(A, Bm, C, D, max_iter=4, tol=1e-7, scan_semantics=True)
    finally:
        del os.environ["GECON_CR_KERNEL"]
    for i in range(len(th)):
        To, no = osol.cycle_reduction_scan(A[i], Bm[i], C[i], 50, 1e-7)
        assert full.n_iter[i] == no and full.converged[i] and rel_fro(full.T[i], To) <= 1e-9
        Ts, ns = osol.cycle_reduction_scan(A[i], Bm[i], C[i], 4, 1e-7)
        assert short.n_iter[i] == ns == 4 and not short.converged[i] and rel_fro(short.T[i], Ts) <= 1e-9


# ------------------------------------------------------------------------------------------- configure options
BACKWARD_SPEC = {
    "name": "backward_var2", "linear": True, "variables": ["x", "y"], "assumptions": {"x": {}, "y": {}}, "shocks": ["e"],
    "free_params": {"rho": 0.8, "a": 0.5, "b": 0.3}, "deterministic_params": {}, "calibrated_params": {}, "hyper_params": {},
    "steady_state": {"x": "0", "y": "0"}, "equations": ["x__t - rho*x__tm1 - e__t", "y__t - a*x__t - b*y__tm1"], "bounds": {},
}  # fmt: skip


def test_backward_direct_pipeline(B):
    """solver="backward_direct" (statespace.py:205-206, backward_looking.py:8-60): T = -B^-1 A, R = -B^-1 D, then the filter."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    cm = CompiledModel(BACKWARD_SPEC)
    ss = BatchedStateSpace(cm).configure(observed_states=["y"], solver="backward_direct")
    rng = np.random.default_rng(5)
    N = 9
    th = cm.theta_vector() * (1.0 + 0.2 * (2.0 * rng.random((N, 3)) - 1.0))
    sig = np.full((N, 1), 0.1)
    Y = 0.3 * rng.standard_normal((50, 1))
    ll, st = ss.loglik(np.hstack([th, sig]), Y)
    assert (st == 0).all()
    for i in range(N):
        rho, a, b = th[i]
        T = np.array([[rho, 0.0], [a * rho, b]])
        R = np.array([[1.0], [a]])
        ref = oss.kalman_loglik(Y, T, R, np.array([[0.01]]), np.array([[0.0, 1.0]]), np.zeros((1, 1)))
        assert abs(ll[i] - ref) <= 1e-7, (i, ll[i], ref)


def test_constant_params_and_scan_solver_pipeline():
    """constant_params drop columns from the parameter vector (statespace.py:741-753); solver="scan_cycle_reduction" gives the
    same likelihood as cycle_reduction on solvable draws and reports the scan's step count."""
    import torch

    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model("full_nk")
    cm = CompiledModel("full_nk")
    observed = mod.spec["observed_default"]
    th = draws(mod, 16, seed=14, width=0.03, valid=True)
    Y = simulate_obs(mod, 60, seed=5, sigma_err=SIGMA_ERR)
    tail = np.hstack([np.full((16, mod.k), SIGMA_SHOCK), np.full((16, len(observed)), SIGMA_ERR)])
    base = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-8, max_iter=50)
    ll0, st0 = base.loglik(np.hstack([th, tail]), Y)
    frozen = ["beta", "delta"]
    th_c = th.copy()
    for p in frozen:
        th_c[:, mod.param_names.index(p)] = mod.theta_vector()[mod.param_names.index(p)]
    ll_ref, _ = base.loglik(np.hstack([th_c, tail]), Y)
    cons = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-8, max_iter=50, constant_params=frozen)
    keep = [j for j, p in enumerate(mod.param_names) if p not in frozen]
    assert cons.param_names[: len(keep)] == [mod.param_names[j] for j in keep]
    ll_c, st_c = cons.loglik(np.hstack([th[:, keep], tail]), Y)
    assert np.array_equal(ll_c, ll_ref)
    _llg, grad, _stg = cons.loglik_and_grad(np.hstack([th[:, keep], tail])[:4], Y)
    assert grad.shape == (4, cons.n_param) and np.isfinite(grad).all()
    scan = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-8, max_iter=50, solver="scan_cycle_reduction")
    n_it = torch.empty(16, dtype=torch.int32, device="cuda")
    ll_s, st_s = scan.loglik_device(torch.as_tensor(np.hstack([th, tail]), device="cuda"), torch.as_tensor(Y, device="cuda"), out_n_iter=n_it)
    ok = st0 == 0
    assert ok.sum() >= 8 and np.abs(ll_s.cpu().numpy()[ok] - ll0[ok]).max() <= 1e-7
    A, Bm, C, _D = jacobian_batch(mod, th)
    for i in np.flatnonzero(ok):
        assert n_it[i].item() == osol.cycle_reduction_scan(A[i], Bm[i], C[i], 50, 1e-8)[1]


def test_gensys_on_a_pencil_from_gensys_setup(B):
    """gensys(g0, g1, c, psi, pi) for the pencils its callers assemble (gensys.py:568-614 -> :398-521): G_1[:n, :n] is the policy
    matrix, the expectational rows are (T T)[lead], eu = [1, 1, 0]."""
    from geconpy_b200.solvers import gensys as gs

    mod = model("full_nk")
    A, Bm, C, D = mod.jacobians(mod.theta_vector(), mode="statespace")
    g0, g1, c, psi, pi = gs._gensys_setup(A, Bm, C, D)
    G_1, const, impact, f_mat, f_wt, y_wt, gev, eu, loose = gs.gensys(g0, g1, c, psi, pi)
    n = mod.n
    To, conv, _ = osol.cycle_reduction_core(A, Bm, C, tol=1e-12)
    Ro = osol.selection_matrix(Bm, C, D, To)
    lead = np.flatnonzero(np.abs(C).sum(axis=0) > 1e-8)
    assert eu == [1, 1, 0] and G_1.shape == g0.shape and not G_1[:, n:].any() and not const.any()
    assert rel_fro(G_1[:n, :n], To) <= 1e-9 and rel_fro(impact[:n], Ro) <= 1e-9
    assert rel_fro(G_1[n:, :n], (To @ To)[lead]) <= 1e-9 and rel_fro(impact[n:], (To @ Ro)[lead]) <= 1e-9
    assert len(gs.gensys(g0, g1, c, psi, pi, return_all_matrices=False)) == 4


def test_prior_solvability_check():
    """perturbation_diagnostics.py:526-579: QMC draws over the prior bounds -> the solvability_check frame, all on the device."""
    from geconpy_b200.model.compiled import CompiledModel
    from geconpy_b200.model.statistics.perturbation_diagnostics import prior_solvability_check, solvability_check

    cm = CompiledModel("rbc")
    out = prior_solvability_check(cm, 256, seed=0, method="sobol")
    assert len(out) == 256 and {"failure_step", "norm_deterministic", "norm_stochastic"} <= set(out.columns)
    assert set(out.columns[:6]) == set(cm.lin.spec["bounds"])
    ok = out["failure_step"].isna() | out["failure_step"].isin([None])
    assert ok.sum() >= 128 and (out.loc[ok, "norm_deterministic"] <= 1e-8).all()
    again = solvability_check(cm, out[list(cm.lin.spec["bounds"])])
    assert again["failure_step"].equals(out["failure_step"]) and np.array_equal(again["norm_stochastic"].values, out["norm_stochastic"].values, equal_nan=True)
    sub = prior_solvability_check(cm, 64, seed=1, param_subset=["beta", "alpha"], method="lhs")
    assert list(sub.columns[:2]) == ["beta", "alpha"]
    with pytest.raises(NotImplementedError, match="prior distributions"):
        prior_solvability_check(cm, 8, method="sobol_ppf")
    with pytest.raises(ValueError, match="param_subset"):
        prior_solvability_check(cm, 8, param_subset=["nope"])


# ------------------------------------------------------------------------------------------- fused entry point
@pytest.mark.parametrize("name", ["rbc", "full_nk", "nk_complete_more_shocks", "nk_rbc_composite"])
def test_compact_jacobian_equals_the_dense_one(name):
    """gecon_model_jacobian_compact writes exactly the structural non-zeros of the dense kernel's A, B, C, D (same expressions, same
    CSE), in the order gecon_model_structure describes; everything outside the structure is zero in the dense matrices."""
    import torch

    from geconpy_b200.model.compiled import CompiledModel

    cm = CompiledModel(name)
    mod = model(name)
    th = draws(mod, 16, seed=21, width=0.05)
    A, Bm, C, D, _xss, st = cm.jacobian(th)
    nnz, table, off, col_ranges, lead = cm.structure()
    assert off[0] == 0 and off[4] == nnz == len(table) and np.array_equal(lead, cm.permuted_lead_var_idx) and col_ranges == cm.col_ranges
    # parameter rows wider than n_theta, read in place
    wide = np.hstack([th, np.full((len(th), 3), 7.0)])
    vals = torch.full((len(th), nnz), float("nan"), dtype=torch.float64, device="cuda")
    st_c = torch.empty(len(th), dtype=torch.int32, device="cuda")
    cm.jacobian_compact_device(torch.as_tensor(wide, device="cuda"), vals, st_c, torch.cuda.current_stream().cuda_stream)
    vals = vals.cpu().numpy()
    assert np.array_equal(st_c.cpu().numpy(), st)
    seen = [np.zeros_like(M[0], dtype=bool) for M in (A, Bm, C, D)]
    for q, M in enumerate((A, Bm, C, D)):
        for e in range(off[q], off[q + 1]):
            r, c = table[e] >> 16, table[e] & 0xFFFF
            assert np.array_equal(vals[:, e], M[:, r, c], equal_nan=True), (name, q, r, c)
            seen[q][r, c] = True
        assert not np.nan_to_num(M[:, ~seen[q]]).any()
    a_cols, c_cols = np.flatnonzero(seen[0].any(0)), np.flatnonzero(seen[2].any(0))
    assert a_cols.min() >= col_ranges[0] and a_cols.max() < col_ranges[1] and c_cols.min() >= col_ranges[2] and c_cols.max() < col_ranges[3]


@pytest.mark.parametrize("name,with_err", [("rbc", False), ("full_nk", True), ("nk_complete_more_shocks", False), ("nk_rbc_composite", True)])
def test_fused_entry_point_equals_the_kernel_by_kernel_pipeline(name, with_err):
    """gecon_model_loglik (compact Jacobian, C-side chunk loop, scales read in place) against the Python-orchestrated pipeline on
    the same draws, failing ones included: identical status words and iteration counts, log-likelihoods equal to the last bit
    (both feed the same numbers to the same kernels) -- and to the oracle within 1e-7 on a subsample."""
    import torch

    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    mod = model(name)
    cm = CompiledModel(name)
    observed = mod.spec["observed_default"]
    meas = observed[:2] if with_err else []
    kw = dict(observed_states=observed, measurement_error=meas, tol=1e-8, max_iter=100, chunk=48)
    fused = BatchedStateSpace(cm).configure(fused=True, **kw)
    plain = BatchedStateSpace(cm).configure(fused=False, **kw)
    assert fused.fused and not plain.fused
    th = np.vstack([draws(mod, 80, seed=22, width=0.03, valid=True), draws(mod, 40, seed=23, width=0.12, valid=False)])
    th[-1, mod.param_names.index("beta")] = 1.05  # a NaN steady state in every model
    Y = simulate_obs(mod, 40, seed=6, sigma_err=SIGMA_ERR if with_err else 0.0)
    full = np.hstack([th, np.full((len(th), mod.k), SIGMA_SHOCK), np.full((len(th), len(meas)), SIGMA_ERR)])
    full_d, Y_d = torch.as_tensor(full, device="cuda"), torch.as_tensor(Y, device="cuda")
    it_f, it_p = (torch.empty(len(th), dtype=torch.int32, device="cuda") for _ in range(2))
    ll_f, st_f = fused.loglik_device(full_d, Y_d, out_n_iter=it_f)
    ll_p, st_p = plain.loglik_device(full_d, Y_d, out_n_iter=it_p)
    ll_f, st_f, ll_p, st_p = (x.cpu().numpy() for x in (ll_f, st_f, ll_p, st_p))
    CERT = 0x800
    assert np.array_equal(st_f & ~CERT, st_p & ~CERT) and np.array_equal(it_f.cpu().numpy(), it_p.cpu().numpy())
    assert np.array_equal(ll_f, ll_p, equal_nan=True)
    assert (st_f != 0).sum() >= 1 and (st_f == 0).sum() >= 40
    err = np.full(len(meas), SIGMA_ERR) if with_err else None
    for i in list(range(0, 80, 16)) + [85, 100]:
        if with_err:  # the error variances belong to the FIRST len(meas) observables (statespace.py:800-808)
            h = np.zeros(len(observed))
            h[: len(meas)] = SIGMA_ERR
            ref = oss.loglik(mod, th[i], Y, observed, np.full(mod.k, SIGMA_SHOCK), h, tol=1e-8, max_iter=100)
        else:
            ref = oss.loglik(mod, th[i], Y, observed, np.full(mod.k, SIGMA_SHOCK), err, tol=1e-8, max_iter=100)
        if ref["ok"] and np.isfinite(ref["ll"]):
            assert st_f[i] == 0 and abs(ll_f[i] - ref["ll"]) <= 1e-7, (name, i, ll_f[i], ref["ll"])
        else:
            assert np.isneginf(ll_f[i]) and st_f[i] != 0
    # the host-array call (H2D, fused pipeline, D2H) and the per-stage timing hook
    ll_h, st_h = fused.loglik(full, Y)
    assert np.array_equal(ll_h, ll_f, equal_nan=True)
    ev = []
    fused.loglik_device(full_d, Y_d, events=ev)
    assert ev and ev[0][0] == "__fused_ms__" and ev[0][1]["kalman_ll"] > 0.0 and ev[0][1]["cr_solve"] > 0.0


def test_fused_is_refused_for_configurations_it_does_not_cover():
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    cm = CompiledModel("rbc")
    with pytest.raises(ValueError, match="plain state space"):
        BatchedStateSpace(cm).configure(observed_states=["Y", "C"], measurement_error=["Y", "C"], ss_obs_intercept=["Y"], fused=True)
    ss = BatchedStateSpace(cm).configure(observed_states=["Y", "C"], measurement_error=["Y", "C"], temporal_aggregation={"Y": "sum"})
    assert not ss.fused
