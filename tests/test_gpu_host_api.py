"""GPU tests of the reference-facing host API (geconpy_b200.solvers / model.perturbation): same calls as the reference's
own tests, checked against the committed golden fixtures (reference outputs) and the oracle."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from helpers import model, rel_fro
from oracle import solvers as osol

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_cycle_reduction_numpy_matches_the_reference_functions_outputs():
    """GPU ``cycle_reduction_numpy`` / ``solve_policy_function_with_cycle_reduction`` vs the outputs of the REFERENCE's
    functions of the same name, recorded from /root/reference by tests/golden/make_goldens.py."""
    from geconpy_b200.solvers.cycle_reduction import cycle_reduction_numpy, solve_policy_function_with_cycle_reduction

    g = np.load(GOLD / "ref_cycle_reduction.npz")
    keys = sorted({"/".join(k.split("/")[:2]) for k in g.files if k.split("/")[1].isdigit()})
    n_ok = 0
    for key in keys:
        A, B, C, D = (g[f"{key}/{m}"] for m in "ABCD")
        X, res, msg, log_norm = cycle_reduction_numpy(A, B, C, max_iter=1000, tol=1e-9)
        T, R, msg2, _ = solve_policy_function_with_cycle_reduction(A, B, C, D, max_iter=1000, tol=1e-9, verbose=False)
        if np.isnan(g[f"{key}/X"]).all():
            assert X is None and res is None and T is None and R is None and msg != "Optimization successful"
            continue
        n_ok += 1
        assert msg == msg2 == str(g[f"{key}/msg"]) == "Optimization successful"
        assert rel_fro(X, g[f"{key}/X"]) <= 1e-9 and rel_fro(T, g[f"{key}/T"]) <= 1e-9 and rel_fro(R, g[f"{key}/R"]) <= 1e-9
        assert np.abs(res).max() < 1e-8 and T.flags["C_CONTIGUOUS"]
    assert n_ok >= 12


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_cycle_reduction_numpy_failure_tuple(name):
    from geconpy_b200.solvers.cycle_reduction import cycle_reduction_numpy

    g = np.load(GOLD / "ref_cycle_reduction.npz")
    mod = model(name)
    A, B, C, D = mod.jacobians(mod.theta_vector())
    X, res, msg, log_norm = cycle_reduction_numpy(A, B, C, max_iter=3, tol=1e-9)
    assert X is None and res is None
    assert msg == str(g[f"{name}/short/msg"])
    assert abs(log_norm - float(g[f"{name}/short/log_norm"])) < 1e-6


@pytest.mark.parametrize("name", ["one_block_1_ss", "rbc_2_block_ss", "full_nk"])
def test_gpu_policy_matches_dynare(name):
    """tests/model/test_model.py:532-562 with the GPU solver in place of gensys: atol = rtol = 1e-5."""
    from geconpy_b200.model.compiled import CompiledModel
    from geconpy_b200.solvers.gensys import solve_policy_function_with_gensys

    g = np.load(GOLD / "ref_dynare_policy.npz")
    cm = CompiledModel(name, log_linearize=False)
    A, B, C, D, _xss, st = cm.jacobian(cm.theta_vector())
    assert st[0] == 0
    G_1, constant, impact, f_mat, f_wt, y_wt, gev, eu, loose = solve_policy_function_with_gensys(A[0], B[0], C[0], D[0], tol=1e-8)
    assert eu[:2] == [1, 1]
    n = cm.n
    T, R = G_1[:n, :n], impact[:n, :]  # the slicing Model._solve_with_gensys applies (model.py:1700-1705)
    inv = cm.inv_var_order
    T, R = T[inv][:, inv], R[inv]
    rows = [cm.var_names.index(v) for v in g[f"{name}/rows"]]
    cols = [cm.var_names.index(v) for v in g[f"{name}/state_cols"]]
    np.testing.assert_allclose(T[rows][:, cols], g[f"{name}/ghx"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(R[rows], g[f"{name}/ghu"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["one_block_1_ss", "rbc_2_block_ss", "full_nk"])
def test_gpu_jacobians_match_reference_goldens(name):
    """tests/model/test_model.py:405-421 (test_linearize) through the generated kernel: atol 1e-8.  The goldens use the
    Model path's static log-linearisation flags: variables with a non-positive steady state stay in levels."""
    from geconpy_b200.model.compiled import CompiledModel

    g = np.load(GOLD / "ref_linearization.npz")
    mod = model(name)
    th = mod.theta_vector()
    xss = mod.steady_state(th)
    not_loglin = [v for v, x in zip(mod.var_names, xss) if not x > 1e-8]
    cm = CompiledModel(name, not_loglin_variables=not_loglin)
    A, B, C, D, _x, st = cm.jacobian(th)
    assert st[0] == 0
    inv_e, inv_v = np.argsort(cm.eq_order), cm.inv_var_order
    for nm, M in zip("ABC", (A[0], B[0], C[0])):
        np.testing.assert_allclose(M[inv_e][:, inv_v], g[f"{name}/{nm}"], atol=1e-8, err_msg=f"{name} {nm}")
    np.testing.assert_allclose(D[0][inv_e], g[f"{name}/D"], atol=1e-8)


def test_gensys_failure_codes_on_pert_fails():
    """tests/model/test_model.py:501-529: eu = [1, 0, 2], message text, (None, None)."""
    from geconpy_b200.solvers.gensys import interpret_gensys_output, solve_policy_function_with_gensys

    mod = model("pert_fails")
    A, B, C, D = mod.jacobians(mod.theta_vector(), mode="model")
    out = solve_policy_function_with_gensys(A, B, C, D, tol=1e-8)
    G_1, impact, eu = out[0], out[2], out[7]
    assert G_1 is None and impact is None
    assert list(eu) == [1, 0, 2]
    assert interpret_gensys_output(eu).endswith("Solution exists, but is not unique.")


def test_gensys_batched_agrees_with_cycle_reduction_and_flags():
    from geconpy_b200.solvers.gensys import gensys_batched

    mod = model("full_nk")
    th = np.tile(mod.theta_vector(), (3, 1))
    th[1, mod.param_names.index("rho_technology")] = 1.08  # unit-root shock: BK violated
    mats = [mod.jacobians(t) for t in th]
    A, B, C, D = (np.stack([m[i] for m in mats]) for i in range(4))
    T, R, ok = gensys_batched(A, B, C, D, lead_idx=mod.permuted_lead_var_idx)
    assert list(ok) == [True, False, True]
    Tg, Rg, success, _ = osol.gensys_policy(A[0], B[0], C[0], D[0])
    assert success and rel_fro(T[0], Tg) <= 1e-8 and rel_fro(R[0], Rg) <= 1e-8  # tests/model/test_perturbation.py:205-206


def test_backward_looking_api(rng):
    """tests/solvers/test_backward_looking.py:38-63: A + B T = 0, B R + D = 0."""
    from geconpy_b200.solvers import backward_looking as bl

    n, k = 6, 2
    A = rng.standard_normal((n, n)) * 0.3
    B = rng.standard_normal((n, n)) + 2 * np.eye(n)
    D = rng.standard_normal((n, k))
    T = bl.solve_backward_policy(A, B)
    R = bl.solve_backward_shock_matrix(B, D)
    assert np.abs(A + B @ T).max() < 1e-12 and np.abs(B @ R + D).max() < 1e-12
    T2, R2 = bl.solve_policy_function_with_backward_direct(A, B, np.zeros((n, n)), D)
    assert np.abs(T2 - T).max() < 1e-12 and np.abs(R2 - R).max() < 1e-12
    with pytest.raises(ValueError):
        bl.solve_policy_function_with_backward_direct(A, B, np.eye(n), D)


def test_check_bk_condition_api():
    """tests/model/test_model.py:593-651: n_forward == n_unstable on rbc_linearized; bool / dataframe / raise."""
    from geconpy_b200.model.perturbation import check_bk_condition, check_bk_condition_pt
    from geconpy_b200.pytensorf.real_eig import count_outside_unit_circle

    mod = model("rbc_linearized")
    A, B, C, D = mod.jacobians(mod.theta_vector())
    assert check_bk_condition(A, B, C, D, verbose=False, return_value="bool") is True
    df = check_bk_condition(A, B, C, D, verbose=False)  # the reference's per-eigenvalue table (perturbation.py:576-582)
    assert list(df.columns) == ["Modulus", "Real", "Imaginary"] and len(df) == mod.n + len(mod.permuted_lead_var_idx)
    assert int((df["Modulus"] > 1).sum()) == len(mod.permuted_lead_var_idx)
    ok, n_fwd, n_unst = check_bk_condition_pt(A, B, C, D, mod.permuted_lead_var_idx)
    assert bool(ok) and n_fwd == int(n_unst)
    bad = model("pert_fails")
    Ab, Bb, Cb, Db = bad.jacobians(bad.theta_vector(), mode="model")
    assert check_bk_condition(Ab, Bb, Cb, Db, verbose=False, return_value="bool") is False
    with pytest.raises(ValueError):
        check_bk_condition(Ab, Bb, Cb, Db, verbose=False, on_failure="raise")
    M = np.diag([0.5, 1.5, -2.0, 0.99, 0.0])
    M[0, 1] = 3.0
    assert int(count_outside_unit_circle(M)) == 2


@pytest.mark.parametrize("name", ["rbc", "full_nk"])
def test_solvability_check_matches_the_per_draw_pipeline(name):
    """the reference's batch-over-draws driver (perturbation_diagnostics.py:362-450) as one GPU pass,
    against the oracle's restatement of _check_one_draw (same labels, norms to 1e-9 relative + 1e-14)."""
    import pandas as pd

    from helpers import draws
    from geconpy_b200.model.compiled import CompiledModel
    from geconpy_b200.model.statistics import solvability_check

    mod = model(name)
    th = np.vstack([draws(mod, 20, seed=51, width=0.05, valid=True), draws(mod, 12, seed=52, width=0.08, valid=False)])
    cols = mod.param_names[:4]
    base = mod.theta_vector()
    samples = pd.DataFrame({c: th[:, mod.param_names.index(c)] for c in cols})
    out = solvability_check(CompiledModel(name), samples, tol=1e-8, max_iter=100, norm_tol=1e-8)
    assert list(out.columns) == cols + ["failure_step", "norm_deterministic", "norm_stochastic"] and len(out) == len(samples)
    seen = set()
    for i in range(len(samples)):
        t = base.copy()
        for c in cols:
            t[mod.param_names.index(c)] = samples[c][i]
        step, nd, ns = osol.solvability_one(mod, t, tol=1e-8, max_iter=100, norm_tol=1e-8)
        seen.add(step)
        got = out["failure_step"][i]
        assert (got is None and step is None) or got == step, (name, i, got, step)
        if np.isnan(nd):
            assert np.isnan(out["norm_deterministic"][i]) and np.isnan(out["norm_stochastic"][i])
        else:
            # converged draws leave rounding noise (1e-15..1e-13) unless the 1e-8 truncation of T, R bites
            assert abs(out["norm_deterministic"][i] - nd) <= 1e-12 + 1e-6 * nd
            assert abs(out["norm_stochastic"][i] - ns) <= 1e-12 + 1e-6 * ns
    assert None in seen
