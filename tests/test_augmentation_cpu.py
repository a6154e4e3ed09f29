"""State-augmentation bookkeeping: host logic and oracle against golden vectors produced by the
reference's own ``_make_design_matrix`` / ``prepare_mixed_frequency_data`` (tests/golden/make_augmentation_goldens.py)."""

from __future__ import annotations

import json

from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from oracle import statespace as oss

from geconpy_b200.model.augmentation import StateAugmentation, prepare_mixed_frequency_data

GOLD = json.loads((Path(__file__).parent / "golden" / "ref_augmentation.json").read_text())


@pytest.mark.parametrize("g", GOLD["design"], ids=lambda g: "-".join(f"{k}{v}" for k, v in g["case"]["ta"].items()) or "none")
def test_design_matrix_and_names_match_the_reference(g):
    c = g["case"]
    aug = StateAugmentation(c["states"], c["observed"], c["ta"], c["period"])
    Zref = np.array(g["Z"])
    assert aug._cumulator_variables == g["cumulator_variables"]
    assert aug._cumulator_state_names == g["cumulator_state_names"]
    assert aug._n_cumulator_states == g["n_cumulator_states"]
    assert np.array_equal(aug.design_matrix(), Zref)
    # the oracle restatement agrees with the reference too
    assert np.array_equal(oss.design_matrix(c["states"], c["observed"], c["ta"], c["period"]), Zref)


@pytest.mark.parametrize("g", GOLD["design"], ids=lambda g: "-".join(f"{k}{v}" for k, v in g["case"]["ta"].items()) or "none")
def test_augmented_transition_is_a_lag_chain(g):
    """Known answer: with x_t = T x_{t-1}, the augmented state carries x_{t-1}, ..., x_{t-s+1} of each aggregated
    variable, so Z s_t equals the windowed sum / mean of the un-augmented path (statespace.py:598-650)."""
    c = g["case"]
    rng = np.random.default_rng(0)
    k = len(c["states"])
    T = 0.4 * rng.standard_normal((k, k))
    aug = StateAugmentation(c["states"], c["observed"], c["ta"], c["period"])
    Ta = aug.augment_transition(T)
    assert np.array_equal(Ta, oss.augment_transition(T, c["states"], c["ta"], c["period"]))
    assert np.array_equal(aug.augment_selection(np.ones((k, 2)))[k:], np.zeros((aug._n_cumulator_states, 2)))
    x = [rng.standard_normal(k)]
    s = np.concatenate([x[0], np.zeros(aug._n_cumulator_states)])
    Z = aug.design_matrix()
    for t in range(1, 12):
        x.append(T @ x[-1])
        s = Ta @ s
        if t >= c["period"]:
            for i, name in enumerate(c["observed"]):
                j = c["states"].index(name)
                m = c["ta"].get(name)
                if m in ("sum", "mean"):
                    w = sum(x[t - d][j] for d in range(c["period"]))
                    w = w / c["period"] if m == "mean" else w
                else:
                    w = x[t][j]
                assert abs(Z[i] @ s - w) < 1e-12


def test_batched_augment_matches_single():
    aug = StateAugmentation(["a", "b", "c"], ["b"], {"b": "sum"}, 3)
    T = np.random.default_rng(1).standard_normal((5, 3, 3))
    Ta = aug.augment_transition(T)
    assert Ta.shape == (5, 5, 5)
    for i in range(5):
        assert np.array_equal(Ta[i], oss.augment_transition(T[i], ["a", "b", "c"], {"b": "sum"}, 3))


def test_validation_errors():
    with pytest.raises(ValueError, match="not in observed_states"):
        StateAugmentation(["a", "b"], ["a"], {"b": "sum"}, 4)
    with pytest.raises(ValueError, match="aggregation_period"):
        StateAugmentation(["a", "b"], ["a"], {"a": "sum"}, 1)
    with pytest.raises(ValueError, match="Unknown temporal aggregation"):
        StateAugmentation(["a", "b"], ["a"], {"a": "median"}, 4)


@pytest.mark.parametrize("g", GOLD["mixed_frequency"], ids=lambda g: f"{g['position']}-{g['period']}")
def test_prepare_mixed_frequency_data_matches_the_reference(g):
    annual = pd.DataFrame({"GDP": [100.0, 110.0, 121.0], "R": [0.05, 0.04, 0.03]}, index=pd.to_datetime(["2020", "2021", "2022"]))
    df = prepare_mixed_frequency_data(annual, high_freq=g["freq"], aggregation_period=g["period"], observation_position=g["position"])
    assert [str(t.date()) for t in df.index] == g["index"]
    ref = np.array([[np.nan if x is None else x for x in row] for row in g["values"]])
    assert np.array_equal(df.to_numpy(), ref, equal_nan=True)


def test_oracle_intercept_and_sum_rule():
    d = oss.obs_intercept(np.array([2.0, 3.0, 4.0]), ["a", "b", "c"], ["c", "a", "b"], ["a", "c"], {"a"}, {"c": "sum", "a": "mean"}, 4)
    assert np.allclose(d, [16.0, np.log(2.0), 0.0])


# ------------------------------------------------------------------------------------------------ observation equations
OBS_EQS = {"Yobs": "log(Y[])", "dC": "alpha * (log(C[]) - log(C[-1])) + beta", "mix": "Y[] / C[-2] + K[ss] * A[-1] ^ 2"}


def test_observation_equation_linearisation_matches_the_complex_step_oracle(tmp_path):
    """codegen's sympy linearisation + generated kernel text (compiled with g++ -DGECON_HOST_CHECK) against the oracle,
    which differentiates numerically by a complex step: intercepts, coefficients, lag layout, aggregation broadcast."""
    import ctypes as C
    import subprocess

    from helpers import draws, model

    from geconpy_b200.model.codegen import LinearizedModel, load_spec

    root = Path(__file__).resolve().parent.parent
    lin = LinearizedModel(load_spec(root / "geconpy_b200" / "model" / "specs" / "rbc.json"))
    mod = model("rbc")
    observed, ta, period = ["Yobs", "dC", "mix"], {"dC": "sum", "mix": "mean"}, 3
    lin_terms = {k: lin.linearize_observation_equation(*lin.parse_observation_equation(k, e)) for k, e in OBS_EQS.items()}
    depths = StateAugmentation.required_obs_lag_depths({k: t[1].keys() for k, t in lin_terms.items()}, ta, period)
    aug = StateAugmentation(list(mod.var_names), observed, ta, period, obs_equation_names=tuple(OBS_EQS), obs_lag_depths=depths)
    assert aug._cumulator_variables == [] and aug._n_obs_lag_states == sum(depths.values())
    # deepest effective lag: C at lag 2 + (3 - 1) aggregation headroom, Y at lag 0 + 2, A at lag 1 + 2 (statespace.py:1040-1053)
    assert depths == {"C": 4, "Y": 2, "A": 3}
    assert aug._obs_lag_state_names[:4] == ["C_obs_lag1", "C_obs_lag2", "C_obs_lag3", "C_obs_lag4"]
    ns, p = aug.k_states, len(observed)
    z_cells, d_cells = {}, {}
    for i, name in enumerate(observed):
        icpt, coeffs = lin_terms[name]
        for col, terms in aug.design_cells(name, coeffs).items():
            z_cells[i * ns + col] = sum(w * coeffs[key] for w, key in terms)
        d_cells[i] = icpt * (period if ta.get(name) == "sum" else 1)
    src = tmp_path / "obs.cpp"
    src.write_text(lin.obs_source(z_cells, d_cells, "test"))
    so = tmp_path / "obs.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-DGECON_HOST_CHECK", "-o", str(so), str(src)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.gecon_obs_host_check.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    th = draws(mod, 4, seed=7, width=0.05, valid=True)
    N = len(th)
    Z, d = np.zeros((N, p, ns)), np.zeros((N, p))
    lib.gecon_obs_host_check(th.ctypes.data, N, Z.ctypes.data, p * ns, d.ctypes.data, p)
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((6, p))
    for i in range(N):
        ref = oss.loglik_augmented(mod, th[i], Y, observed, [0.01], [0.01] * p, temporal_aggregation=ta, aggregation_period=period,
                                   observation_equations=OBS_EQS)  # fmt: skip
        assert ref["T_aug"].shape == (ns, ns)
        assert np.array_equal(ref["T_aug"][mod.n :], aug.transition_rows())
        np.testing.assert_allclose(Z[i], ref["Z"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(d[i], ref["d"], rtol=1e-12, atol=1e-14)


def test_observation_equation_errors_follow_the_reference():
    from geconpy_b200.model.codegen import LinearizedModel, load_spec

    root = Path(__file__).resolve().parent.parent
    lin = LinearizedModel(load_spec(root / "geconpy_b200" / "model" / "specs" / "rbc.json"))
    with pytest.raises(ValueError, match="lead reference"):
        lin.parse_observation_equation("x", "log(Y[1])")
    with pytest.raises(ValueError, match="unknown model variable"):
        lin.parse_observation_equation("x", "log(Q[])")
    with pytest.raises(ValueError):
        lin.parse_observation_equation("x", "not_a_parameter * Y[]")
    icpt, coeffs = lin.linearize_observation_equation(*lin.parse_observation_equation("x", "log(Y[]) - log(Y[-1])"))
    assert icpt == 0 and coeffs == {("Y", 0): 1, ("Y", -1): -1}


# ------------------------------------------------------------------------------------------------ configure() host logic
@pytest.fixture(scope="module")
def rbc_statespace():
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    return lambda: BatchedStateSpace(CompiledModel("rbc"))  # builds (or finds) the generated Jacobian library; no device needed


def test_configure_validation_follows_the_reference(rbc_statespace):
    """Argument checks of DSGEStateSpace.configure (statespace.py:930-1005), same conditions and wording."""
    ss = rbc_statespace()
    with pytest.raises(ValueError, match="unknown observed states"):
        ss.configure(observed_states=["Q"])
    with pytest.raises(ValueError, match="stochastic singularity"):
        ss.configure(observed_states=["Y", "C"])  # one shock, no measurement error
    with pytest.raises(ValueError, match="measurement error on unobserved"):
        ss.configure(observed_states=["Y"], measurement_error=["C"])
    with pytest.raises(ValueError, match="ss_obs_intercept entries are not in observed_states"):
        ss.configure(observed_states=["Y"], ss_obs_intercept=["C"])
    with pytest.raises(ValueError, match="temporal_aggregation entries are not in observed_states"):
        ss.configure(observed_states=["Y"], temporal_aggregation={"C": "sum"})
    with pytest.raises(ValueError, match="aggregation_period must be >= 2"):
        ss.configure(observed_states=["Y"], temporal_aggregation={"Y": "sum"}, aggregation_period=1)
    with pytest.raises(ValueError, match="observation_equations entries are not in observed_states"):
        ss.configure(observed_states=["Y"], observation_equations={"dY": "log(Y[])"})
    with pytest.raises(ValueError, match="both observation_equations and ss_obs_intercept"):
        ss.configure(observed_states=["dY"], measurement_error=["dY"], observation_equations={"dY": "log(Y[])"}, ss_obs_intercept=["dY"])
    with pytest.raises(NotImplementedError, match="augmented state dimension"):
        ss.configure(observed_states=["Y"], temporal_aggregation={"Y": "sum"}, aggregation_period=80)


def test_configure_layout_and_parameter_names(rbc_statespace):
    ss = rbc_statespace().configure(
        observed_states=["Y", "dC"], measurement_error=["Y", "dC"], temporal_aggregation={"Y": "mean"}, aggregation_period=3,
        ss_obs_intercept=["Y"], observation_equations={"dC": "log(C[]) - log(C[-1])"},
    )  # fmt: skip
    m = ss.model
    # filter states: lagged variables + observed model variables + variables an observation equation refers to
    names = [m.lin.vars_perm[int(u)] for u in ss.filter_vars]
    assert {"Y", "C"} <= set(names) and ss.n_filter == len(names) < m.n
    assert ss.aug.augmented_state_names[ss.n_filter :] == ["Y_cumulator_lag1", "Y_cumulator_lag2", "C_obs_lag1"]
    assert ss.n_aug == ss.n_filter + 3 and ss.dense_Z.shape == (2, ss.n_aug)
    assert np.count_nonzero(ss.dense_Z[0]) == 3 and np.allclose(ss.dense_Z[0][ss.dense_Z[0] != 0], 1 / 3)
    assert not ss.dense_Z[1].any()  # the equation's row is written per draw by the generated observation kernel
    assert ss.param_names == list(m.param_names) + ["sigma_epsilon_A" if "epsilon_A" in m.shock_names else f"sigma_{m.shock_names[0]}"] + [
        "error_sigma_Y", "error_sigma_dC"]
    full = rbc_statespace().configure(observed_states=["Y"], full_shock_covariance=True)
    assert full.param_names[len(m.param_names) :] == ["state_cov[0,0]"]
    # reduce_state=False keeps every variable, in solver order
    allv = rbc_statespace().configure(observed_states=["Y"], reduce_state=False)
    assert allv.n_filter == m.n and list(allv.filter_vars) == list(range(m.n)) and allv.filter_t_cols == 0


def test_filter_variables_come_lagged_first_and_t_cols_counts_them():
    """The filter variables are ordered [lagged | observed only]: T = -A1hat^-1 A has non-zero columns only at the lagged variables, so
    ``filter_t_cols`` (gecon_kalman_args.t_cols) is the number of lagged variables whenever something follows them."""
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    cm = CompiledModel("nk_complete_more_shocks")
    lag_lo, lag_hi = cm.col_ranges[0], cm.col_ranges[1]
    ss = BatchedStateSpace(cm).configure(observed_states=["Y", "C", "I", "N", "pi", "i", "w"])
    fv = [int(v) for v in ss.filter_vars]
    t = ss.filter_t_cols
    assert (ss.n_filter, t) == (19, 16) == (len(fv), lag_hi - lag_lo)
    assert fv[:t] == list(range(lag_lo, lag_hi)) and all(not (lag_lo <= v < lag_hi) for v in fv[t:]) and fv[t:] == sorted(fv[t:])
    names = [cm.lin.vars_perm[v] for v in fv]
    assert [names[i] for i in ss.obs_idx_filter] == ["Y", "C", "I", "N", "pi", "i", "w"]
    # every observable a lagged variable: nothing follows the lagged block, T is dense in the filter's eyes
    lagged_names = names[:t]
    only = BatchedStateSpace(cm).configure(observed_states=lagged_names[:2])
    assert only.n_filter == t and only.filter_t_cols == 0
