"""Golden vectors for the state-augmentation bookkeeping, produced by the REFERENCE's own code.

Run in the build container only (needs /root/reference):  python tests/golden/make_augmentation_goldens.py
Executes, through the import shim (pytensor & co. are placeholders; these code paths are pure numpy / pandas):

* ``DSGEStateSpace._make_design_matrix`` on its selector path (gEconpy/model/statespace.py:279-296),
  ``_cumulator_variables``, ``_cumulator_state_names``, ``_n_cumulator_states`` (:556-584) on a stand-in ``self``
  carrying exactly the attributes those methods read;
* ``prepare_mixed_frequency_data`` (:1432-1509).

Output: tests/golden/ref_augmentation.json
"""

import json
import sys
import types

from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import _ref_shim  # noqa: E402

_ref_shim.install()
from gEconpy.model import statespace as ref_ss  # noqa: E402

CASES = [
    dict(states=["A", "C", "K", "Y", "r"], observed=["Y"], ta={"Y": "sum"}, period=4),
    dict(states=["A", "C", "K", "Y", "r"], observed=["Y", "C"], ta={"Y": "mean"}, period=3),
    dict(states=["A", "C", "K", "Y", "r"], observed=["C", "Y", "r"], ta={"Y": "sum", "r": "last", "C": "mean"}, period=4),
    dict(states=["A", "C", "K", "Y", "r"], observed=["r", "K"], ta={"r": "first"}, period=4),
    dict(states=["A", "C", "K", "Y", "r"], observed=["K", "A"], ta={"A": "sum", "K": "sum"}, period=2),
    dict(states=["A", "C", "K", "Y", "r"], observed=["Y"], ta={}, period=4),
]


class _Var:
    def __init__(self, name):
        self.base_name = name


def stand_in(case):
    cls = ref_ss.DSGEStateSpace
    ns = types.SimpleNamespace()
    ns.variables = [_Var(v) for v in case["states"]]
    ns._temporal_aggregation = dict(case["ta"])
    ns._aggregation_period = case["period"]
    ns._obs_equations = {}
    ns._k_orig_states = len(case["states"])
    ns.observed_states = list(case["observed"])
    ns.k_endog = len(case["observed"])
    ns._orig_state_names = cls._orig_state_names.fget(ns)
    ns._cumulator_variables = cls._cumulator_variables.fget(ns)
    ns._n_cumulator_states = cls._n_cumulator_states.fget(ns)
    ns._cumulator_state_names = cls._cumulator_state_names.fget(ns)
    ns.k_states = ns._k_orig_states + ns._n_cumulator_states
    return cls, ns


def main():
    out = {"design": [], "mixed_frequency": []}
    for case in CASES:
        cls, ns = stand_in(case)
        Z = cls._make_design_matrix(ns)
        assert isinstance(Z, np.ndarray)
        out["design"].append(
            dict(case=case, Z=Z.tolist(), cumulator_variables=ns._cumulator_variables, cumulator_state_names=ns._cumulator_state_names,
                 n_cumulator_states=ns._n_cumulator_states)
        )
    annual = pd.DataFrame({"GDP": [100.0, 110.0, 121.0], "R": [0.05, 0.04, 0.03]}, index=pd.to_datetime(["2020", "2021", "2022"]))
    for pos in ("first", "last"):
        for period, freq in ((4, "QS"), (12, "MS")):
            df = ref_ss.prepare_mixed_frequency_data(annual, high_freq=freq, aggregation_period=period, observation_position=pos)
            out["mixed_frequency"].append(
                dict(position=pos, period=period, freq=freq, index=[str(t.date()) for t in df.index],
                     values=[[None if np.isnan(x) else x for x in row] for row in df.to_numpy()])
            )
    (HERE / "ref_augmentation.json").write_text(json.dumps(out, indent=1))
    print("wrote ref_augmentation.json:", len(out["design"]), "design cases,", len(out["mixed_frequency"]), "mixed-frequency cases")


if __name__ == "__main__":
    main()
