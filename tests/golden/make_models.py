"""Generate the model-spec fixtures ``tests/golden/models/*.json`` from the reference's GCN files.

TEST/ORACLE INFRASTRUCTURE.  Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_models.py

For every GCN file listed in ``MODELS`` the *reference's own front-end* (parser, FOC derivation,
``simplify_tryreduce``/``simplify_constants``, steady-state propagation -- i.e. the reference function
``gEconpy/model/build.py:331-463 _compile_gcn``, imported through ``_ref_shim``) is executed and its
output -- variables, shocks, first-order conditions, free/deterministic parameters, analytic steady state
-- is written as a small JSON document in this repo's own model-spec format (see
``geconpy_b200/model_spec.py``).  The GCN parser and the symbolic model derivation are out of scope of
this repo; the specs are the *input* of the hot path, exactly what
``statespace_from_gcn`` hands to ``linearize_model`` (``gEconpy/model/build.py:681-687``).

Time-indexed symbols are written ``<name>__tm1`` / ``<name>__t`` / ``<name>__tp1`` / ``<name>__ss``.
"""

from __future__ import annotations

import json
import re
import sys
import warnings

from pathlib import Path

import sympy as sp

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import _ref_shim  # noqa: E402

REF = Path(_ref_shim.REFERENCE_ROOT)

MODELS = {
    # name: (gcn path relative to the reference root, observed states for the benchmark configs)
    "rbc": ("gEconpy/data/GCN Files/RBC.gcn", ["Y"]),
    "rbc_extended": ("gEconpy/data/GCN Files/RBC_extended.gcn", ["Y", "C", "I"]),
    "new_keynesian": ("gEconpy/data/GCN Files/New_Keynesian.gcn", ["Y", "pi", "r_G"]),
    "nk_complete_more_shocks": (
        "gEconpy/data/GCN Files/sims_2024/nk_complete_more_shocks.gcn",
        ["Y", "C", "I", "N", "pi", "i", "w"],
    ),
    "one_block_1_ss": ("tests/_resources/test_gcns/one_block_1_ss.gcn", ["Y"]),
    "rbc_2_block_ss": ("tests/_resources/test_gcns/rbc_2_block_ss.gcn", ["Y"]),
    "full_nk": ("tests/_resources/test_gcns/full_nk.gcn", ["Y", "pi", "r_G"]),
    "rbc_linearized": ("tests/_resources/test_gcns/rbc_linearized.gcn", ["Y"]),
    "open_rbc": ("tests/_resources/test_gcns/open_rbc.gcn", ["Y"]),
    "pert_fails": ("tests/_resources/test_gcns/pert_fails.gcn", ["Y"]),
    "basic_rbc": ("tests/_resources/test_gcns/basic_rbc.gcn", ["Y"]),
}

_SUFFIX = {-1: "__tm1", 0: "__t", 1: "__tp1", "ss": "__ss"}


def _plain(expr, TimeAwareSymbol):
    """Replace every TimeAwareSymbol by a plain sympy Symbol carrying the time index in its name."""
    expr = sp.sympify(expr)
    repl = {}
    for atom in expr.atoms(TimeAwareSymbol):
        if atom.time_index not in _SUFFIX:
            raise ValueError(f"time index {atom.time_index} of {atom} is outside t-1..t+1")
        repl[atom] = sp.Symbol(atom.base_name + _SUFFIX[atom.time_index])
    return expr.xreplace(repl)


def _bounds_from_gcn_text(text: str) -> dict[str, list[float]]:
    out = {}
    pat = re.compile(r"^\s*([A-Za-z_][A-Za-z_0-9]*)\s*~[^;]*?lower\s*=\s*([-+0-9.eE]+)\s*,\s*upper\s*=\s*([-+0-9.eE]+)", re.M)
    for m in pat.finditer(text):
        out[m.group(1)] = [float(m.group(2)), float(m.group(3))]
    return out


def export(name: str, rel_path: str, observed: list[str]) -> dict:
    from gEconpy.classes.time_aware_symbol import TimeAwareSymbol
    from gEconpy.model.build import _compile_gcn

    gcn_path = REF / rel_path
    objects, dictionaries, _priors, options = _compile_gcn(gcn_path, verbose=False, on_unused_parameters="ignore")
    variables, shocks, equations, _ss_rel, _ss_eqs, ss_solution_dict = objects
    param_dict, hyper_param_dict, deterministic_dict, calib_dict = dictionaries

    ss_sympy = ss_solution_dict.to_sympy() if ss_solution_dict else {}
    ss_by_name = {k.base_name: v for k, v in ss_sympy.items()}
    det_sympy = deterministic_dict.to_sympy() if deterministic_dict else {}
    calib_sympy = calib_dict.to_sympy() if calib_dict else {}

    spec = {
        "name": name,
        "derived_from": rel_path,
        "derived_by": "tests/golden/make_models.py (reference front-end _compile_gcn, gEconpy/model/build.py:331-463)",
        "linear": bool(options.get("linear", False)) if options else False,
        "variables": [v.base_name for v in variables],
        "assumptions": {
            v.base_name: {k: bool(val) for k, val in v.assumptions0.items() if k in ("positive", "negative")}
            for v in variables
        },
        "shocks": [s.base_name for s in shocks],
        "free_params": {str(k): float(v) for k, v in param_dict.to_string().items()},
        "hyper_params": {str(k): float(v) for k, v in (hyper_param_dict.to_string().items() if hyper_param_dict else [])},
        "deterministic_params": {str(k): str(_plain(v, TimeAwareSymbol)) for k, v in det_sympy.items()},
        "calibrated_params": {str(k): str(_plain(v, TimeAwareSymbol)) for k, v in calib_sympy.items()},
        "steady_state": {
            v.base_name: (str(_plain(ss_by_name[v.base_name], TimeAwareSymbol)) if v.base_name in ss_by_name else None)
            for v in variables
        },
        "equations": [str(_plain(eq, TimeAwareSymbol)) for eq in equations],
        "bounds": _bounds_from_gcn_text(gcn_path.read_text()),
        "observed_default": observed,
    }
    spec["analytic_steady_state"] = all(v is not None for v in spec["steady_state"].values())
    return spec


def main():
    warnings.simplefilter("ignore")
    _ref_shim.install()
    out_dir = HERE / "models"
    out_dir.mkdir(exist_ok=True)
    for name, (rel_path, observed) in MODELS.items():
        try:
            spec = export(name, rel_path, observed)
        except Exception as e:  # keep going: some fixtures are deliberately broken models
            print(f"{name}: FAILED {type(e).__name__}: {e}")
            continue
        (out_dir / f"{name}.json").write_text(json.dumps(spec, indent=1) + "\n")
        print(
            f"{name}: n={len(spec['variables'])} k={len(spec['shocks'])} "
            f"params={len(spec['free_params'])} det={len(spec['deterministic_params'])} "
            f"analytic_ss={spec['analytic_steady_state']} linear={spec['linear']}"
        )


if __name__ == "__main__":
    main()
