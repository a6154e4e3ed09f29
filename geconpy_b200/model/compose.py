"""Block-diagonal composition of two model specs with disjoint, renamed symbols.

SURVEY.md section 8(d), config 4b: the reference ships nothing at 40-60 states, so the Smets-Wouters-scale workload is
the composition of ``nk_complete_more_shocks`` (n = 31, k = 9) with ``rbc_extended`` (n = 14, k = 4): n = 45, k = 13,
39 free parameters.  The two economies do not interact -- A, B, C, D are block diagonal up to the solver's
[static | lagged | mixed | forward] permutation -- but the solver, the determinacy check and the filter see one dense
45-variable system, which is what the size-dependent kernels are exercised on.

Every variable, shock and parameter of the second spec gets ``suffix`` appended to its base name (time suffixes
``__tm1 | __t | __tp1 | __ss`` are kept), in every expression string.
"""

from __future__ import annotations

import copy
import re

_IDENT = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")
_TIME = ("__tm1", "__tp1", "__ss", "__t")


def _rename_expr(text: str, names: set, suffix: str) -> str:
    def sub(m):
        tok = m.group(0)
        base, tail = tok, ""
        for t in _TIME:
            if tok.endswith(t):
                base, tail = tok[: -len(t)], t
                break
        return base + suffix + tail if base in names else tok

    return _IDENT.sub(sub, str(text))


def rename_spec(spec: dict, suffix: str) -> dict:
    """Copy of ``spec`` with ``suffix`` appended to every variable, shock and parameter name."""
    names = set(spec["variables"]) | set(spec["shocks"]) | set(spec["free_params"]) | set(spec.get("hyper_params", {}))
    names |= set(spec.get("deterministic_params", {})) | set(spec.get("calibrated_params", {}))
    r = lambda s: s + suffix  # noqa: E731
    out = copy.deepcopy(spec)
    out["variables"] = [r(v) for v in spec["variables"]]
    out["assumptions"] = {r(k): v for k, v in spec.get("assumptions", {}).items()}
    out["shocks"] = [r(s) for s in spec["shocks"]]
    out["free_params"] = {r(k): v for k, v in spec["free_params"].items()}
    out["hyper_params"] = {r(k): v for k, v in spec.get("hyper_params", {}).items()}
    out["deterministic_params"] = {r(k): _rename_expr(v, names, suffix) for k, v in spec.get("deterministic_params", {}).items()}
    out["calibrated_params"] = {r(k): _rename_expr(v, names, suffix) for k, v in spec.get("calibrated_params", {}).items()}
    out["steady_state"] = {r(k): (None if v is None else _rename_expr(v, names, suffix)) for k, v in spec["steady_state"].items()}
    out["equations"] = [_rename_expr(e, names, suffix) for e in spec["equations"]]
    out["bounds"] = {r(k): v for k, v in spec.get("bounds", {}).items()}
    out["observed_default"] = [r(v) for v in spec.get("observed_default", [])]
    return out


def compose_specs(a: dict, b: dict, name: str, suffix: str = "_2") -> dict:
    """Spec of the economy made of ``a`` and (renamed) ``b`` side by side.  Observables default to those of ``a``."""
    b2 = rename_spec(b, suffix)
    clash = (set(a["variables"]) | set(a["shocks"]) | set(a["free_params"])) & (set(b2["variables"]) | set(b2["shocks"]) | set(b2["free_params"]))
    if clash:
        raise ValueError(f"symbols clash after renaming: {sorted(clash)}")
    out = copy.deepcopy(a)
    out["name"] = name
    out["derived_from"] = f"{a.get('derived_from', a['name'])} (+) {b.get('derived_from', b['name'])}"
    out["derived_by"] = "geconpy_b200/model/compose.py (block-diagonal composition, SURVEY.md section 8d config 4b)"
    out["linear"] = bool(a.get("linear")) and bool(b.get("linear"))
    out["variables"] = list(a["variables"]) + b2["variables"]
    out["shocks"] = list(a["shocks"]) + b2["shocks"]
    for key in ("assumptions", "free_params", "hyper_params", "deterministic_params", "calibrated_params", "steady_state", "bounds"):
        out[key] = {**a.get(key, {}), **b2.get(key, {})}
    out["equations"] = list(a["equations"]) + b2["equations"]
    out["observed_default"] = list(a.get("observed_default", []))
    out["analytic_steady_state"] = bool(a.get("analytic_steady_state")) and bool(b.get("analytic_steady_state"))
    return out


if __name__ == "__main__":
    import json
    import sys

    from pathlib import Path

    spec_dir = Path(__file__).resolve().parent / "specs"
    a = json.loads((spec_dir / "nk_complete_more_shocks.json").read_text())
    b = json.loads((spec_dir / "rbc_extended.json").read_text())
    out = compose_specs(a, b, "nk_rbc_composite")
    (spec_dir / "nk_rbc_composite.json").write_text(json.dumps(out, indent=1))
    print("wrote", spec_dir / "nk_rbc_composite.json", len(out["variables"]), "variables,", len(out["shocks"]), "shocks,",
          len(out["free_params"]), "parameters", file=sys.stderr)
