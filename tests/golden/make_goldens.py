"""Generate the numeric golden fixtures under ``tests/golden/`` from the reference's own test resources and code.

TEST/ORACLE INFRASTRUCTURE.  Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_goldens.py

Outputs (small ``.npz`` files, committed):

* ``ref_linearization.npz``  A, B, C, D goldens of the reference for one_block_1_ss / rbc_2_block_ss / full_nk, read
                             from ``tests/_resources/expected_matrices.py`` (the arrays ``test_linearize`` asserts
                             against, tests/model/test_model.py:405-421), in the reference's ORIGINAL equation x
                             variable order as ``Model.linearize_model`` returns them.
* ``ref_dynare_policy.npz``  Dynare ghx / ghu goldens (``tests/_resources/dynare_outputs/*.mat`` through the
                             reference's own loader ``tests/_resources/load_dynare.py``) with their row / column
                             variable names (``test_solve_matches_dynare``, tests/model/test_model.py:532-562).
* ``ref_cycle_reduction.npz``  inputs and outputs of the REFERENCE's ``cycle_reduction_numpy`` and
                             ``solve_policy_function_with_cycle_reduction`` (gEconpy/solvers/cycle_reduction.py:23-114,
                             328-398) executed from ``/root/reference`` on Jacobians of five models: pins the
                             oracle's restatement of the iteration against the real function.
* ``ref_gensys_components.npz``  the MATLAB-derived alpha/beta vectors and expected outputs the reference uses in
                             ``tests/solvers/test_gensys.py:12-70``.
"""

from __future__ import annotations

import importlib.util
import sys
import warnings

from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(ROOT))

import _ref_shim  # noqa: E402

REF = Path(_ref_shim.REFERENCE_ROOT)


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def linearization_goldens():
    em = _load_module("expected_matrices", REF / "tests/_resources/expected_matrices.py")
    out = {}
    for gcn, d in em.expected_linearization_result.items():
        name = gcn.replace(".gcn", "")
        for mat in "ABCD":
            out[f"{name}/{mat}"] = np.asarray(d[mat], dtype=np.float64)
        out[f"{name}/param_names"] = np.array(list(d["param_dict"].keys()))
        out[f"{name}/param_values"] = np.array([float(v) for v in d["param_dict"].values()])
    np.savez_compressed(HERE / "ref_linearization.npz", **out)
    print("ref_linearization.npz:", sorted({k.split("/")[0] for k in out}))


def dynare_goldens():
    sys.path.insert(0, str(REF / "tests" / "_resources"))
    import os

    os.chdir(REF)  # the loader uses paths relative to the reference's root
    ld = _load_module("load_dynare", REF / "tests/_resources/load_dynare.py")
    out = {}
    for name in ("one_block_1_ss", "rbc_2_block_ss", "full_nk", "basic_rbc", "basic_rbc_loglinear"):
        try:
            res = ld.load_dynare_outputs(name)
        except Exception as e:
            print(f"dynare {name}: FAILED {type(e).__name__}: {e}")
            continue
        T, R = res["T"], res["R"]
        out[f"{name}/ghx"] = T.to_numpy(dtype=np.float64)
        out[f"{name}/ghu"] = R.to_numpy(dtype=np.float64)
        out[f"{name}/rows"] = np.array([str(x) for x in T.index])
        out[f"{name}/state_cols"] = np.array([str(x) for x in T.columns])
        out[f"{name}/shock_cols"] = np.array([str(x) for x in R.columns])
    os.chdir(ROOT)
    np.savez_compressed(HERE / "ref_dynare_policy.npz", **out)
    print("ref_dynare_policy.npz:", sorted({k.split("/")[0] for k in out}))


def cycle_reduction_goldens():
    """Run the reference's own numpy cycle reduction on the oracle's Jacobians."""
    _ref_shim.install()
    from gEconpy.solvers.cycle_reduction import cycle_reduction_numpy, solve_policy_function_with_cycle_reduction

    from oracle.model import OracleModel

    out = {}
    for name in ("rbc", "one_block_1_ss", "rbc_2_block_ss", "full_nk", "nk_complete_more_shocks"):
        mod = OracleModel(name)
        th0 = mod.theta_vector()
        rng = np.random.default_rng(11)
        thetas = [th0] + [th0 * (1 + 0.02 * (2 * rng.random(th0.size) - 1)) for _ in range(3)]
        for d, th in enumerate(thetas):
            A, B, C, D = mod.jacobians(th, mode="statespace")
            if not all(np.isfinite(M).all() for M in (A, B, C, D)):
                continue
            X, res, msg, log_norm = cycle_reduction_numpy(A, B, C, max_iter=1000, tol=1e-9)
            if X is None:
                # the reference wraps the failed `None` in a 0-d object array and then crashes on `C @ T`
                # (cycle_reduction.py:381-396); the solved-policy outputs only exist for converged draws
                T = R = None
            else:
                T, R, _msg2, _ = solve_policy_function_with_cycle_reduction(A, B, C, D, max_iter=1000, tol=1e-9, verbose=False)
            key = f"{name}/{d}"
            out[f"{key}/theta"] = th
            for nm, M in zip("ABCD", (A, B, C, D)):
                out[f"{key}/{nm}"] = M
            out[f"{key}/X"] = np.full_like(A, np.nan) if X is None else X
            out[f"{key}/T"] = np.full_like(A, np.nan) if T is None else T
            out[f"{key}/R"] = np.full_like(D, np.nan) if R is None else R
            out[f"{key}/msg"] = np.array(msg)
        # a truncated run exercises the failure tuple of the numpy twin
        A, B, C, D = mod.jacobians(th0, mode="statespace")
        X, res, msg, log_norm = cycle_reduction_numpy(A, B, C, max_iter=3, tol=1e-9)
        out[f"{name}/short/X_is_none"] = np.array(X is None)
        out[f"{name}/short/msg"] = np.array(msg)
        out[f"{name}/short/log_norm"] = np.array(float(log_norm))
    np.savez_compressed(HERE / "ref_cycle_reduction.npz", **out)
    print("ref_cycle_reduction.npz:", len(out), "arrays")


def gensys_component_goldens():
    """alpha/beta test vectors of tests/solvers/test_gensys.py (values typed in the reference test from MATLAB)."""
    src = (REF / "tests/solvers/test_gensys.py").read_text()
    ns = {"np": np}
    # the test module builds its vectors inline; evaluate just the literal assignments we need
    import ast

    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            nm = node.targets[0].id
            if nm in ("A", "B", "a", "b", "alpha", "beta", "div", "n_unstable", "expected"):
                try:
                    val = eval(compile(ast.Expression(node.value), "<gensys-test>", "eval"), ns)  # noqa: S307
                    arr = np.asarray(val)
                    if arr.dtype != object:
                        out.setdefault(nm, arr)
                except Exception:
                    pass
    np.savez_compressed(HERE / "ref_gensys_components.npz", **out)
    print("ref_gensys_components.npz:", {k: v.shape for k, v in out.items()})


def main():
    warnings.simplefilter("ignore")
    linearization_goldens()
    dynare_goldens()
    cycle_reduction_goldens()
    gensys_component_goldens()


if __name__ == "__main__":
    main()
