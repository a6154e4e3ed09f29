"""Dump golden (theta, Y) -> logp vectors from the REAL reference estimation graph, wherever it can be imported.

TEST/ORACLE INFRASTRUCTURE.  The Kalman filter of the reference lives in ``pymc_extras`` (third party; call site
``gEconpy/model/statespace.py:1151-1157``), which is not installed in the build container, so the filter stage of
``oracle/statespace.py`` is "parity unpinned" (DESIGN.md section 4).  This script closes that hole the first time it runs
on a machine where ``import gEconpy, pymc, pymc_extras`` works (a driver-side install under ``baseline/_ref``, a
developer's conda environment, ...):

    python tests/golden/make_kalman_goldens.py            # writes tests/golden/ref_kalman_logp.npz (+ .json summary)

For every case it builds the reference's own graph -- ``statespace_from_gcn -> configure(solver="cycle_reduction") ->
build_statespace_graph -> pm.Model.compile_logp`` (``gEconpy/model/build.py:566-713``, ``statespace.py:822-1215``) -- with
``pm.Flat`` placeholders for every parameter, so that the compiled logp IS the data log-likelihood, evaluates it at a
handful of parameter draws, and then checks which of the oracle's option combinations reproduces it:

    mvn_const      "per_obs" (p log 2 pi per step) | "bare" (log 2 pi)          SURVEY.md A.5 item (i)
    mask_intercept False (v_i = -d_i at missing entries) | True (v_i = 0)       ADVICE round 1, kalman.cuh:379

The matching combination is recorded per case in the fixture (``<case>/mvn_const``, ``<case>/mask_intercept``);
``tests/test_kalman_reference_golden.py`` consumes the file when it exists (CPU: oracle vs. logp; GPU: kernels vs. logp)
and ``bench.py --impl reference`` uses the same builder to time the real reference when it is importable.
Nothing here is imported by the product package.
"""

from __future__ import annotations

import json
import sys
import warnings

from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

SIGMA_SHOCK = 0.01
SIGMA_ERR = 1e-3

# case -> (spec name in geconpy_b200/model/specs, GCN path relative to the reference SOURCE tree, configure kwargs, T_obs,
#          fraction of missing entries)
CASES = {
    "rbc": ("rbc", "gEconpy/data/GCN Files/RBC.gcn", dict(observed_states=["Y"]), 100, 0.0),
    "rbc_two_obs_err": ("rbc", "gEconpy/data/GCN Files/RBC.gcn", dict(observed_states=["Y", "C"], measurement_error=["Y", "C"]), 80, 0.0),
    "rbc_missing_intercept": ("rbc", "gEconpy/data/GCN Files/RBC.gcn",
                              dict(observed_states=["Y", "C"], measurement_error=["Y", "C"], ss_obs_intercept=["Y"]), 80, 0.25),
    "full_nk": ("full_nk", "tests/_resources/test_gcns/full_nk.gcn",
                dict(observed_states=["Y", "pi", "r_G"], measurement_error=["Y", "pi", "r_G"]), 200, 0.0),
}
SOLVER = dict(solver="cycle_reduction", tol=1e-8, max_iter=100)


def find_reference():
    """Returns (gEconpy module, source root or None).  Search order: an importable install, baseline/_ref, /root/reference."""
    for extra in (None, ROOT / "baseline" / "_ref", Path("/root/reference")):
        if extra is not None:
            if not Path(extra).exists():
                continue
            sys.path.insert(0, str(extra))
        try:
            import gEconpy  # noqa: F401
            import pymc  # noqa: F401
            import pymc_extras  # noqa: F401

            pkg = Path(gEconpy.__file__).resolve().parent
            src = pkg.parent if (pkg.parent / "tests" / "_resources").exists() else (Path("/root/reference") if Path("/root/reference/tests").exists() else None)
            return gEconpy, src
        except Exception:
            if extra is not None and str(extra) in sys.path:
                sys.path.remove(str(extra))
            for m in [k for k in sys.modules if k.split(".")[0] in ("gEconpy",)]:
                sys.modules.pop(m, None)
    return None, None


def build_reference_logp(gcn_path, cfg, data_frame, missing_fill_value=-9999.0):
    """The reference's own compiled data log-likelihood: returns (f, names) with f(dict name -> value) -> float."""
    import pymc as pm

    from gEconpy.model.build import statespace_from_gcn

    ss_mod = statespace_from_gcn(str(gcn_path), verbose=False)
    ss_mod.configure(**cfg, **SOLVER, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pm.Model(coords=getattr(ss_mod, "coords", None)) as m:
            for name in ss_mod.param_names:
                info = ss_mod.param_info.get(name, {}) if hasattr(ss_mod, "param_info") else {}
                shape = info.get("shape", ()) or ()
                pm.Flat(name, shape=shape)
            ss_mod.build_statespace_graph(data_frame, add_norm_check=False, missing_fill_value=missing_fill_value)
            f = m.compile_logp()
    return f, list(ss_mod.param_names), ss_mod


def oracle_logp(om, theta, Y, cfg, sig, err, mvn_const, mask_intercept):
    from oracle import statespace as oss

    if cfg.get("ss_obs_intercept"):
        r = oss.loglik_augmented(om, theta, Y, cfg["observed_states"], sig, err if len(err) else None, ss_obs_intercept=cfg["ss_obs_intercept"],
                                 tol=SOLVER["tol"], max_iter=SOLVER["max_iter"], mvn_const=mvn_const, mask_intercept=mask_intercept)
        return r["ll_raw"]
    r = oss.loglik(om, theta, Y, cfg["observed_states"], sig, err if len(err) else None, tol=SOLVER["tol"], max_iter=SOLVER["max_iter"],
                   mvn_const=mvn_const)
    return r["ll_raw"]


def main():
    import pandas as pd

    from oracle import statespace as oss
    from oracle.model import OracleModel

    gE, src = find_reference()
    if gE is None:
        print("gEconpy / pymc / pymc_extras are not importable here: nothing written (the fixture stays absent and the "
              "Kalman stage stays 'parity unpinned').")
        return 2
    import pymc_extras

    out, summary = {}, {"pymc_extras": getattr(pymc_extras, "__version__", "?"), "gEconpy": getattr(gE, "__version__", "?"), "cases": {}}
    pkg_root = Path(gE.__file__).resolve().parent.parent
    for case, (spec, rel, cfg, tobs, miss) in CASES.items():
        gcn = next((p for p in ((src / rel) if src else None, pkg_root / rel) if p is not None and p.exists()), None)
        if gcn is None:
            print(f"{case}: {rel} not found, skipped")
            continue
        om = OracleModel(spec)
        th0 = om.theta_vector()
        rng = np.random.default_rng(7)
        thetas = th0 * (1.0 + 0.02 * (2.0 * rng.random((5, th0.size)) - 1.0))
        thetas[0] = th0
        observed = cfg["observed_states"]
        meas = cfg.get("measurement_error", [])
        k, p = om.k, len(observed)
        sig = np.full(k, SIGMA_SHOCK)
        err = np.full(len(meas), SIGMA_ERR)
        r0 = oss.loglik(om, th0, np.zeros((1, p)), observed, sig)
        x = oss.simulate(r0["T"], r0["R"], sig, tobs, seed=0)
        Y = x[:, [om.var_names.index(v) for v in observed]] + SIGMA_ERR * np.random.default_rng(1).standard_normal((tobs, p)) * (len(meas) > 0)
        if cfg.get("ss_obs_intercept"):
            xss = om.steady_state(th0)
            for v in cfg["ss_obs_intercept"]:
                Y[:, observed.index(v)] += np.log(xss[om.var_names.index(v)])
        if miss > 0:
            Y[np.random.default_rng(2).random(Y.shape) < miss] = np.nan
            Y[3] = np.nan
        df = pd.DataFrame(Y, index=pd.date_range("2000-01-01", periods=tobs, freq="QS"), columns=observed)
        f, names, _ = build_reference_logp(gcn, cfg, df)
        logp = []
        for th in thetas:
            point = dict(zip(om.param_names, th))
            point.update({f"sigma_{s}": SIGMA_SHOCK for s in om.shock_names})
            point.update({f"error_sigma_{v}": SIGMA_ERR for v in meas})
            missing = [n for n in names if n not in point]
            if missing:
                raise RuntimeError(f"{case}: the reference graph wants parameters the spec does not have: {missing}")
            logp.append(float(f({n: np.asarray(point[n], dtype=np.float64) for n in names})))
        logp = np.array(logp)
        best = None
        for mc in ("per_obs", "bare"):
            for mi in (False, True):
                mine = np.array([oracle_logp(om, th, Y, cfg, sig, err, mc, mi) for th in thetas])
                dev = float(np.nanmax(np.abs(mine - logp)))
                if best is None or dev < best[0]:
                    best = (dev, mc, mi)
        out[f"{case}/theta"], out[f"{case}/Y"], out[f"{case}/logp"] = thetas, Y, logp
        out[f"{case}/sigma_shock"], out[f"{case}/sigma_err"] = sig, err
        out[f"{case}/config"] = np.array(json.dumps(dict(cfg, spec=spec, **SOLVER)))
        out[f"{case}/mvn_const"], out[f"{case}/mask_intercept"] = np.array(best[1]), np.array(best[2])
        out[f"{case}/oracle_max_abs_dev"] = np.array(best[0])
        summary["cases"][case] = dict(max_abs_dev=best[0], mvn_const=best[1], mask_intercept=bool(best[2]), logp0=float(logp[0]))
        print(f"{case}: oracle reproduces the reference logp to {best[0]:.3e} with mvn_const={best[1]!r}, mask_intercept={best[2]}")
    if out:
        out["meta/pymc_extras_version"] = np.array(summary["pymc_extras"])
        out["meta/jitter"] = np.array(oss.JITTER_DEFAULT)
        np.savez_compressed(HERE / "ref_kalman_logp.npz", **out)
        (HERE / "ref_kalman_logp.json").write_text(json.dumps(summary, indent=1))
        print(f"wrote {HERE / 'ref_kalman_logp.npz'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
