"""Golden vectors for the posterior-batched moments / simulation row.

TEST/ORACLE INFRASTRUCTURE.  Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_moment_goldens.py

Executes the REFERENCE's own functions -- ``stationary_covariance_matrix``, ``_compute_autocovariance_matrix``
(gEconpy/model/statistics/covariance.py), ``impulse_response_function``, ``simulate`` (gEconpy/model/simulate.py) --
imported from ``/root/reference`` through ``_ref_shim`` (xarray is absent here: the DataArray wrapper is replaced by
the identity, the numbers are the reference's) on policy matrices produced by the oracle, and writes
``tests/golden/ref_moments.npz``.
"""

from __future__ import annotations

import sys

from pathlib import Path
from types import SimpleNamespace

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import _ref_shim  # noqa: E402


def main():
    _ref_shim.install()
    from gEconpy.model import simulate as rs
    from gEconpy.model.statistics import covariance as rc

    rs._irf_to_xarray = lambda data, variable_names, shock_names: data
    rs.xr = SimpleNamespace(DataArray=lambda data, **kw: data)
    rc.xr = SimpleNamespace(DataArray=lambda data, **kw: data)

    from helpers import draws, model
    from oracle import solvers as osol

    out = {}
    for name in ("rbc", "full_nk"):
        mod = model(name)
        fake = SimpleNamespace(variables=[SimpleNamespace(base_name=v) for v in mod.var_names],
                               shocks=[SimpleNamespace(base_name=s) for s in mod.shock_names])
        th = draws(mod, 3, seed=17, width=0.02, valid=True)
        k = mod.k
        rng = np.random.default_rng(3)
        Lq = np.tril(rng.standard_normal((k, k))) * 0.01 + 0.02 * np.eye(k)
        Qfull = Lq @ Lq.T
        traj = rng.standard_normal((12, k)) * 0.01
        sizes = 0.5 + rng.random(k)
        for d in range(len(th)):
            A, B, C, D = mod.jacobians(th[d], mode="statespace")
            T, conv, _ = osol.cycle_reduction_core(A, B, C, max_iter=1000, tol=1e-12)
            assert conv
            R = osol.selection_matrix(B, C, D, T)
            T, R = mod.unpermute_policy(T, R)
            key = f"{name}/{d}"
            out[f"{key}/T"], out[f"{key}/R"] = T, R
            out[f"{key}/Sigma_std"] = rc.stationary_covariance_matrix(fake, T=T, R=R, shock_std=0.01, return_df=False)
            out[f"{key}/Sigma_cov"] = rc.stationary_covariance_matrix(fake, T=T, R=R, shock_cov_matrix=Qfull, return_df=False)
            out[f"{key}/acov"] = rc._compute_autocovariance_matrix(T, out[f"{key}/Sigma_std"], n_lags=6, correlation=False)
            out[f"{key}/acorr"] = rc._compute_autocovariance_matrix(T, out[f"{key}/Sigma_std"], n_lags=6, correlation=True)
            out[f"{key}/irf_unit"] = rs.impulse_response_function(fake, T=T, R=R, simulation_length=25, shock_size=1.0)
            out[f"{key}/irf_sizes_joint"] = rs.impulse_response_function(fake, T=T, R=R, simulation_length=25, shock_size=sizes,
                                                                        return_individual_shocks=False)
            out[f"{key}/irf_dict"] = rs.impulse_response_function(fake, T=T, R=R, simulation_length=10,
                                                                 shock_size={mod.shock_names[-1]: 2.0})
            out[f"{key}/irf_traj"] = rs.impulse_response_function(fake, T=T, R=R, shock_trajectory=traj)
            out[f"{key}/irf_cov"] = rs.impulse_response_function(fake, T=T, R=R, simulation_length=8, shock_cov=Qfull, random_seed=9)
            out[f"{key}/sim"] = rs.simulate(fake, T=T, R=R, n_simulations=3, simulation_length=20, shock_std=0.01, random_seed=5)
        out[f"{name}/Qfull"], out[f"{name}/traj"], out[f"{name}/sizes"] = Qfull, traj, sizes
    np.savez_compressed(HERE / "ref_moments.npz", **out)
    print("ref_moments.npz:", len(out), "arrays", {k: v.shape for k, v in out.items() if k.startswith("full_nk/0/")})


if __name__ == "__main__":
    main()
