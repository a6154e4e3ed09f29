"""Development timing of the individual kernels on replicated oracle Jacobians (not the bench)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from helpers import *
from geconpy_b200 import batched as B

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
for name, N, Tobs in [("rbc", 65536, 200), ("full_nk", 32768, 200), ("nk_complete_more_shocks", 16384, 200)]:
    mod = model(name)
    th = draws(mod, 64, seed=1, width=0.02)
    A, Bm, C, D = jacobian_batch(mod, th)
    fin = np.isfinite(A).all(axis=(1, 2)) & np.isfinite(Bm).all(axis=(1, 2)) & np.isfinite(C).all(axis=(1, 2))
    A, Bm, C, D = A[fin], Bm[fin], C[fin], D[fin]
    rep = -(-N // len(A))
    dA, dB, dC, dD = (torch.as_tensor(np.tile(x, (rep, 1, 1))[:N], device="cuda") for x in (A, Bm, C, D))
    Y = torch.as_tensor(simulate_obs(mod, Tobs, seed=0), device="cuda")
    obs = observed_idx(mod, permuted=True)
    lead = mod.permuted_lead_var_idx.astype(np.int32)
    q = torch.full((mod.k,), SIGMA_SHOCK**2, device="cuda", dtype=torch.float64)
    h = torch.full((len(obs),), SIGMA_ERR**2, device="cuda", dtype=torch.float64)
    res = B.cr_solve(dA, dB, dC, dD, tol=1e-9, resid_tol=1e-8)
    t_cr = timeit(lambda: B.cr_solve(dA, dB, dC, dD, tol=1e-9, resid_tol=1e-8))
    t_bk = timeit(lambda: B.bk_count(dA, dB, dC, lead))
    t_kf = timeit(lambda: B.kalman_loglik(res.T, res.R, q, Y, obs_idx=obs, hdiag=h))
    ll, st = B.kalman_loglik(res.T, res.R, q, Y, obs_idx=obs, hdiag=h)
    info = {w: B.kernel_info(w, mod.n if w != "bk_count" else mod.n + len(lead), len(obs), Tobs) for w in ("cr_solve", "kalman_ll", "bk_count")}
    out[name] = dict(N=N, n=mod.n, ms_cr=t_cr, ms_bk=t_bk, ms_kf=t_kf, evals_per_s=N / (t_cr + t_bk + t_kf) * 1e3,
                     mean_iter=float(res.n_iter.double().mean()), ll0=float(ll[0]), bad=int((st != 0).sum()), info=info)
    print(name, json.dumps(out[name]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/quick_time.json").write_text(json.dumps(out, indent=1))
