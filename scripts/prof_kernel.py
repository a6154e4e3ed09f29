"""Run the three kernels once on replicated oracle Jacobians of one model (ncu target; not the bench).
Usage: python scripts/prof_kernel.py <model> <N> [Tobs]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from helpers import *
from geconpy_b200 import batched as B

name, N = sys.argv[1], int(sys.argv[2])
Tobs = int(sys.argv[3]) if len(sys.argv) > 3 else 200
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
for _ in range(2):
    res = B.cr_solve(dA, dB, dC, dD, tol=1e-8, resid_tol=1e-8, lead_idx=lead)
    ll, st = B.kalman_loglik(res.T, res.R, q, Y, obs_idx=obs, hdiag=h)
torch.cuda.synchronize()
print("ok", float(ll[0]), int((st != 0).sum()))
