"""Times BatchedStateSpace.loglik_and_grad_device next to loglik_device on one GPU: medium NK, T_obs = 200."""
import json
import sys

from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel  # noqa: E402


def main():
    name, observed, N, Tobs = "full_nk", ["Y", "pi", "r_G"], 32768, 200
    cm = CompiledModel(name)
    ss = BatchedStateSpace(cm).configure(observed_states=observed, measurement_error=observed, tol=1e-8, max_iter=100)
    rng = np.random.default_rng(0)
    th0 = cm.theta_vector()
    th = th0 * (1.0 + 0.02 * (2.0 * rng.random((N, th0.size)) - 1.0))
    full = np.hstack([th, np.full((N, cm.k), 0.01), np.full((N, 3), 1e-3)])
    Y = 0.01 * rng.standard_normal((Tobs, 3))
    thd, Yd = torch.as_tensor(full, device="cuda"), torch.as_tensor(Y, device="cuda")
    out = {}
    for label, fn in (("loglik", lambda ev: ss.loglik_device(thd, Yd, events=ev)), ("loglik_and_grad", lambda ev: ss.loglik_and_grad_device(thd, Yd, events=ev))):
        for _ in range(2):
            fn(None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev = []
        e0.record()
        res = fn(ev)
        e1.record()
        torch.cuda.synchronize()
        per = {}
        for nm, a, b in ev:
            if nm == "__fused_ms__":  # the fused entry point times its own kernels
                for kn, ms in a.items():
                    per[kn] = per.get(kn, 0.0) + ms
            else:
                per[nm] = per.get(nm, 0.0) + a.elapsed_time(b)
        ms = e0.elapsed_time(e1)
        out[label] = dict(ms=ms, evals_per_s=N / ms * 1e3, kernel_ms=per, ok=int((res[-1] == 0).sum().item()), N=N)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
