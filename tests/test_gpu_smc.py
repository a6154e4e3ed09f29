"""Tempered SMC sweep on one GPU (BASELINE.json config 5; the multi-rank exchange is covered by the gloo test)."""

from __future__ import annotations

import numpy as np
import pytest

from helpers import SIGMA_SHOCK, draws, model, simulate_obs

pytestmark = pytest.mark.gpu


def _sweep(seed, n_particles=2048, n_stages=4):
    import torch

    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel
    from geconpy_b200.smc import TemperedSMC

    mod = model("rbc")
    cm = CompiledModel("rbc")
    ss = BatchedStateSpace(cm).configure(observed_states=["Y"], tol=1e-8, max_iter=100)
    Y = torch.as_tensor(simulate_obs(mod, 80, seed=3), device="cuda")
    th0 = mod.theta_vector()
    lo, hi = th0 * 0.9, th0 * 1.1
    for j, p in enumerate(mod.param_names):
        if p == "beta":
            hi[j] = min(hi[j], 0.999)
        if p.startswith("rho_"):
            hi[j] = min(hi[j], 0.99)
    rng = np.random.default_rng(seed)
    theta = lo + rng.random((n_particles, th0.size)) * (hi - lo)
    smc = TemperedSMC(ss, torch.as_tensor(lo, device="cuda"), torch.as_tensor(hi, device="cuda"),
                      torch.full((1, mod.k), SIGMA_SHOCK, dtype=torch.float64, device="cuda"), Y, step_scale=0.05, seed=seed)
    smc.initialise(torch.as_tensor(theta, device="cuda"))
    ll_prior = float(smc.ll[torch.isfinite(smc.ll)].mean())
    stats = smc.run(n_stages)
    return smc, stats, ll_prior, (lo, hi)


def test_smc_sweep_concentrates_and_is_deterministic():
    import torch

    smc, stats, ll_prior, (lo, hi) = _sweep(seed=5)
    assert len(stats) == 4 and abs(stats[-1].phi - 1.0) < 1e-12
    for st in stats:
        assert 1.0 <= st.ess <= 2048.0 and 0.0 < st.accept_rate < 1.0 and np.isfinite(st.mean_ll)
    # tempering towards the posterior raises the population's mean log-likelihood above the prior population's
    assert stats[-1].mean_ll > ll_prior
    th = smc.theta.cpu().numpy()
    assert (th >= lo - 1e-12).all() and (th <= hi + 1e-12).all()
    # the particles' stored log-likelihoods are the kernel's values at their parameters
    ll_check = torch.empty_like(smc.ll)
    st_check = torch.empty_like(smc.status)
    smc._eval(smc.theta, ll_check, st_check)
    fin = torch.isfinite(smc.ll)
    assert torch.equal(ll_check[fin], smc.ll[fin])
    # same seed, same sweep
    smc2, stats2, _, _ = _sweep(seed=5)
    assert torch.equal(smc2.theta, smc.theta) and [s.ess for s in stats2] == [s.ess for s in stats]


def test_systematic_ancestors_properties():
    import torch

    from geconpy_b200.smc import systematic_ancestors

    lw = torch.log(torch.tensor([0.1, 0.2, 0.0, 0.7], dtype=torch.float64, device="cuda"))
    counts = torch.zeros(4, dtype=torch.long, device="cuda")
    for seed in range(200):
        anc = systematic_ancestors(lw.repeat(64), seed)  # 256 weights
        assert anc.shape == (256,) and bool((anc[1:] >= anc[:-1]).all())
        counts += torch.bincount(anc % 4, minlength=4)
    freq = (counts.double() / counts.sum()).cpu().numpy()
    # (every block of 4 shares the offset u0: 200 independent offsets -> a few per cent of Monte-Carlo error)
    assert freq[2] == 0.0 and np.abs(freq - np.array([0.1, 0.2, 0.0, 0.7])).max() < 3e-2
    assert torch.equal(systematic_ancestors(lw, 3), systematic_ancestors(lw.clone(), 3))
