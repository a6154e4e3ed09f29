"""Sampler plumbing of geconpy_b200.hmc on the CPU with an analytic target (no CUDA library involved)."""

from __future__ import annotations

import torch

from geconpy_b200.hmc import BatchedHMC


def _gaussian_target(mu, sd):
    def f(th):
        z = (th - mu) / sd
        return -0.5 * (z * z).sum(dim=1), -z / sd

    return f


def test_batched_hmc_samples_a_gaussian_inside_the_box():
    torch.manual_seed(0)
    d, N = 3, 4096
    mu, sd = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64), torch.tensor([0.1, 0.3, 0.2], dtype=torch.float64)
    lo, hi = mu - 8 * sd, mu + 8 * sd
    th0 = lo + (hi - lo) * torch.rand((N, d), dtype=torch.float64)
    hmc = BatchedHMC(_gaussian_target(mu, sd), lo, hi, step_scale=0.02, n_leapfrog=5, seed=1).initialise(th0)
    hmc.run(60)  # trajectory length 5 x 0.32 sd = about a quarter period of the slowest mode (a half period would not mix)
    assert 0.6 < hmc.stats[-1].accept_rate <= 1.0
    assert hmc.stats[-1].max_energy_error < 1.0
    m, s = hmc.theta.mean(0), hmc.theta.std(0)
    assert torch.all((m - mu).abs() < 5 * sd / N**0.5 + 0.05 * sd)
    assert torch.all((s / sd - 1).abs() < 0.08)


def test_leapfrog_is_reversible_and_rejects_outside_the_box_or_gated_points():
    d, N = 2, 64
    mu, sd = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    base = _gaussian_target(mu, sd)

    def gated(th):  # a region the "solver" refuses: logp = -inf there
        lp, g = base(th)
        bad = th[:, 0] > 0.9
        return torch.where(bad, torch.full_like(lp, float("-inf")), lp), g

    lo, hi = -torch.ones(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    th0 = 0.5 * torch.ones((N, d), dtype=torch.float64)
    hmc = BatchedHMC(gated, lo, hi, step_scale=0.2, n_leapfrog=5, seed=3).initialise(th0)
    for _ in range(20):
        hmc.step()
        assert torch.all(hmc.theta >= lo) and torch.all(hmc.theta <= hi)
        assert torch.all(hmc.theta[:, 0] <= 0.9) and torch.isfinite(hmc.logp).all()
    assert any(s.accept_rate < 1.0 for s in hmc.stats)  # some trajectories did leave the box / hit the gated region


def test_dual_averaging_finds_a_step_size_with_the_target_acceptance():
    torch.manual_seed(1)
    d, N = 4, 2048
    mu, sd = torch.zeros(d, dtype=torch.float64), torch.tensor([0.05, 0.1, 0.2, 0.4], dtype=torch.float64)
    lo, hi = mu - 10 * sd, mu + 10 * sd
    th0 = mu + sd * torch.randn((N, d), dtype=torch.float64)
    for start in (1e-4, 0.2):  # far too small and far too large
        hmc = BatchedHMC(_gaussian_target(mu, sd), lo, hi, step_scale=start, n_leapfrog=5, seed=2).initialise(th0)
        eps = hmc.warmup(150, target_accept=0.8)
        assert 1e-3 < eps < 0.2
        hmc.stats.clear()
        hmc.run(20)
        acc = sum(s.accept_rate for s in hmc.stats) / len(hmc.stats)
        assert 0.65 < acc < 0.95, (start, eps, acc)
