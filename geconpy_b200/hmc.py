"""Batched Hamiltonian Monte Carlo over many chains at once, on top of the gradient path.

The reference estimates with NUTS through PyMC, one chain per process, each leapfrog step one compiled ``dlogp`` call
(gEconpy README "Estimation"; the solver Ops' ``pullback`` -> ``o1_policy_function_adjoints``).  On the GPU the natural
shape is the other way round: thousands of chains advance in lock-step and every leapfrog step is ONE batched
``BatchedStateSpace.loglik_and_grad_device`` call (seven kernel launches for the whole population).  This module holds the
sampler plumbing only -- momentum draws, leapfrog, accept/reject masks, all in torch on the device; the arithmetic of the
log-posterior and its gradient is the CUDA library's.

The target of chain i is ``phi * loglik(theta_i) + log prior(theta_i)`` with a uniform prior on the box ``[lo, hi]`` (the same
prior as ``geconpy_b200.smc.TemperedSMC``): a trajectory that leaves the box, or lands on a parameter vector the solver
gates (-inf log-likelihood), is rejected.  Step sizes are per dimension, ``step_scale * (hi - lo)``.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch


@dataclass
class HMCStats:
    accept_rate: float
    mean_logp: float
    max_energy_error: float
    n_failed: int


@dataclass
class BatchedHMC:
    """logp_and_grad: callable theta[N, d] -> (logp[N], grad[N, d]) on the device (``-inf`` logp marks a gated draw, whose
    gradient is ignored).  lo, hi: box of the uniform prior ([d] tensors on the device)."""

    logp_and_grad: object
    lo: torch.Tensor
    hi: torch.Tensor
    step_scale: float = 0.01
    n_leapfrog: int = 8
    seed: int = 0
    stats: list = field(default_factory=list)

    def initialise(self, theta: torch.Tensor):
        self.theta = theta.clone()
        self.gen = torch.Generator(device=theta.device).manual_seed(self.seed)
        self.eps = self.step_scale * (self.hi - self.lo)
        self.logp, self.grad = self.logp_and_grad(self.theta)
        return self

    def _inside(self, th):
        return ((th >= self.lo) & (th <= self.hi)).all(dim=1)

    def step(self) -> HMCStats:
        """One HMC transition of every chain: fresh unit-mass momenta, ``n_leapfrog`` leapfrog steps, Metropolis test."""
        th0, lp0, g0 = self.theta, self.logp, self.grad
        r0 = torch.randn(th0.shape, generator=self.gen, dtype=th0.dtype, device=th0.device)
        th, r, g = th0.clone(), r0.clone(), torch.nan_to_num(g0)
        alive = torch.isfinite(lp0)
        lp = lp0
        for _ in range(self.n_leapfrog):
            r = r + 0.5 * self.eps * g
            th = th + self.eps * r
            alive = alive & self._inside(th)
            th_eval = torch.where(alive[:, None], th, th0)  # dead trajectories are evaluated at a harmless point
            lp, g = self.logp_and_grad(th_eval)
            alive = alive & torch.isfinite(lp)
            g = torch.where(alive[:, None], torch.nan_to_num(g), torch.zeros_like(g))
            r = r + 0.5 * self.eps * g
        h0 = -lp0 + 0.5 * (r0 * r0).sum(dim=1)
        h1 = -lp + 0.5 * (r * r).sum(dim=1)
        dh = torch.where(alive, h1 - h0, torch.full_like(h0, float("inf")))
        log_u = torch.log(torch.rand(th0.shape[0], generator=self.gen, dtype=th0.dtype, device=th0.device))
        accept = alive & (log_u < -dh)
        self.last_accept_prob = torch.clamp(torch.exp(-dh), max=1.0)  # per chain; 0 for rejected-by-construction trajectories
        self.theta = torch.where(accept[:, None], th, th0)
        self.logp = torch.where(accept, lp, lp0)
        self.grad = torch.where(accept[:, None], g, g0)
        fin = torch.isfinite(self.logp)
        st = HMCStats(
            accept_rate=float(accept.double().mean()),
            mean_logp=float(self.logp[fin].mean()) if bool(fin.any()) else float("nan"),
            max_energy_error=float(dh[alive].abs().max()) if bool(alive.any()) else float("nan"),
            n_failed=int((~fin).sum()),
        )
        self.stats.append(st)
        return st

    def run(self, n_steps: int):
        for _ in range(n_steps):
            self.step()
        return self.stats

    def warmup(self, n_steps: int, target_accept: float = 0.8, gamma: float = 0.05, t0: float = 10.0, kappa: float = 0.75):
        """Dual-averaging adaptation of the (common) step-size scale towards a mean acceptance probability of
        ``target_accept`` over the population (Hoffman & Gelman 2014, section 3.2, with the population mean in place of a
        single chain's statistic -- thousands of chains make it a low-noise signal).  Ends with the averaged step size."""
        import math

        mu = math.log(10.0 * self.step_scale)
        log_eps_bar, h_bar = 0.0, 0.0
        for m_ in range(1, n_steps + 1):
            self.step()
            alpha = float(self.last_accept_prob.mean())
            h_bar = (1.0 - 1.0 / (m_ + t0)) * h_bar + (target_accept - alpha) / (m_ + t0)
            log_eps = mu - math.sqrt(m_) / gamma * h_bar
            eta = m_ ** (-kappa)
            log_eps_bar = eta * log_eps + (1.0 - eta) * log_eps_bar
            self.step_scale = math.exp(log_eps)
            self.eps = self.step_scale * (self.hi - self.lo)
        self.step_scale = math.exp(log_eps_bar)
        self.eps = self.step_scale * (self.hi - self.lo)
        return self.step_scale


def statespace_target(statespace, Y, fixed_tail, phi: float = 1.0):
    """``logp_and_grad`` for ``BatchedHMC``: ``phi * log-likelihood`` of a configured ``BatchedStateSpace`` as a function of the
    free parameters; ``fixed_tail`` holds the trailing (not sampled) columns of the parameter vector (shock / error sigmas)."""

    def f(theta):
        full = torch.cat([theta, fixed_tail.expand(theta.shape[0], -1)], dim=1)
        ll, grad, _st = statespace.loglik_and_grad_device(full, Y)
        # gated draws stay -inf for every phi (0 * -inf would be NaN and slip through the isfinite tests as "alive = False"
        # only by accident): the support of the target does not depend on the temperature
        lp = torch.where(torch.isfinite(ll), phi * ll, torch.full_like(ll, float("-inf")))
        return lp, phi * grad[:, : theta.shape[1]]

    return f
