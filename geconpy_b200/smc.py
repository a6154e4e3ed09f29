"""Tempered SMC sweep over a particle population, sharded across GPUs (BASELINE.json config 5).

Per stage every rank (one process per GPU) mutates its shard with one random-walk Metropolis step -- the step that
costs a batched likelihood evaluation, i.e. the four kernels of ``BatchedStateSpace.loglik_device`` -- reweights by the
tempering increment, exchanges the particles with ONE ``all_gather_into_tensor`` (log-weight, log-likelihood and theta
packed in a single [N_local, 2 + d] buffer; NCCL over NVLink), and then every rank draws the same systematic-resampling
ancestors from the gathered weights (shared seed, computed redundantly on the device), keeping the slice that is its
shard.  No scatter, no host round trip.  The likelihood evaluation is the data path and has no collective in it.

The prior is uniform on a box (``lo``, ``hi``); what the reference would run instead is PyMC's ``sample_smc`` calling the
compiled logp particle by particle.  torch is used for the plumbing (random numbers,
accept/reject masks, prefix sum, searchsorted); the arithmetic of the likelihood is the CUDA library's.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def systematic_ancestors(log_weights: torch.Tensor, seed: int) -> torch.Tensor:
    """Systematic resampling on the device the weights live on; deterministic in (weights, seed), so every rank that
    holds the same gathered weights obtains the same ancestors."""
    lw = torch.nan_to_num(log_weights.to(torch.float64), nan=float("-inf"))
    n = lw.numel()
    w = torch.exp(lw - lw.max())
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1].clone()
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    u0 = float(torch.rand(1, generator=gen, dtype=torch.float64))
    positions = (u0 + torch.arange(n, dtype=torch.float64, device=lw.device)) / n
    return torch.searchsorted(cdf, positions).clamp_(max=n - 1)


@dataclass
class SMCStageStats:
    phi: float
    ess: float
    accept_rate: float
    mean_ll: float
    n_failed: int


@dataclass
class TemperedSMC:
    """statespace: a configured ``BatchedStateSpace``; lo/hi: prior box of the free parameters (device tensors, [d]);
    fixed_tail: the trailing columns of the full parameter vector that are not sampled (shock / measurement sigmas)."""

    statespace: object
    lo: torch.Tensor
    hi: torch.Tensor
    fixed_tail: torch.Tensor
    Y: torch.Tensor
    step_scale: float = 0.02
    seed: int = 0
    stats: list = field(default_factory=list)

    def initialise(self, theta_local: torch.Tensor):
        """theta_local: this rank's shard of the initial (prior) population, [N_local, d] on the device."""
        self.rank, self.world = _world()
        self.theta = theta_local.clone()
        self.n_local, self.d = self.theta.shape
        dev = self.theta.device
        self.gen = torch.Generator(device=dev).manual_seed(self.seed * 1000003 + self.rank)
        self.ll = torch.empty(self.n_local, dtype=torch.float64, device=dev)
        self.status = torch.empty(self.n_local, dtype=torch.int32, device=dev)
        self._ll_prop = torch.empty_like(self.ll)
        self._st_prop = torch.empty_like(self.status)
        self._pack = torch.empty((self.n_local, 2 + self.d), dtype=torch.float64, device=dev)
        self._gath = torch.empty((self.world * self.n_local, 2 + self.d), dtype=torch.float64, device=dev)
        self.phi = 0.0
        self._eval(self.theta, self.ll, self.status)
        return self

    def _eval(self, theta, out_ll, out_status):
        full = torch.cat([theta, self.fixed_tail.expand(theta.shape[0], -1)], dim=1)
        self.statespace.loglik_device(full, self.Y, out_ll=out_ll, out_status=out_status)

    def stage(self, phi_next: float, stage_index: int) -> SMCStageStats:
        """One tempering stage: Metropolis mutation at phi, reweighting to phi_next, exchange, resampling."""
        dev = self.theta.device
        # ---- mutation: theta' = theta + scale (hi - lo) z, one batched likelihood evaluation of the proposals
        z = torch.randn(self.theta.shape, generator=self.gen, dtype=torch.float64, device=dev)
        prop = self.theta + self.step_scale * (self.hi - self.lo) * z
        inside = ((prop >= self.lo) & (prop <= self.hi)).all(dim=1)
        prop = torch.where(inside[:, None], prop, self.theta)  # outside the box: prior density 0, never accepted
        self._eval(prop, self._ll_prop, self._st_prop)
        log_u = torch.log(torch.rand(self.n_local, generator=self.gen, dtype=torch.float64, device=dev))
        accept = inside & (log_u < self.phi * (self._ll_prop - self.ll)) & torch.isfinite(self._ll_prop)
        if self.phi == 0.0:  # prior stage: every in-box proposal with a finite likelihood is a draw from the prior
            accept = inside & torch.isfinite(self._ll_prop)
        self.theta = torch.where(accept[:, None], prop, self.theta)
        self.ll = torch.where(accept, self._ll_prop, self.ll)
        self.status = torch.where(accept, self._st_prop, self.status)
        # ---- reweighting and the stage's only collective
        self._pack[:, 0] = (phi_next - self.phi) * self.ll
        self._pack[:, 1] = self.ll
        self._pack[:, 2:] = self.theta
        if self.world > 1:
            dist.all_gather_into_tensor(self._gath, self._pack)
            gath = self._gath
        else:
            gath = self._pack
        lw = torch.nan_to_num(gath[:, 0], nan=float("-inf"))
        w = torch.exp(lw - lw.max())
        ess = float(w.sum() ** 2 / (w * w).sum())
        # ---- resampling: same ancestors on every rank, each keeps its slice
        anc = systematic_ancestors(lw, seed=self.seed * 7919 + stage_index)
        mine = anc[self.rank * self.n_local : (self.rank + 1) * self.n_local]
        self.ll = gath[mine, 1].contiguous()
        self.theta = gath[mine, 2:].contiguous()
        self.phi = float(phi_next)
        st = SMCStageStats(phi=self.phi, ess=ess, accept_rate=float(accept.double().mean()),
                           mean_ll=float(self.ll[torch.isfinite(self.ll)].mean()), n_failed=int((~torch.isfinite(self.ll)).sum()))
        self.stats.append(st)
        return st

    def run(self, n_stages: int = 10):
        """phi_s = (s / n_stages)^2 tempering schedule (finer steps early, as adaptive schedules end up with)."""
        for s in range(1, n_stages + 1):
            self.stage((s / n_stages) ** 2, s)
        return self.stats
