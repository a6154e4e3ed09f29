"""Tempered SMC sweep over a particle population, sharded across GPUs (BASELINE.json config 5).

Per stage every rank (one process per GPU) mutates its shard with one random-walk Metropolis step -- the step that
costs a batched likelihood evaluation, i.e. the four kernels of ``BatchedStateSpace.loglik_device`` -- reweights by the
tempering increment, all-gathers (log-weight, log-likelihood, status) -- 24 bytes per particle, what the north star names --
with ONE ``all_gather_into_tensor`` (NCCL over NVLink), and then every rank draws the same systematic-resampling ancestors
from the gathered weights (shared seed, computed redundantly on the device) and keeps the slice that is its shard; the
surviving parameter rows that live on other ranks are fetched with one ``all_to_all_single`` (``parallel.fetch_rows``: the
ancestors are sorted, so at most N_local rows arrive, most of them from the rank itself).  No scatter, no host round trip.
The likelihood evaluation is the data path and has no collective in it.

The prior is uniform on a box (``lo``, ``hi``); what the reference would run instead is PyMC's ``sample_smc`` calling the
compiled logp particle by particle.  torch is used for the plumbing (random numbers,
accept/reject masks, prefix sum, searchsorted); the arithmetic of the likelihood is the CUDA library's.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import parallel
from .parallel import systematic_ancestors  # noqa: F401  (re-exported: the resampler of the sweep)


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
    fixed_tail: the trailing columns of the full parameter vector that are not sampled (shock / measurement sigmas).
    ``exchange``: "rows" (default: gather 3 doubles per particle, then fetch only the surviving parameter rows) or "allgather"
    (round 1: all-gather the parameter rows as well; kept for the timing comparison in bench.py)."""

    statespace: object
    lo: torch.Tensor
    hi: torch.Tensor
    fixed_tail: torch.Tensor
    Y: torch.Tensor
    step_scale: float = 0.02
    seed: int = 0
    exchange: str = "rows"
    profile: bool = False  # CUDA events around the exchange (all-gather + row fetch) of every stage -> self.exchange_ms
    stats: list = field(default_factory=list)
    exchange_ms: list = field(default_factory=list)

    def initialise(self, theta_local: torch.Tensor):
        """theta_local: this rank's shard of the initial (prior) population, [N_local, d] on the device."""
        self.rank, self.world = parallel.world()
        self.theta = theta_local.clone()
        self.n_local, self.d = self.theta.shape
        parallel.equal_shards(self.n_local)
        dev = self.theta.device
        self.gen = torch.Generator(device=dev).manual_seed(self.seed * 1000003 + self.rank)
        self.ll = torch.empty(self.n_local, dtype=torch.float64, device=dev)
        self.status = torch.empty(self.n_local, dtype=torch.int32, device=dev)
        self._ll_prop = torch.empty_like(self.ll)
        self._st_prop = torch.empty_like(self.status)
        width = 3 + (self.d if self.exchange == "allgather" else 0)
        self._pack = torch.empty((self.n_local, width), dtype=torch.float64, device=dev)
        self._gath = torch.empty((self.world * self.n_local, width), dtype=torch.float64, device=dev) if self.world > 1 else None
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
        # ---- reweighting and the stage's collective: (log-weight, log-likelihood, status) of every particle, 24 bytes each.
        # (phi' - phi) * ll with ll = -inf is -inf, except at phi' = phi where it is NaN: the resampler maps NaN to -inf
        if self.profile:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
        self._pack[:, 0] = (phi_next - self.phi) * self.ll
        self._pack[:, 1] = self.ll
        self._pack[:, 2] = self.status.to(torch.float64)  # (int32 bit field: exact in a double)
        if self.exchange == "allgather":
            self._pack[:, 3:] = self.theta
        gath = parallel.gather_rows(self._pack, self._gath)
        lw = torch.nan_to_num(gath[:, 0], nan=float("-inf"), neginf=float("-inf"), posinf=float("inf"))
        w = torch.exp(lw - lw.max())
        ess = float(w.sum() ** 2 / (w * w).sum())
        # ---- resampling: same (sorted) ancestors on every rank, each keeps its slice; only surviving rows travel
        anc = parallel.systematic_ancestors(lw, seed=self.seed * 7919 + stage_index)
        mine = anc[self.rank * self.n_local : (self.rank + 1) * self.n_local]
        self.ll = gath[mine, 1].contiguous()
        self.status = gath[mine, 2].to(torch.int32)
        self.theta = gath[mine, 3:].contiguous() if self.exchange == "allgather" else parallel.fetch_rows(self.theta, anc)
        if self.profile:
            ev[1].record()
            ev[1].synchronize()
            self.exchange_ms.append(ev[0].elapsed_time(ev[1]))
        self.phi = float(phi_next)
        fin = torch.isfinite(self.ll)
        st = SMCStageStats(phi=self.phi, ess=ess, accept_rate=float(accept.double().mean()),
                           mean_ll=float(self.ll[fin].mean()) if bool(fin.any()) else float("nan"), n_failed=int((~fin).sum()))
        self.stats.append(st)
        return st

    def run(self, n_stages: int = 10):
        """phi_s = (s / n_stages)^2 tempering schedule (finer steps early, as adaptive schedules end up with)."""
        for s in range(1, n_stages + 1):
            self.stage((s / n_stages) ** 2, s)
        return self.stats
