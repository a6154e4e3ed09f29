"""Multi-GPU plumbing: draws shard across ranks with no data-path collective; the only exchange is the SMC-stage
all-gather of per-draw log-likelihoods (8 bytes per draw), after which every rank resamples redundantly.

One process per GPU (``torch.distributed``: NCCL over NVLink on the B200 box, gloo in the CPU tests).  The reference's
only batch mechanism is a fork pool over draws (gEconpy/model/statistics/perturbation_diagnostics.py:470-490); the
partition here is the same idea -- contiguous blocks of draws -- with ranks instead of worker processes.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_draws: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of draws owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n_draws), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_loglik(local_ll: torch.Tensor, n_draws: int) -> torch.Tensor:
    """All-gather the per-draw log-likelihoods of every rank's shard into one ``[n_draws]`` tensor in draw order.
    Equal shards use a single ``all_gather_into_tensor``; ragged shards are padded to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_ll
    world = dist.get_world_size()
    sizes = [shard_bounds(n_draws, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    out = torch.empty(world * width, dtype=local_ll.dtype, device=local_ll.device)
    if all(hi - lo == width for lo, hi in sizes):
        dist.all_gather_into_tensor(out, local_ll.contiguous())
        return out
    padded = torch.full((width,), float("-inf"), dtype=local_ll.dtype, device=local_ll.device)
    padded[: local_ll.numel()] = local_ll
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width : r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def systematic_resample(log_weights: torch.Tensor, seed: int) -> torch.Tensor:
    """Ancestor indices by systematic resampling from (unnormalised) log-weights.  Deterministic in ``seed``: every
    rank calls it on the gathered weights and obtains the same ancestors, so no scatter is needed."""
    lw = torch.nan_to_num(log_weights.detach().to(torch.float64).cpu(), nan=float("-inf"))
    n = lw.numel()
    w = torch.exp(lw - torch.max(lw))
    w = w / w.sum()
    gen = torch.Generator().manual_seed(int(seed))
    u0 = torch.rand(1, generator=gen, dtype=torch.float64)
    positions = (u0 + torch.arange(n, dtype=torch.float64)) / n
    cdf = torch.cumsum(w, 0)
    cdf[-1] = 1.0
    return torch.searchsorted(cdf, positions).clamp_(max=n - 1)
