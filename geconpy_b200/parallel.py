"""Multi-GPU plumbing: draws shard across ranks with no data-path collective; the only exchanges are those of an SMC stage.

One process per GPU (``torch.distributed``: NCCL over NVLink on the B200 box, gloo in the CPU tests).  The reference's
only batch mechanism is a fork pool over draws (gEconpy/model/statistics/perturbation_diagnostics.py:470-490); the
partition here is the same idea -- contiguous blocks of draws -- with ranks instead of worker processes.

Used by ``geconpy_b200.smc`` (``gather_rows``, ``systematic_ancestors``, ``fetch_rows``, ``equal_shards``) and ``bench.py``
(``gather_loglik``); covered by the world-size-2 gloo tests (``tests/_gloo_worker.py``, ``tests/_gloo_smc_worker.py``).
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    """(rank, world_size); (0, 1) outside ``torch.distributed``."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_draws: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of draws owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(n_draws), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_loglik(local_ll: torch.Tensor, n_draws: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather the per-draw log-likelihoods of every rank's shard into one ``[n_draws]`` tensor in draw order.
    Equal shards use a single ``all_gather_into_tensor`` (into ``out`` when given); ragged shards are padded to the largest."""
    rank, nworld = world()
    if nworld == 1:
        return local_ll
    sizes = [shard_bounds(n_draws, r, nworld) for r in range(nworld)]
    width = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == width for lo, hi in sizes):
        if out is None:
            out = torch.empty(nworld * width, dtype=local_ll.dtype, device=local_ll.device)
        dist.all_gather_into_tensor(out, local_ll.contiguous())
        return out
    out = torch.empty(nworld * width, dtype=local_ll.dtype, device=local_ll.device)
    padded = torch.full((width,), float("-inf"), dtype=local_ll.dtype, device=local_ll.device)
    padded[: local_ll.numel()] = local_ll
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width : r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def equal_shards(n_local: int) -> int:
    """Checks that every rank holds ``n_local`` rows (the packed all-gathers below need equal shards); returns the world size."""
    rank, nworld = world()
    if nworld > 1:
        sizes = [None] * nworld
        dist.all_gather_object(sizes, int(n_local))
        if len(set(sizes)) != 1:
            raise ValueError(f"ranks hold different numbers of particles {sizes}: pad or re-shard to equal shards")
    return nworld


def gather_rows(local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[n_local, c] on every rank -> [world * n_local, c] in rank order: ONE ``all_gather_into_tensor``."""
    rank, nworld = world()
    if nworld == 1:
        return local
    if out is None:
        out = torch.empty((nworld * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def systematic_ancestors(log_weights: torch.Tensor, seed: int) -> torch.Tensor:
    """Systematic resampling on the device the weights live on; deterministic in (weights, seed), so every rank that holds the
    same gathered weights obtains the same ancestors and no scatter is needed.  The ancestors are non-decreasing.  Raises if no
    weight is finite (a population whose every particle was rejected cannot be resampled)."""
    lw = torch.nan_to_num(log_weights.to(torch.float64), nan=float("-inf"), neginf=float("-inf"), posinf=float("inf"))
    n = lw.numel()
    top = lw.max()
    if not bool(torch.isfinite(top)):
        raise RuntimeError("systematic_ancestors: no particle has a finite weight (every log-likelihood is -inf / NaN)")
    w = torch.exp(lw - top)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1].clone()
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    u0 = float(torch.rand(1, generator=gen, dtype=torch.float64))
    positions = (u0 + torch.arange(n, dtype=torch.float64, device=lw.device)) / n
    return torch.searchsorted(cdf, positions).clamp_(max=n - 1)


def fetch_rows(local: torch.Tensor, ancestors: torch.Tensor) -> torch.Tensor:
    """Rows of the GLOBAL population (equal shards, rank order) selected by this rank's slice of ``ancestors``.

    ``ancestors`` is the full, non-decreasing ancestor vector every rank computed (``systematic_ancestors``).  Rank s knows
    which of its rows every other rank needs, so one ``all_to_all_single`` with those split sizes moves exactly the surviving
    rows: at most ``n_local`` rows arrive per rank, most of them from the rank itself -- instead of all-gathering the whole
    population (``world * n_local`` rows per rank)."""
    rank, nworld = world()
    n = local.shape[0]
    if nworld == 1:
        return local.index_select(0, ancestors)
    owner = torch.div(ancestors, n, rounding_mode="floor")
    send_idx, send_counts, recv_counts = [], [], []
    for r in range(nworld):
        sl = slice(r * n, (r + 1) * n)
        mine = owner[sl] == rank                   # rows of MY shard that rank r keeps
        send_idx.append(ancestors[sl][mine] - rank * n)
        send_counts.append(int(mine.sum()))
        recv_counts.append(int((owner[rank * n : (rank + 1) * n] == r).sum()))
    send = local.index_select(0, torch.cat(send_idx)).contiguous()
    recv = torch.empty((n, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts)
    return recv  # sources arrive in rank order and the ancestors are sorted: already in the order of this rank's slice
