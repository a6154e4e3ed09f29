"""Worker for test_draw_sharding_and_allgather_world_size_2 (run under torch.distributed.run, gloo, CPU)."""
import sys

from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from geconpy_b200.parallel import fetch_rows, gather_loglik, gather_rows, shard_bounds, systematic_ancestors  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N = 1001  # ragged on purpose
lo, hi = shard_bounds(N, rank, world)
all_bounds = [shard_bounds(N, r, world) for r in range(world)]
assert all_bounds[0][0] == 0 and all_bounds[-1][1] == N and all(a[1] == b[0] for a, b in zip(all_bounds, all_bounds[1:]))
# each rank "evaluates" its shard: ll_i = -i  (stands in for the kernel output; the exchange is what is tested)
local = -torch.arange(lo, hi, dtype=torch.float64)
full = gather_loglik(local, N)
assert full.shape == (N,) and torch.equal(full, -torch.arange(N, dtype=torch.float64)), full[:5]
idx = systematic_ancestors(full, seed=7)
idx2 = systematic_ancestors(full, seed=7)
assert torch.equal(idx, idx2) and idx.shape == (N,) and int(idx.min()) >= 0 and int(idx.max()) < N
# every rank derives the same ancestors from the gathered weights: no scatter needed
gathered = [torch.empty_like(idx) for _ in range(world)]
dist.all_gather(gathered, idx)
assert all(torch.equal(g, idx) for g in gathered)
# equal shards take the single all_gather_into_tensor path
N2 = 64
lo2, hi2 = shard_bounds(N2, rank, world)
full2 = gather_loglik(torch.arange(lo2, hi2, dtype=torch.float64), N2)
assert torch.equal(full2, torch.arange(N2, dtype=torch.float64))
# the device-side resampler used by the SMC sweep (geconpy_b200/smc.py) is deterministic in (weights, seed) as well
from geconpy_b200.smc import systematic_ancestors  # noqa: E402

anc = systematic_ancestors(full, seed=11)
gathered = [torch.empty_like(anc) for _ in range(world)]
dist.all_gather(gathered, anc)
assert all(torch.equal(g, anc) for g in gathered) and int(anc.max()) < N
# surviving rows only: all_to_all_single with split sizes derived from the (sorted) ancestors == slicing the all-gathered rows
n_loc = 37
rows = torch.arange(rank * n_loc, (rank + 1) * n_loc, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 10.0, 100.0]], dtype=torch.float64)
lw = torch.sin(torch.arange(world * n_loc, dtype=torch.float64))  # same on every rank
anc2 = systematic_ancestors(lw, seed=3)
assert bool((anc2[1:] >= anc2[:-1]).all())
got = fetch_rows(rows, anc2)
ref = gather_rows(rows)[anc2[rank * n_loc : (rank + 1) * n_loc]]
assert torch.equal(got, ref), (rank, got[:3], ref[:3])
try:
    systematic_ancestors(torch.full((8,), float("-inf"), dtype=torch.float64), seed=1)
    raise AssertionError("expected a RuntimeError")
except RuntimeError as e:
    assert "finite weight" in str(e)
if rank == 0:
    print("gloo sharding ok")
dist.destroy_process_group()
