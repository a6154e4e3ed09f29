"""Worker for test_tempered_smc_stage_world_size_2 (torch.distributed.run, gloo, CPU): TemperedSMC.stage with a stub likelihood
(an analytic Gaussian in theta; status = a function of theta) so that the pack / gather / ancestor-slice / row-fetch path and
its determinism across ranks run without a GPU."""
import sys

from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from geconpy_b200.smc import TemperedSMC  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d, n_local = 3, 64


def status_of(theta):
    return (theta[:, 0] > 0.5).to(torch.int32) * 16  # a stand-in for a Blanchard-Kahn failure region


class StubStateSpace:
    """loglik_device(full, Y, out_ll, out_status): ll = -50 |theta - 0.3|^2, -inf where the status is non-zero."""

    def loglik_device(self, full, Y, out_ll=None, out_status=None):
        th = full[:, :d]
        st = status_of(th)
        ll = -50.0 * ((th - 0.3) ** 2).sum(dim=1)
        out_ll.copy_(torch.where(st == 0, ll, torch.full_like(ll, float("-inf"))))
        out_status.copy_(st)
        return out_ll, out_status


def make(exchange):
    g = torch.Generator().manual_seed(100 + rank)
    theta0 = torch.rand((n_local, d), generator=g, dtype=torch.float64)
    lo, hi = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)
    return TemperedSMC(StubStateSpace(), lo, hi, torch.zeros((1, 2), dtype=torch.float64), torch.zeros((4, 1), dtype=torch.float64),
                       step_scale=0.05, seed=5, exchange=exchange).initialise(theta0)


a, b = make("rows"), make("allgather")
for s in range(1, 5):
    sa, sb = a.stage((s / 4) ** 2, s), b.stage((s / 4) ** 2, s)
    # both exchange modes hold the same particles, likelihoods and status words
    assert torch.equal(a.theta, b.theta) and torch.equal(a.ll, b.ll) and torch.equal(a.status, b.status), (rank, s)
    assert sa.ess == sb.ess and sa.accept_rate == sb.accept_rate
    # the status word travels with its particle and the stored likelihood is the particle's own
    assert torch.equal(a.status, status_of(a.theta))
    ll_chk = torch.empty_like(a.ll)
    st_chk = torch.empty_like(a.status)
    a._eval(a.theta, ll_chk, st_chk)
    assert torch.equal(ll_chk, a.ll)
    # resampled particles are finite-likelihood particles: the -inf region never survives
    assert bool(torch.isfinite(a.ll).all()) and int((a.status != 0).sum()) == 0
    # every rank saw the same global weights: ESS agrees across ranks
    ess = [None] * world
    dist.all_gather_object(ess, sa.ess)
    assert len(set(ess)) == 1
# the union of the shards is a resample of the previous global population: all rows come from the pre-stage population
# ragged shards are refused with a clear message
try:
    bad = TemperedSMC(StubStateSpace(), torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64),
                      torch.zeros((1, 2), dtype=torch.float64), torch.zeros((4, 1), dtype=torch.float64))
    bad.initialise(torch.rand((n_local + rank, d), dtype=torch.float64))
    raise AssertionError("expected a ValueError")
except ValueError as e:
    assert "different numbers of particles" in str(e)
if rank == 0:
    print("gloo smc ok")
dist.destroy_process_group()
