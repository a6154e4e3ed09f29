#!/usr/bin/env python
"""bench.py -- fp64 likelihood evaluations / second (solve + Kalman) on B200, and the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload nk|rbc|large|large45] [--draws D] [--impl reference]

One "step" = one pass of the hot path (theta -> A,B,C,D -> cycle reduction -> R, residual -> Blanchard-Kahn count ->
P0 -> Kalman log-likelihood over T_obs = 200 -> gating) over the whole population of draws of the workload.
Workloads are BASELINE.json's configs: ``nk`` = medium New-Keynesian model (full_nk, n = 24, k = 4, p = 3) with
262,144 draws per GPU -- the config the north-star target is quoted on and the default; ``rbc`` = RBC, 65,536 draws;
``large`` = nk_complete_more_shocks (n = 31, k = 9, p = 7), 131,072 draws per GPU; ``large45`` = its synthetic 45-state
composition with rbc_extended (config 4b).  Draws shard across ranks with no
data-path collective; for N > 1 each step ends with the SMC-stage all-gather of the log-likelihoods (NCCL).

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU restatement of the reference path (oracle/, all
host cores, bounded sample) and prints the same line with "impl": "reference".
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fp64 likelihood evals/sec (solve+Kalman)"
UNIT = "evals/s"
FP64_PEAK_FILE = ROOT / "profiles" / "r01_fp64_peak_microbench.json"

WORKLOADS = {
    # name: (model spec, observed states, measurement-error states, draws per GPU, T_obs, box half-width)
    "rbc": dict(model="rbc", observed=["Y"], meas=[], draws=65536, tobs=200, width=None,
                desc="RBC (n=9,k=1,p=1), Sobol draws on the GCN prior bounds, T_obs=200"),
    "nk": dict(model="full_nk", observed=["Y", "pi", "r_G"], meas=["Y", "pi", "r_G"], draws=262144, tobs=200, width=0.10,
               desc="medium NK full_nk (n=24,k=4,p=3), Sobol draws in +-10% boxes around defaults, T_obs=200"),
    "large": dict(model="nk_complete_more_shocks", observed=["Y", "C", "I", "N", "pi", "i", "w"], meas=[], draws=131072, tobs=200,
                  width=0.05, desc="large NK nk_complete_more_shocks (n=31,k=9,p=7), Sobol draws in +-5% boxes, T_obs=200"),
    "large45": dict(model="nk_rbc_composite", observed=["Y", "C", "I", "N", "pi", "i", "w"], meas=[], draws=131072, tobs=200,
                    width=0.05, desc="synthetic 45-state composite nk_complete_more_shocks (+) rbc_extended (n=45,k=13,p=7; SURVEY 8d config 4b), "
                                     "Sobol draws in +-5% boxes, T_obs=200"),
}
SIGMA_SHOCK = 0.01
SIGMA_ERR = 1e-3


# --------------------------------------------------------------------------------------------------- inputs
def make_draws(spec: dict, n_draws: int, width, seed: int, skip: int = 0) -> np.ndarray:
    """Scrambled-Sobol draws (scipy.stats.qmc, as gEconpy/model/sampling.py:122-145 does): on the GCN's prior bounds when
    width is None, else in +-width relative boxes around the defaults intersected with the bounds."""
    from scipy.stats import qmc

    names = list(spec["free_params"])
    th0 = np.array([spec["free_params"][p] for p in names], dtype=np.float64)
    lo, hi = th0.copy(), th0.copy()
    bounds = spec.get("bounds", {})
    for j, p in enumerate(names):
        if width is None and p in bounds:
            lo[j], hi[j] = bounds[p]
        else:
            w = 0.10 if width is None else width
            lo[j], hi[j] = th0[j] - w * abs(th0[j]), th0[j] + w * abs(th0[j])
            if p in bounds:
                eps = 1e-6 * (bounds[p][1] - bounds[p][0])
                lo[j], hi[j] = max(lo[j], bounds[p][0] + eps), min(hi[j], bounds[p][1] - eps)
    # validity bounds: discount factor and AR coefficients strictly inside the unit
    # interval, steady-state targets pinned -- otherwise the steady state is NaN or the model has a unit root
    for j, p in enumerate(names):
        if p in ("beta",):
            lo[j], hi[j] = min(lo[j], 0.998), min(hi[j], 0.999)
        elif p.startswith("rho_"):
            lo[j], hi[j] = min(lo[j], 0.98), min(hi[j], 0.99)
        elif p in ("pi_bar", "phi_pi_obj"):
            lo[j] = hi[j] = th0[j]
    eng = qmc.Sobol(d=len(names), scramble=True, seed=seed)
    if skip:
        eng.fast_forward(skip)
    u = eng.random(n_draws)
    th = lo + u * (hi - lo)
    return np.ascontiguousarray(th)


def full_params(theta: np.ndarray, k: int, n_err: int) -> np.ndarray:
    n = theta.shape[0]
    return np.ascontiguousarray(np.hstack([theta, np.full((n, k), SIGMA_SHOCK), np.full((n, n_err), SIGMA_ERR)]))


def simulate_from_policy(T, R, k, tobs, obs_idx, seed=0):
    """x_t = T x_{t-1} + R eps_t (the recursion of gEconpy/model/simulate.py:171-183), observed columns only."""
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal((tobs, k)) * SIGMA_SHOCK
    x = np.zeros(T.shape[0])
    out = np.zeros((tobs, len(obs_idx)))
    for t in range(tobs):
        x = T @ x + R @ eps[t]
        out[t] = x[obs_idx]
    return out


# --------------------------------------------------------------------------------------------------- flop model
def flop_model(n, k, p, tobs, i_cr, j_lyap):
    """Algorithmic FLOPs per evaluation, SURVEY.md section 8(d) (dense counts, multiply-add = 2)."""
    n3 = float(n) ** 3
    f_cr = i_cr * (38.0 / 3.0) * n3 + (8.0 / 3.0) * n3
    f_r = (8.0 / 3.0) * n3 + 2.0 * n * n * k
    f_res = 6.0 * n3
    f_lyap = j_lyap * 6.0 * n3 + 2.0 * n * n * k + 2.0 * n * k * k
    f_kf = tobs * (8.0 * n3 + 6.0 * n * n * p + 4.0 * n * p * p + 2.0 * n * n + 4.0 * n * p + p**3 / 3.0)
    return dict(cr_solve=f_cr + f_r + f_res, kalman_ll=f_lyap + f_kf, total=f_cr + f_r + f_res + f_lyap + f_kf)


# --------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._idx = gpu_index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------- CPU arm
_ORACLE = {}


def _oracle_eval(args):
    name, theta, sig, herr, observed, Y = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import fast
    from oracle.model import OracleModel

    if name not in _ORACLE:
        _ORACLE[name] = OracleModel(str(ROOT / "geconpy_b200" / "model" / "specs" / f"{name}.json"))
    return fast.loglik(_ORACLE[name], theta, Y, observed, sig, herr if len(herr) else None, tol=1e-8, max_iter=100)


def cpu_reference_rate(wl: dict, Y: np.ndarray, n_sample: int, cores: int, seed: int = 0):
    """Times the oracle (CPU restatement of the reference path) on `n_sample` draws of the workload over a fork pool of
    `cores` workers (the reference's own batch mechanism, perturbation_diagnostics.py:470-490); returns evals/s."""
    import multiprocessing as mp

    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    k = len(spec["shocks"])
    theta = make_draws(spec, n_sample, wl["width"], seed)
    herr_full = np.zeros(len(wl["observed"]))
    for v in wl["meas"]:
        herr_full[wl["observed"].index(v)] = SIGMA_ERR
    sig = np.full(k, SIGMA_SHOCK)
    jobs = [(wl["model"], theta[i], sig, herr_full if wl["meas"] else np.zeros(0), wl["observed"], Y) for i in range(n_sample)]
    _oracle_eval(jobs[0])  # builds the sympy model in the parent so that forked workers inherit it
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if cores > 1:
        with ctx.Pool(cores) as pool:
            lls = pool.map(_oracle_eval, jobs, chunksize=max(1, n_sample // (cores * 4)))
    else:
        lls = [_oracle_eval(j) for j in jobs]
    dt = time.perf_counter() - t0
    return n_sample / dt, dt, int(np.isfinite(lls).sum())


# --------------------------------------------------------------------------------------------------- SMC sweep (config 5)
def prior_box(spec: dict, width):
    """The box make_draws samples from (lo, hi per free parameter)."""
    names = list(spec["free_params"])
    th0 = np.array([spec["free_params"][p] for p in names], dtype=np.float64)
    lo, hi = th0.copy(), th0.copy()
    bounds = spec.get("bounds", {})
    for j, p in enumerate(names):
        lo[j], hi[j] = th0[j] - width * abs(th0[j]), th0[j] + width * abs(th0[j])
        if p in bounds:
            eps = 1e-6 * (bounds[p][1] - bounds[p][0])
            lo[j], hi[j] = max(lo[j], bounds[p][0] + eps), min(hi[j], bounds[p][1] - eps)
        if p in ("beta",):
            lo[j], hi[j] = min(lo[j], 0.998), min(hi[j], 0.999)
        elif p.startswith("rho_"):
            lo[j], hi[j] = min(lo[j], 0.98), min(hi[j], 0.99)
        elif p in ("pi_bar", "phi_pi_obj"):
            lo[j] = hi[j] = th0[j]
    return lo, hi


def run_smc(args, rank, world, local_rank):
    """BASELINE.json config 5: tempered SMC over `--particles` particles per GPU of the medium NK model; one timed
    "step" = one tempering stage = Metropolis mutation (one batched likelihood evaluation of every particle) +
    reweighting + ONE all-gather of (weight, ll, theta) + redundant systematic resampling on every rank."""
    import torch
    import torch.distributed as dist

    from geconpy_b200 import batched
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel
    from geconpy_b200.smc import TemperedSMC

    wl = WORKLOADS["nk"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    n, k, p, tobs = len(spec["variables"]), len(spec["shocks"]), len(wl["observed"]), wl["tobs"]
    cm = CompiledModel(wl["model"])
    ss = BatchedStateSpace(cm).configure(observed_states=wl["observed"], measurement_error=wl["meas"], tol=1e-8, max_iter=100)
    sol = ss.solve(cm.theta_vector()[None], device=str(dev))
    Y = simulate_from_policy(sol["T"][0], sol["R"][0], k, tobs, [cm.var_names.index(v) for v in wl["observed"]])
    Y_d = torch.as_tensor(Y, device=dev)
    n_local = args.particles
    theta0 = torch.as_tensor(make_draws(spec, n_local, wl["width"], seed=0, skip=rank * n_local), device=dev)
    lo, hi = (torch.as_tensor(x, device=dev) for x in prior_box(spec, wl["width"]))
    tail = torch.as_tensor(np.concatenate([np.full(k, SIGMA_SHOCK), np.full(len(wl["meas"]), SIGMA_ERR)]), device=dev)[None]
    smc = TemperedSMC(ss, lo, hi, tail, Y_d, step_scale=0.02, seed=0).initialise(theta0)
    n_stages = args.warmup + args.steps

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    launches0 = None
    evs = []
    with ClockSampler(local_rank) as clk:
        for s in range(1, n_stages + 1):
            if s == args.warmup + 1:
                barrier()
                launches0 = batched.launch_count() + cm.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smc.stage((s / n_stages) ** 2, s)
            e1.record()
            if s > args.warmup:
                evs.append((e0, e1))
        barrier()
    launches = batched.launch_count() + cm.launches - launches0
    ms = float(sum(a.elapsed_time(b) for a, b in evs))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * n_local / (ms_per_step * 1e-3)
    d = theta0.shape[1]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"SMC sweep, medium NK full_nk (n={n},k={k},p={p}), T_obs={tobs}: {n_local} particles per GPU, "
                                   f"{n_stages} tempering stages (first {args.warmup} untimed); one step = one stage = Metropolis mutation "
                                   "(one likelihood evaluation per particle) + all-gather + systematic resampling",
                       "n": n, "k": k, "p": p, "T_obs": tobs, "particles_per_gpu": n_local, "stages": n_stages,
                       "l2": "working set of a stage (A,B,C,D,T,R of 65,536-draw chunks: >1 GB) exceeds L2"},
            "clocks": clk.summary(),
            # particles live on the device from stage to stage: the sweep has no per-step host traffic besides 3 scalars
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 24, "ms_per_step": ms_per_step},
            "gpu_launches": int(launches),
            "smc": {"allgather_bytes_per_rank_per_stage": int(n_local * (2 + d) * 8), "collective": "nccl all_gather_into_tensor" if world > 1 else None,
                    "stages": [dict(phi=round(st.phi, 4), ess=round(st.ess, 1), accept_rate=round(st.accept_rate, 4), mean_ll=round(st.mean_ll, 3),
                                    failed=st.n_failed) for st in smc.stats]}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="nk", choices=sorted(WORKLOADS) + ["smc"])
    ap.add_argument("--particles", type=int, default=131072, help="smc workload: particles per GPU")
    ap.add_argument("--draws", type=int, default=0, help="draws per GPU (default: the workload's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="draws timed on the CPU (default: sized for ~20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gradient", action="store_true", help="skip the gradient-path extra")
    ap.add_argument("--extra-workloads", action="store_true", help="also time the other workloads (kernel-only) at N=1")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "smc":
        if args.impl == "reference":
            args.workload = "nk"  # the CPU comparator of the sweep is the per-particle likelihood rate of the same model
        else:
            return run_smc(args, rank, world, local_rank)
    wl = WORKLOADS[args.workload]
    draws_per_gpu = args.draws or wl["draws"]
    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    n, k, p, tobs = len(spec["variables"]), len(spec["shocks"]), len(wl["observed"]), wl["tobs"]
    cores = os.cpu_count() or 1
    config = {"workload": f"{wl['desc']}; {draws_per_gpu} draws per GPU", "model": None, "n": n, "k": k, "p": p, "T_obs": tobs,
              "draws_per_gpu": draws_per_gpu, "solver": "cycle_reduction tol=1e-8 max_iter=100 (solvability_check defaults) + BK count + resid gate 1e-8",
              "l2": "working set (A,B,C,D,T,R of a 65,536-draw chunk: >1 GB) exceeds L2; a 256 MiB buffer is also written between steps"}
    config.pop("model")

    # ---------------------------------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        # Y from the oracle itself at the default parameters (no GPU on this arm)
        from oracle import statespace as oss
        from oracle.model import OracleModel

        om = OracleModel(str(ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json"))
        th0 = om.theta_vector()
        r0 = oss.loglik(om, th0, np.zeros((1, p)), wl["observed"], np.full(k, SIGMA_SHOCK))
        Y = simulate_from_policy(r0["T"], r0["R"], k, tobs, [om.var_names.index(v) for v in wl["observed"]])
        n_sample = args.cpu_sample or 256 * cores
        rates = []
        for s in range(args.warmup + args.steps):
            rate, dt, nfin = cpu_reference_rate(wl, Y, n_sample, cores, seed=s)
            if s >= args.warmup:
                rates.append((rate, dt))
        value = float(np.mean([r for r, _ in rates]))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([d for _, d in rates]) * 1e3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{n_sample} draws of the workload per step, fork pool of {cores} workers, "
                                           "numba-compiled restatement of the reference path (oracle/fast.py)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist

    from geconpy_b200 import batched
    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cm = CompiledModel(wl["model"])
    ss = BatchedStateSpace(cm).configure(observed_states=wl["observed"], measurement_error=wl["meas"], tol=1e-8, max_iter=100)
    # observations: simulated once from the model at its default parameters, seed 0 (same on every rank)
    sol = ss.solve(cm.theta_vector()[None], device=str(dev))
    Y = simulate_from_policy(sol["T"][0], sol["R"][0], k, tobs, [cm.var_names.index(v) for v in wl["observed"]])
    Y_d = torch.as_tensor(Y, device=dev)
    # this rank's shard of the population (independent Sobol blocks per rank)
    theta = make_draws(spec, draws_per_gpu, wl["width"], seed=0, skip=rank * draws_per_gpu)
    params = full_params(theta, k, len(wl["meas"]))
    params_pinned = torch.from_numpy(params).pin_memory()
    params_d = params_pinned.to(dev)
    ll_d = torch.empty((draws_per_gpu,), dtype=torch.float64, device=dev)
    st_d = torch.empty((draws_per_gpu,), dtype=torch.int32, device=dev)
    it_d = torch.empty((draws_per_gpu,), dtype=torch.int32, device=dev)
    ll_host = torch.empty((draws_per_gpu,), dtype=torch.float64).pin_memory()
    st_host = torch.empty((draws_per_gpu,), dtype=torch.int32).pin_memory()
    gathered = torch.empty((world * draws_per_gpu,), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_device(events=None):
        ss.loglik_device(params_d, Y_d, out_ll=ll_d, out_status=st_d, out_n_iter=it_d, events=events)
        if world > 1:
            dist.all_gather_into_tensor(gathered, ll_d)

    def step_e2e():
        pd = params_pinned.to(dev, non_blocking=True)
        ss.loglik_device(pd, Y_d, out_ll=ll_d, out_status=st_d)
        ll_host.copy_(ll_d, non_blocking=True)
        st_host.copy_(st_d, non_blocking=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, ll_d)
        torch.cuda.synchronize(dev)

    def timed(fn, steps, with_events=False):
        """Times exactly `steps` calls, CUDA events on the launching stream, L2 flush between steps (outside events)."""
        evs, kern_events = [], []
        barrier()
        for _ in range(steps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if with_events:
                ke = []
                fn(ke)
                kern_events.append(ke)
            else:
                fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = float(sum(a.elapsed_time(b) for a, b in evs))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kern_events

    launches0 = batched.launch_count() + cm.launches
    for _ in range(args.warmup):
        step_device()
    barrier()
    with ClockSampler(local_rank) as clk:
        total_ms, kern_events = timed(step_device, args.steps, with_events=True)
    launches = batched.launch_count() + cm.launches - launches0 - 0
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    e2e_ms, _ = timed(step_e2e, args.steps)

    ms_per_step = total_ms / args.steps
    value = world * draws_per_gpu / (ms_per_step * 1e-3)
    e2e_value = world * draws_per_gpu / (e2e_ms / args.steps * 1e-3)

    # ---- extras (reported beside the contract's numbers, never instead of them)
    # (1) the same step with the chunks alternating between two CUDA streams, so that the drain of one chunk's kernels
    #     overlaps the next chunk's; per-kernel event times are not meaningful under overlap, hence a separate pass
    extras = {}
    if draws_per_gpu > ss.chunk:
        ss.n_streams = 2
        for _ in range(2):
            step_device()
        ov_ms, _ = timed(step_device, args.steps)
        ss.n_streams = 1
        extras["two_stream_overlap"] = {"n_streams": 2, "value": world * draws_per_gpu / (ov_ms / args.steps * 1e-3), "unit": UNIT,
                                        "ms_per_step": ov_ms / args.steps}
    # (2) the gradient path: log-likelihood + d/d(theta, sigma) for a 32,768-draw slice
    if ss.n_aug <= 48 and not args.no_gradient:
        ng = min(draws_per_gpu, 32768)
        for _ in range(2):
            ss.loglik_and_grad_device(params_d[:ng], Y_d)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gev = []
        g0.record()
        _, gr, gst = ss.loglik_and_grad_device(params_d[:ng], Y_d, events=gev)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1)
        gk = {}
        for nm, a, b in gev:
            gk[nm] = gk.get(nm, 0.0) + a.elapsed_time(b)
        extras["gradient"] = {"value": world * ng / (gms * 1e-3), "unit": "likelihood+gradient evals/s", "draws": int(ng), "ms": gms,
                              "kernel_ms": gk, "n_param": int(gr.shape[1]), "finite": bool(torch.isfinite(gr).all().item()),
                              "ok": int((gst == 0).sum().item())}

    # ---- per-kernel durations measured live (CUDA events on the launching stream), roofline of the dominant kernel
    kt = {}
    for ke in kern_events:
        for name, a, b in ke:
            kt.setdefault(name, []).append(a.elapsed_time(b))
    n_chunks = max(1, len(kern_events[0]) // max(1, len({nm for nm, _, _ in kern_events[0]}))) if kern_events else 1
    kernel_ms = {nm: float(np.sum(v) / args.steps) for nm, v in kt.items()}  # per step (all chunks)
    status = st_d.cpu().numpy()
    n_iter = it_d.cpu().numpy()
    ok = status == 0
    i_cr = float(n_iter[(status & 0x207) == 0].mean()) if ((status & 0x207) == 0).any() else float("nan")
    j_lyap = 11.0
    fm = flop_model(n, k, p, tobs, i_cr, j_lyap)
    # the Kalman kernel runs on the variables the likelihood depends on (states + observed): its own algorithmic count
    fm["kalman_ll_dense_n"] = fm["kalman_ll"]
    fm["kalman_ll"] = flop_model(ss.n_filter, k, p, tobs, i_cr, j_lyap)["kalman_ll"]
    fm["n_filter"] = ss.n_filter
    n_eval_kf = int(((status & 0x400) == 0).sum())  # draws the Kalman kernel actually filtered (not gated out)
    peak = json.loads(FP64_PEAK_FILE.read_text())["dfma_tflops"] if FP64_PEAK_FILE.exists() else 36.6
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else "kalman_ll"
    units = n_eval_kf if dom == "kalman_ll" else draws_per_gpu
    launches_per_step = max(1, math.ceil(draws_per_gpu / ss.chunk))
    achieved = (fm.get(dom, 0.0) * units) / (kernel_ms[dom] * 1e-3) / 1e12 if dom in fm and kernel_ms.get(dom) else None
    # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture (per-draw figure x draws)
    traffic = None
    tfile = ROOT / "profiles" / "r01_ncu_traffic.json"
    if tfile.exists() and args.workload == "nk":
        for kname, v in json.loads(tfile.read_text()).items():
            if isinstance(v, dict) and kname.startswith(dom):  # e.g. "kalman_ll_warp_kernel<16, 3, 4>" for dom = "kalman_ll"
                traffic = v["dram_bytes_per_draw"] * min(draws_per_gpu, ss.chunk)
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": "measured DFMA burst on this pool's B200 (profiles/r01_fp64_peak_microbench.json); "
                               "MEASURED_PEAKS.json holds no fp64 figure",
                "flops_per_eval": fm, "mean_cr_iterations": i_cr, "assumed_lyapunov_doublings": j_lyap,
                "kernel_ms_per_step": kernel_ms, "launches_per_step": launches_per_step,
                "avg_launch_ms": (kernel_ms[dom] / launches_per_step) if kernel_ms.get(dom) else None,
                "note": "achieved = SURVEY 8(d) dense algorithmic FLOPs of the draws the kernel processed / its CUDA-event time"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params.nbytes),
                    "d2h_bytes_per_step": int(draws_per_gpu * (8 + 4)), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roofline, "extras": extras,
            "draw_outcomes": {"ok": int(ok.sum()), "gated_minus_inf": int(((status & 0x400) != 0).sum()),
                              "bk_violated": int(((status & 0x10) != 0).sum()), "bk_inconclusive": int(((status & 0x20) != 0).sum()),
                              "cr_not_converged": int(((status & 0x1) != 0).sum()), "jacobian_nonfinite": int(((status & 0x200) != 0).sum()),
                              "not_pd": int(((status & 0x80) != 0).sum()), "lyap": int(((status & 0x40) != 0).sum()),
                              "of": int(draws_per_gpu), "rank": 0}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # parity at the benchmark's own size: 48 draws of THIS run, spread over the whole population, against the CPU
        # restatement (checker only; nothing of it is timed here)
        pick = np.unique(np.linspace(0, draws_per_gpu - 1, 48).astype(np.int64))
        ll_gpu = ll_d.cpu().numpy()[pick]
        herr_full = np.zeros(len(wl["observed"]))
        for v in wl["meas"]:
            herr_full[wl["observed"].index(v)] = SIGMA_ERR
        ll_cpu = np.array([_oracle_eval((wl["model"], theta[i], np.full(k, SIGMA_SHOCK), herr_full if wl["meas"] else np.zeros(0),
                                        wl["observed"], Y)) for i in pick])
        both = np.isfinite(ll_gpu) & np.isfinite(ll_cpu)
        line["parity_spot_check"] = {"draws": int(pick.size), "finite_on_both": int(both.sum()),
                                     "flags_agree": bool((np.isfinite(ll_gpu) == np.isfinite(ll_cpu)).all()),
                                     "max_abs_ll_error": float(np.abs(ll_gpu[both] - ll_cpu[both]).max()) if both.any() else None,
                                     "tolerance": 1e-7}
        n_sample = args.cpu_sample or 256 * cores
        rate, dt, nfin = cpu_reference_rate(wl, Y, n_sample, cores, seed=0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_sample} draws of the same workload in {dt:.1f} s, fork pool of {cores} workers, "
                                          "numba-compiled restatement of the reference path (oracle/fast.py)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
