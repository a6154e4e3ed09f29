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
    # a prior wide enough that every failure class of the path is populated (and a short iteration budget, as configure()'s
    # default max_iter = 50 is for gensys): NaN steady states (alpha > 1), cycle reduction out of iterations, Blanchard-Kahn
    # violated (unit-root shocks, passive monetary policy), residual gate -- each >= 15 % of the draws
    "nk_wide": dict(model="full_nk", observed=["Y", "pi", "r_G"], meas=["Y", "pi", "r_G"], draws=65536, tobs=200, width=None, max_iter=16,
                    box={"alpha": (0.25, 1.15), "beta": (0.97, 0.999), "delta": (0.01, 0.05), "eta_p": (0.5, 0.9), "eta_w": (0.5, 0.9),
                         "gamma_I": (5.0, 15.0), "gamma_R": (0.5, 0.95), "gamma_Y": (0.0, 0.3), "gamma_pi": (0.85, 2.5), "phi_H": (0.3, 0.7),
                         "phi_pi_obj": (1.0, 1.0), "psi_p": (0.4, 0.8), "psi_w": (0.5, 0.9), "rho_pi_dot": (0.7, 1.03),
                         "rho_preference": (0.7, 1.03), "rho_technology": (0.7, 1.03), "sigma_C": (1.2, 3.0), "sigma_L": (1.0, 2.5)},
                    desc="medium NK full_nk, WIDE prior box (every failure class populated), max_iter=16, Sobol draws, T_obs=200"),
    "large45": dict(model="nk_rbc_composite", observed=["Y", "C", "I", "N", "pi", "i", "w"], meas=[], draws=131072, tobs=200,
                    width=0.05, desc="synthetic 45-state composite nk_complete_more_shocks (+) rbc_extended (n=45,k=13,p=7; SURVEY 8d config 4b), "
                                     "Sobol draws in +-5% boxes, T_obs=200"),
}
SIGMA_SHOCK = 0.01
SIGMA_ERR = 1e-3


# --------------------------------------------------------------------------------------------------- inputs
def make_draws(spec: dict, n_draws: int, width, seed: int, skip: int = 0, box=None) -> np.ndarray:
    """Scrambled-Sobol draws (scipy.stats.qmc, as gEconpy/model/sampling.py:122-145 does): on the GCN's prior bounds when
    width is None, else in +-width relative boxes around the defaults intersected with the bounds; ``box`` = explicit
    {parameter: (lo, hi)} (no validity clamps: the wide-prior workload WANTS invalid draws)."""
    from scipy.stats import qmc

    names = list(spec["free_params"])
    if box is not None:
        lo = np.array([box[p][0] for p in names], dtype=np.float64)
        hi = np.array([box[p][1] for p in names], dtype=np.float64)
        eng = qmc.Sobol(d=len(names), scramble=True, seed=seed)
        if skip:
            eng.fast_forward(skip)
        return np.ascontiguousarray(lo + eng.random(n_draws) * (hi - lo))
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
    return fast.loglik(_ORACLE[name], theta, Y, observed, sig, herr if len(herr) else None, tol=1e-8, max_iter=int(os.environ.get("GECON_BENCH_MAX_ITER", "100")))


def cpu_reference_rate(wl: dict, Y: np.ndarray, n_sample: int, cores: int, seed: int = 0):
    """Times the oracle (CPU restatement of the reference path) on `n_sample` draws of the workload over a fork pool of
    `cores` workers (the reference's own batch mechanism, perturbation_diagnostics.py:470-490); returns evals/s."""
    import multiprocessing as mp

    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    k = len(spec["shocks"])
    theta = make_draws(spec, n_sample, wl["width"], seed, box=wl.get("box"))
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
            "smc": {"allgather_bytes_per_rank_per_stage": int(n_local * 3 * 8),
                    "collective": "nccl all_gather_into_tensor of (log-weight, ll, status) + all_to_all_single of the surviving rows" if world > 1 else None,
                    "stages": [dict(phi=round(st.phi, 4), ess=round(st.ess, 1), accept_rate=round(st.accept_rate, 4), mean_ll=round(st.mean_ll, 3),
                                    failed=st.n_failed) for st in smc.stats]}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------- CPU arm
def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def try_real_reference(wl: dict, Y: np.ndarray):
    """The REAL reference's compiled data log-likelihood for this workload, if gEconpy + pymc + pymc_extras import here (a
    driver-side install under baseline/_ref, a developer environment): ``f(theta_row) -> float`` and a description, else None.
    Built by the same code that writes the Kalman goldens (tests/golden/make_kalman_goldens.py)."""
    try:
        sys.path.insert(0, str(ROOT / "tests" / "golden"))
        import make_kalman_goldens as mk

        gE, src = mk.find_reference()
        if gE is None:
            return None
        import pandas as pd

        spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
        rel = spec.get("derived_from")
        pkg_root = Path(gE.__file__).resolve().parent.parent
        gcn = next((pth for pth in ((src / rel) if src and rel else None, (pkg_root / rel) if rel else None) if pth is not None and pth.exists()), None)
        if gcn is None:
            return None
        df = pd.DataFrame(Y, index=pd.date_range("2000-01-01", periods=len(Y), freq="QS"), columns=wl["observed"])
        f, names, _ = mk.build_reference_logp(gcn, dict(observed_states=wl["observed"], measurement_error=wl["meas"] or None), df)
        pn = list(spec["free_params"])
        fixed = {f"sigma_{s_}": SIGMA_SHOCK for s_ in spec["shocks"]}
        fixed.update({f"error_sigma_{v}": SIGMA_ERR for v in wl["meas"]})

        def logp(theta_row):
            point = dict(zip(pn, theta_row))
            point.update(fixed)
            return float(f({nm: np.asarray(point[nm], dtype=np.float64) for nm in names}))

        logp(np.array([spec["free_params"][p_] for p_ in pn], dtype=np.float64))
        return logp, f"gEconpy {getattr(gE, '__version__', '?')} compiled logp (statespace_from_gcn -> configure -> build_statespace_graph)"
    except Exception:  # anything missing or incompatible: the port is the baseline
        return None


def cpu_arm(wl: dict, Y: np.ndarray, cores: int, seconds: float = 6.0) -> dict:
    """Single-core and all-core rates of the CPU implementation of the path on a bounded sample of the workload:
    the real reference when it imports here, else the numba-compiled restatement oracle/fast.py (LAPACK through numba, like the
    reference's own numba kernels).  ~`seconds` of wall clock per leg."""
    os.environ["GECON_BENCH_MAX_ITER"] = str(wl.get("max_iter", 100))
    real = try_real_reference(wl, Y)
    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    if real is not None:
        logp, what = real
        theta = make_draws(spec, 4096, wl["width"], 0, box=wl.get("box"))
        t0, n1 = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds and n1 < len(theta):
            logp(theta[n1])
            n1 += 1
        rate1 = n1 / (time.perf_counter() - t0)
        return {"value": rate1, "unit": UNIT, "cores": 1, "kind": "reference", "single_core": rate1, "all_core": None, "cpu_model": cpu_model_name(),
                "wall_s": seconds,
                "sample": f"{n1} draws of the workload in {seconds:.0f} s on one core: {what}"}
    rate_probe, _dt, _ = cpu_reference_rate(wl, Y, 32, 1, seed=0)  # (also JIT-compiles the port)
    n1 = max(32, int(rate_probe * seconds))
    rate1, dt1, _ = cpu_reference_rate(wl, Y, n1, 1, seed=1)
    nall = max(cores * 32, int(rate1 * cores * seconds * 0.7))
    rate_all, dt_all, _ = cpu_reference_rate(wl, Y, nall, cores, seed=2) if cores > 1 else (rate1, dt1, 0)
    return {"value": rate_all, "unit": UNIT, "cores": cores, "kind": "port", "single_core": rate1, "all_core": rate_all, "cpu_model": cpu_model_name(),
            "wall_s": dt_all,
            "sample": f"{nall} draws of the workload in {dt_all:.1f} s over a fork pool of {cores} workers (single core: {n1} draws in {dt1:.1f} s); "
                      "numba-compiled restatement of the reference path (oracle/fast.py); gEconpy / pymc_extras do not import here"}


def config1_single_draw(dev) -> dict:
    """BASELINE config 1: RBC, T_obs = 100, ONE parameter draw: ms per evaluation on one host core (the CPU restatement, numpy and
    numba) next to the latency of the same single evaluation through the GPU path (host arrays in, host arrays out)."""
    import torch

    from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel
    from oracle import fast
    from oracle import statespace as oss
    from oracle.model import OracleModel

    om = OracleModel(str(ROOT / "geconpy_b200" / "model" / "specs" / "rbc.json"))
    th0 = om.theta_vector()
    r0 = oss.loglik(om, th0, np.zeros((1, 1)), ["Y"], [SIGMA_SHOCK])
    x = oss.simulate(r0["T"], r0["R"], [SIGMA_SHOCK], 100, seed=0)
    Y = np.ascontiguousarray(x[:, [om.var_names.index("Y")]])

    def per_call(fn, budget=1.5):
        fn()
        t0, n_ = time.perf_counter(), 0
        while time.perf_counter() - t0 < budget:
            fn()
            n_ += 1
        return (time.perf_counter() - t0) / n_ * 1e3

    ms_numpy = per_call(lambda: oss.loglik(om, th0, Y, ["Y"], [SIGMA_SHOCK], tol=1e-8, max_iter=100))
    ms_numba = per_call(lambda: fast.loglik(om, th0, Y, ["Y"], np.array([SIGMA_SHOCK]), None, tol=1e-8, max_iter=100))
    ss = BatchedStateSpace(CompiledModel("rbc")).configure(observed_states=["Y"], tol=1e-8, max_iter=100)
    full = np.hstack([th0, [SIGMA_SHOCK]])[None]
    ll_gpu, _st = ss.loglik(full, Y)
    ms_gpu = per_call(lambda: ss.loglik(full, Y))
    ref = oss.loglik(om, th0, Y, ["Y"], [SIGMA_SHOCK], tol=1e-8, max_iter=100)["ll"]
    return {"workload": "RBC (n=9,k=1,p=1), default parameters, T_obs=100, ONE draw (BASELINE config 1)", "cpu_ms_per_eval_numpy_oracle": ms_numpy,
            "cpu_ms_per_eval_numba_port": ms_numba, "cpu_cores": 1, "gpu_ms_per_eval_latency": ms_gpu, "ll_gpu": float(ll_gpu[0]),
            "abs_ll_error_vs_oracle": float(abs(ll_gpu[0] - ref))}


# --------------------------------------------------------------------------------------------------- B200 arm
class Workload:
    """One configured workload on one device: model, state space, synthetic observations, this rank's shard of the draws."""

    def __init__(self, name, draws_per_gpu, rank, dev):
        import torch

        from geconpy_b200.model.compiled import BatchedStateSpace, CompiledModel

        wl = WORKLOADS[name]
        self.name, self.wl, self.dev, self.n_draws = name, wl, dev, int(draws_per_gpu or wl["draws"])
        self.spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
        self.n, self.k, self.p, self.tobs = len(self.spec["variables"]), len(self.spec["shocks"]), len(wl["observed"]), wl["tobs"]
        self.max_iter = int(wl.get("max_iter", 100))
        self.cm = CompiledModel(wl["model"])
        self.ss = BatchedStateSpace(self.cm).configure(observed_states=wl["observed"], measurement_error=wl["meas"], tol=1e-8, max_iter=self.max_iter)
        sol = self.ss.solve(self.cm.theta_vector()[None], device=str(dev))
        self.Y = simulate_from_policy(sol["T"][0], sol["R"][0], self.k, self.tobs, [self.cm.var_names.index(v) for v in wl["observed"]])
        self.Y_d = torch.as_tensor(self.Y, device=dev)
        self.theta = make_draws(self.spec, self.n_draws, wl["width"], seed=0, skip=rank * self.n_draws, box=wl.get("box"))
        self.params = full_params(self.theta, self.k, len(wl["meas"]))
        self.params_pinned = torch.from_numpy(self.params).pin_memory()
        self.params_d = self.params_pinned.to(dev)
        self.ll_d = torch.empty((self.n_draws,), dtype=torch.float64, device=dev)
        self.st_d = torch.empty((self.n_draws,), dtype=torch.int32, device=dev)
        self.it_d = torch.empty((self.n_draws,), dtype=torch.int32, device=dev)

    def config(self):
        path = "fused C entry point gecon_model_loglik (compact Jacobian; no dense A,B,C,D)" if self.ss.fused else "kernel-by-kernel Python pipeline"
        return {"workload": f"{self.wl['desc']}; {self.n_draws} draws per GPU", "n": self.n, "k": self.k, "p": self.p, "T_obs": self.tobs,
                "draws_per_gpu": self.n_draws, "path": path,
                "solver": f"cycle_reduction tol=1e-8 max_iter={self.max_iter} + BK count + resid gate 1e-8",
                "l2": "working set of a 65,536-draw chunk (compact Jacobians, T, R: > 130 MB at n = 24) exceeds L2 and a 256 MiB buffer is "
                      "written between steps"}

    def outcomes(self):
        st = self.st_d.cpu().numpy()
        n_ = float(len(st))
        cls = {"ok": st == 0, "gated_minus_inf": (st & 0x400) != 0, "jacobian_nonfinite": (st & 0x200) != 0, "cr_not_converged": (st & 0x1) != 0,
               "bk_violated": (st & 0x10) != 0, "bk_inconclusive": (st & 0x20) != 0, "resid_gate": (st & 0x8) != 0, "singular": (st & 0x4) != 0,
               "not_pd": (st & 0x80) != 0, "lyap": (st & 0x40) != 0}
        out = {k_: int(v.sum()) for k_, v in cls.items()}
        out["fractions"] = {k_: round(float(v.sum()) / n_, 4) for k_, v in cls.items()}
        out.update(of=int(n_), rank=0)
        return out


def time_workload(w: Workload, steps, warmup, world, flush, peak, with_e2e=True):
    """Device-resident and end-to-end timings of one workload + the roofline of its dominant kernel."""
    import torch
    import torch.distributed as dist

    from geconpy_b200 import batched, parallel

    dev, ss = w.dev, w.ss
    gathered = torch.empty((world * w.n_draws,), dtype=torch.float64, device=dev) if world > 1 else None
    ll_host = torch.empty((w.n_draws,), dtype=torch.float64).pin_memory()
    st_host = torch.empty((w.n_draws,), dtype=torch.int32).pin_memory()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_device(events=None):
        ss.loglik_device(w.params_d, w.Y_d, out_ll=w.ll_d, out_status=w.st_d, out_n_iter=w.it_d, events=events)
        if world > 1:
            parallel.gather_loglik(w.ll_d, world * w.n_draws, out=gathered)

    def step_e2e():
        pd_ = w.params_pinned.to(dev, non_blocking=True)
        ss.loglik_device(pd_, w.Y_d, out_ll=w.ll_d, out_status=w.st_d)
        ll_host.copy_(w.ll_d, non_blocking=True)
        st_host.copy_(w.st_d, non_blocking=True)
        if world > 1:
            parallel.gather_loglik(w.ll_d, world * w.n_draws, out=gathered)
        torch.cuda.synchronize(dev)

    def timed(fn, n_steps, with_events=False):
        """Times exactly `n_steps` calls, CUDA events on the launching stream, L2 flush between steps (outside the events)."""
        evs, kern_events = [], []
        barrier()
        for _ in range(n_steps):
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

    launches0 = batched.launch_count() + w.cm.launches
    for _ in range(warmup):
        step_device()
    total_ms, _ = timed(step_device, steps)                       # the headline: no per-kernel events inside the timed region
    launches = batched.launch_count() + w.cm.launches - launches0
    _kms, kern_events = timed(step_device, steps, with_events=True)  # same steps again with CUDA events around every kernel
    e2e = None
    if with_e2e:
        for _ in range(max(1, warmup // 2)):
            step_e2e()
        e2e_ms, _ = timed(step_e2e, steps)
        e2e = {"value": world * w.n_draws / (e2e_ms / steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(w.params.nbytes),
               "d2h_bytes_per_step": int(w.n_draws * (8 + 4)), "ms_per_step": e2e_ms / steps}
    ms_per_step = total_ms / steps
    kt = {}
    for ke in kern_events:
        for name, a, b in ke:
            if name == "__fused_ms__":  # the fused C entry point times its own kernels (CUDA events on the launching stream)
                for nm, ms in a.items():
                    kt.setdefault(nm, []).append(ms)
            else:
                kt.setdefault(name, []).append(a.elapsed_time(b))
    kernel_ms = {nm: float(np.sum(v) / steps) for nm, v in kt.items()}
    status, n_iter = w.st_d.cpu().numpy(), w.it_d.cpu().numpy()
    solved = (status & 0x207) == 0
    i_cr = float(n_iter[solved].mean()) if solved.any() else float("nan")
    i_all = float(n_iter[(status & 0x200) == 0].mean()) if ((status & 0x200) == 0).any() else float("nan")
    j_lyap = 11.0
    fm = flop_model(w.n, w.k, w.p, w.tobs, i_all, j_lyap)  # every draw with a finite Jacobian iterates (failed ones to max_iter)
    fm["kalman_ll_dense_n"] = fm["kalman_ll"]
    fm["kalman_ll"] = flop_model(ss.n_filter, w.k, w.p, w.tobs, i_all, j_lyap)["kalman_ll"]
    fm["n_filter"] = ss.n_filter
    n_eval_kf = int(((status & 0x400) == 0).sum())
    n_eval_cr = int(((status & 0x200) == 0).sum())
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else "kalman_ll"
    units = n_eval_kf if dom == "kalman_ll" else n_eval_cr
    launches_per_step = max(1, math.ceil(w.n_draws / ss.chunk))
    achieved = (fm.get(dom, 0.0) * units) / (kernel_ms[dom] * 1e-3) / 1e12 if dom in fm and kernel_ms.get(dom) else None
    per_kernel = {}
    for nm, units_k in (("cr_solve", n_eval_cr), ("kalman_ll", n_eval_kf)):
        if kernel_ms.get(nm):
            tf = fm[nm] * units_k / (kernel_ms[nm] * 1e-3) / 1e12
            per_kernel[nm] = {"ms_per_step": kernel_ms[nm], "achieved_tflops": tf, "frac": tf / peak["dfma_tflops"]}
    traffic = None
    tfile = ROOT / "profiles" / "r02_ncu_traffic.json"
    if tfile.exists() and w.name == "nk":
        for kname, v in json.loads(tfile.read_text()).items():
            if isinstance(v, dict) and v.get("stage") == dom:
                traffic = v["dram_bytes_per_draw"] * min(w.n_draws, ss.chunk)
    roofline = {"bound": "fp64", "kernel": dom, "achieved": achieved, "peak": peak["dfma_tflops"], "unit": "TFLOP/s",
                "frac": (achieved / peak["dfma_tflops"]) if achieved else None, "traffic": traffic,
                "peak_source": "DFMA peak measured in this run by gecon_fp64_peak (register-resident chains); MEASURED_PEAKS.json holds no fp64 "
                               f"figure; committed burst figure of round 1: {peak.get('committed_dfma_tflops')} TFLOP/s",
                "dmma_peak_tflops": peak["dmma_tflops"], "flops_per_eval": fm, "mean_cr_iterations_converged": i_cr,
                "mean_cr_iterations_all": i_all, "assumed_lyapunov_doublings": j_lyap, "kernel_ms_per_step": kernel_ms,
                "per_kernel": per_kernel, "launches_per_step": launches_per_step,
                "avg_launch_ms": (kernel_ms[dom] / launches_per_step) if kernel_ms.get(dom) else None,
                "note": "achieved = SURVEY 8(d) dense algorithmic FLOPs (at the dimension the kernel runs: n for the solver, the exactly "
                        "reduced filter dimension for the filter) of the draws the kernel processed / its CUDA-event time"}
    return {"value": world * w.n_draws / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "draw_outcomes": w.outcomes()}


def parity_spot_check(w: Workload, n_pick=48):
    """Parity at the benchmark's own size: draws of THIS run, spread over the whole population, against the CPU restatement
    (checker only; nothing of it is timed here)."""
    os.environ["GECON_BENCH_MAX_ITER"] = str(w.max_iter)
    pick = np.unique(np.linspace(0, w.n_draws - 1, n_pick).astype(np.int64))
    ll_gpu = w.ll_d.cpu().numpy()[pick]
    herr_full = np.zeros(len(w.wl["observed"]))
    for v in w.wl["meas"]:
        herr_full[w.wl["observed"].index(v)] = SIGMA_ERR
    ll_cpu = np.array([_oracle_eval((w.wl["model"], w.theta[i], np.full(w.k, SIGMA_SHOCK), herr_full if w.wl["meas"] else np.zeros(0),
                                    w.wl["observed"], w.Y)) for i in pick])
    both = np.isfinite(ll_gpu) & np.isfinite(ll_cpu)
    n_dis = int((np.isfinite(ll_gpu) != np.isfinite(ll_cpu)).sum())
    return {"draws": int(pick.size), "finite_on_both": int(both.sum()), "flags_agree": n_dis == 0, "gate_disagreements": n_dis,
            "note": ("a wide prior produces draws on which the reference's own Blanchard-Kahn count is a numerical artefact (eigenvalues within its "
                     "LAPACK error of the unit circle, counts that change under its own 1e-8 regularisation): classified and bounded in "
                     "tests/test_gpu_pipeline.py::test_wide_prior_population_failure_classes_at_scale") if w.name == "nk_wide" else None,
            "max_abs_ll_error": float(np.abs(ll_gpu[both] - ll_cpu[both]).max()) if both.any() else None, "tolerance": 1e-7}


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
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="wall clock per leg of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gradient", action="store_true", help="skip the gradient-path extra")
    ap.add_argument("--no-extras", action="store_true", help="only the contract's line: no other workloads, no config 1, no SMC / strong-scaling extras")
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
    cores = os.cpu_count() or 1

    # ---------------------------------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import statespace as oss
        from oracle.model import OracleModel

        spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
        n, k, p, tobs = len(spec["variables"]), len(spec["shocks"]), len(wl["observed"]), wl["tobs"]
        draws_per_gpu = args.draws or wl["draws"]
        om = OracleModel(str(ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json"))
        r0 = oss.loglik(om, om.theta_vector(), np.zeros((1, p)), wl["observed"], np.full(k, SIGMA_SHOCK))
        Y = simulate_from_policy(r0["T"], r0["R"], k, tobs, [om.var_names.index(v) for v in wl["observed"]])
        rates = []
        for s_ in range(args.warmup + args.steps):  # each step: a bounded sample of the workload on all host cores
            arm = cpu_arm(wl, Y, cores, seconds=max(2.0, args.cpu_seconds / 2))
            if s_ >= args.warmup:
                rates.append(arm)
        value = float(np.mean([r["value"] for r in rates]))
        last = rates[-1]
        config = {"workload": f"{wl['desc']}; {draws_per_gpu} draws per GPU", "n": n, "k": k, "p": p, "T_obs": tobs, "draws_per_gpu": draws_per_gpu,
                  "solver": f"cycle_reduction tol=1e-8 max_iter={wl.get('max_iter', 100)} + BK count + resid gate 1e-8"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean([r["wall_s"] for r in rates]) * 1e3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": dict(last, value=value, single_core=float(np.mean([r["single_core"] for r in rates]))),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------------------------------- B200 arm
    import torch
    import torch.distributed as dist

    from geconpy_b200 import batched

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = batched.fp64_peak()
    if FP64_PEAK_FILE.exists():
        peak["committed_dfma_tflops"] = json.loads(FP64_PEAK_FILE.read_text())["dfma_tflops"]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    w = Workload(args.workload, args.draws, rank, dev)
    with ClockSampler(local_rank) as clk:
        res = time_workload(w, args.steps, args.warmup, world, flush, peak)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": w.config(), "clocks": clk.summary(), "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
            "roofline": res["roofline"], "draw_outcomes": res["draw_outcomes"], "fp64_peak_in_run": peak}
    extras = {}
    ss = w.ss
    # (1) the kernel-by-kernel pipeline with its chunks alternating between two CUDA streams (round 1's overlap experiment)
    if w.n_draws > ss.chunk and not args.no_extras:
        fused0, ss.fused, ss.n_streams = ss.fused, False, 2
        for _ in range(2):
            ss.loglik_device(w.params_d, w.Y_d, out_ll=w.ll_d, out_status=w.st_d)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ss.loglik_device(w.params_d, w.Y_d, out_ll=w.ll_d, out_status=w.st_d)
        e1.record()
        torch.cuda.synchronize(dev)
        ov = e0.elapsed_time(e1) / args.steps
        ss.fused, ss.n_streams = fused0, 1
        extras["unfused_two_stream_overlap"] = {"value": w.n_draws / (ov * 1e-3), "unit": UNIT, "ms_per_step": ov, "per_gpu": True}
    # (2) the gradient path: log-likelihood + d/d(theta, sigma) for a 32,768-draw slice
    if ss.n_aug <= 48 and not args.no_gradient:
        ng = min(w.n_draws, 32768)
        for _ in range(2):
            ss.loglik_and_grad_device(w.params_d[:ng], w.Y_d)
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gev = []
        g0.record()
        _, gr, gst = ss.loglik_and_grad_device(w.params_d[:ng], w.Y_d, events=gev)
        g1.record()
        torch.cuda.synchronize(dev)
        gms = g0.elapsed_time(g1)
        gk = {}
        for nm, a, b in gev:
            gk[nm] = gk.get(nm, 0.0) + a.elapsed_time(b)
        extras["gradient"] = {"value": ng / (gms * 1e-3), "unit": "likelihood+gradient evals/s", "draws": int(ng), "ms": gms, "kernel_ms": gk,
                              "n_param": int(gr.shape[1]), "finite": bool(torch.isfinite(gr).all().item()), "ok": int((gst == 0).sum().item()),
                              "per_gpu": True}
    if world == 1 and rank == 0 and not args.no_extras:
        # (3) the other BASELINE configurations, kernel-timed on this GPU (each with its own roofline), and the wide-prior population
        others = {}
        for nm in ("rbc", "large", "large45", "nk_wide"):
            if nm == args.workload:
                continue
            wo = Workload(nm, 0, 0, dev)
            ro = time_workload(wo, 3, 3, 1, flush, peak, with_e2e=False)
            others[nm] = {"config": wo.config(), "value": ro["value"], "unit": UNIT, "ms_per_step": ro["ms_per_step"], "steps": 3, "warmup": 3,
                          "roofline": {k_: ro["roofline"][k_] for k_ in ("kernel", "achieved", "peak", "frac", "kernel_ms_per_step", "per_kernel",
                                                                         "mean_cr_iterations_all")},
                          "draw_outcomes": ro["draw_outcomes"]}
            # parity at each workload's own size: draws of THIS run against the CPU restatement (|ll - oracle| <= 1e-7, same gating)
            others[nm]["parity_spot_check"] = parity_spot_check(wo, {"rbc": 32, "large": 16, "large45": 12, "nk_wide": 96}[nm])
            if nm == "nk_wide":
                # the same population with the Blanchard-Kahn count restricted to the draws the gate has not rejected yet
                # (configure(bk_on_rejected_draws=False) -> gecon_pipeline_args.check_bk = 2): identical log-likelihoods, what a sampler needs
                ll_exact = wo.ll_d.clone()
                wo.ss.configure(observed_states=wo.wl["observed"], measurement_error=wo.wl["meas"], tol=1e-8, max_iter=wo.max_iter,
                                bk_on_rejected_draws=False)
                rg = time_workload(wo, 3, 3, 1, flush, peak, with_e2e=False)
                others[nm]["gate_only_bk"] = {"value": rg["value"], "unit": UNIT, "ms_per_step": rg["ms_per_step"],
                                              "kernel_ms_per_step": rg["roofline"]["kernel_ms_per_step"],
                                              "ll_identical": bool(torch.equal(torch.nan_to_num(ll_exact, neginf=-1e300), torch.nan_to_num(wo.ll_d, neginf=-1e300)))}
            del wo
            torch.cuda.empty_cache()
        extras["workloads"] = others
        # (4) BASELINE config 1: one RBC draw, T_obs = 100, one host core
        extras["config1_single_draw"] = config1_single_draw(dev)
    if world > 1 and not args.no_extras:
        # (5) one SMC stage per rank with 131,072 particles, the exchange timed on its own, both exchange modes; and a
        #     strong-scaling point: 262,144 draws in total, split over the ranks
        extras["smc_stage"] = smc_stage_extra(w, world, rank, dev)
        ws_ = Workload(args.workload, max(1, 262144 // world), rank, dev)
        rs = time_workload(ws_, 3, 2, world, flush, peak, with_e2e=False)
        extras["strong_scaling_point"] = {"total_draws": int(ws_.n_draws * world), "draws_per_gpu": int(ws_.n_draws), "value": rs["value"], "unit": UNIT,
                                          "ms_per_step": rs["ms_per_step"], "scaling": "strong"}
    line["extras"] = extras
    if rank == 0 and not args.no_cpu_baseline:
        # parity at the benchmark's own size at every N: 48 draws of rank 0's shard of THIS run against the CPU restatement
        line["parity_spot_check"] = parity_spot_check(w)
        if world == 1:  # (the CPU baseline is timed at N = 1 only)
            line["cpu_baseline"] = cpu_arm(wl, w.Y, cores, seconds=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def smc_stage_extra(w: Workload, world, rank, dev, n_local=131072, stages=4):
    """Config 5 next to the likelihood sweep: tempering stages of the medium NK posterior with the exchange timed separately, for the
    row-fetch exchange (24 bytes per particle gathered + surviving rows) and for round 1's all-gather of the parameter rows."""
    import torch
    import torch.distributed as dist

    from geconpy_b200.smc import TemperedSMC

    wl = WORKLOADS["nk"]
    spec = json.loads((ROOT / "geconpy_b200" / "model" / "specs" / f"{wl['model']}.json").read_text())
    lo, hi = (torch.as_tensor(x, device=dev) for x in prior_box(spec, wl["width"]))
    tail = torch.as_tensor(np.concatenate([np.full(w.k, SIGMA_SHOCK), np.full(len(wl["meas"]), SIGMA_ERR)]), device=dev)[None]
    theta0 = torch.as_tensor(make_draws(spec, n_local, wl["width"], seed=0, skip=rank * n_local), device=dev)
    out = {"particles_per_gpu": n_local, "stages_timed": stages}
    for mode in ("rows", "allgather"):
        smc = TemperedSMC(w.ss, lo, hi, tail, w.Y_d, step_scale=0.02, seed=0, exchange=mode, profile=True).initialise(theta0)
        smc.stage(0.01, 1)  # warm-up
        smc.exchange_ms.clear()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(2, 2 + stages):
            smc.stage((s_ / (2 + stages)) ** 2, s_)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / stages, float(np.mean(smc.exchange_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d = theta0.shape[1]
        out[mode] = {"ms_per_stage": float(t[0]), "exchange_ms_per_stage": float(t[1]), "value": world * n_local / (float(t[0]) * 1e-3), "unit": UNIT,
                     "gathered_bytes_per_rank_per_stage": int(n_local * 8 * (3 + (d if mode == "allgather" else 0)))}
    return out


if __name__ == "__main__":
    main()
