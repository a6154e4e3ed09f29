"""A model spec compiled to a per-model CUDA library, and the batched likelihood pipeline built on it.

``CompiledModel``      theta -> A, B, C, D (generated kernel; replaces the compiled pytensor function of
                       gEconpy/model/model.py:1647-1664 / build.py:681-695).
``BatchedStateSpace``  the estimation graph of ``DSGEStateSpace`` (gEconpy/model/statespace.py:725-820, 1139-1215)
                       evaluated for a whole population of draws:  Jacobian -> cycle reduction -> R, residual ->
                       Blanchard-Kahn count -> P0 -> Kalman log-likelihood -> gating, four kernel launches per chunk.
"""

from __future__ import annotations

import ctypes as C
import os

from pathlib import Path

import numpy as np

from .. import _lib as L
from ..build import build_model
from .augmentation import StateAugmentation
from .codegen import LinearizedModel, load_spec

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

SPEC_DIR = Path(__file__).resolve().parent / "specs"


class CompiledModel:
    def __init__(self, spec, log_linearize: bool = True, not_loglin_variables=(), force_build: bool = False):
        if isinstance(spec, (str, Path)) and not Path(spec).exists():
            spec = SPEC_DIR / f"{spec}.json"
        self.lin = LinearizedModel(load_spec(spec), log_linearize=log_linearize, not_loglin_variables=tuple(not_loglin_variables))
        self.name = self.lin.name
        self.n, self.k, self.n_theta = self.lin.n, self.lin.k, self.lin.n_theta
        self.source = self.lin.cuda_source()
        self.lib_path = build_model(self.name, self.source, force=force_build)
        try:
            self._lib = C.CDLL(str(self.lib_path))
        except OSError as e:
            raise L.GeconLibraryError(f"cannot load {self.lib_path}: {e}") from e
        self._jac_dev = self._lib.gecon_model_jacobian_batched
        self._jac_dev.restype = C.c_int
        self._jac_dev.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
        self._jac_host = self._lib.gecon_model_jacobian_host
        self._jac_host.restype = C.c_int
        self._jac_host.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        self._vjp_dev = self._lib.gecon_model_vjp_batched
        self._vjp_dev.restype = C.c_int
        self._vjp_dev.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
        self._loglik = self._lib.gecon_model_loglik  # the fused theta -> log-likelihood entry point (gecon_pipeline_args)
        self._loglik.restype = C.c_int
        self._loglik.argtypes = [C.POINTER(L.PipelineArgs), C.c_void_p]
        # the model's own build of the warp-per-draw solver (csrc/cr_warp_spec.cu; same contract as gecon_cr_solve_batched, which it
        # calls itself for arguments it was not built for); models the warp kernel does not cover have no such symbol
        try:
            self._cr_solve = self._lib.gecon_model_cr_solve
            self._cr_solve.restype = C.c_int
            self._cr_solve.argtypes = [C.POINTER(L.CrArgs), C.c_void_p]
        except AttributeError:
            self._cr_solve = None
        self._jac_compact = self._lib.gecon_model_jacobian_compact
        self._jac_compact.restype = C.c_int
        self._jac_compact.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        n, k, nt = C.c_int32(), C.c_int32(), C.c_int32()
        self.launches = 0  # kernel launches made through this model library (bench.py's gpu_launches)
        self._lib.gecon_model_info(C.byref(n), C.byref(k), C.byref(nt))
        assert (n.value, k.value, nt.value) == (self.n, self.k, self.n_theta)

    # structure, in the reference's vocabulary
    var_names = property(lambda self: self.lin.var_names)
    shock_names = property(lambda self: self.lin.shock_names)
    param_names = property(lambda self: self.lin.param_names)
    var_order = property(lambda self: self.lin.var_order)
    eq_order = property(lambda self: self.lin.eq_order)
    inv_var_order = property(lambda self: self.lin.inv_var_order)
    permuted_lead_var_idx = property(lambda self: self.lin.permuted_lead_var_idx)
    col_ranges = property(lambda self: self.lin.col_ranges)

    def theta_vector(self, **updates) -> np.ndarray:
        d = dict(self.lin.defaults)
        unknown = set(updates) - set(d)
        if unknown:
            raise KeyError(f"unknown parameters {sorted(unknown)}")
        d.update(updates)
        return np.array([d[p] for p in self.param_names], dtype=np.float64)

    def jacobian_device(self, theta, A, B, Cm, D, xss, status, stream) -> None:
        """Launch on device pointers (torch tensors); nothing is allocated or copied."""
        N = theta.shape[0]
        rc = self._jac_dev(
            theta.data_ptr(), N, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), D.data_ptr(),
            xss.data_ptr() if xss is not None else None, status.data_ptr() if status is not None else None, C.c_void_p(stream),
        )  # fmt: skip
        self.launches += 1
        if rc != 0:
            raise L.GeconLibraryError(f"gecon_model_jacobian_batched({self.name}) failed with CUDA error {rc}")

    def structure(self):
        """``gecon_model_structure``: (nnz, table[nnz], off[5], col_ranges[4], lead_idx) of the generated library -- the layout of the
        compact Jacobian and the structural ranges, as numpy arrays (copied)."""
        nnz, nlead = C.c_int32(), C.c_int32()
        table, off, lead = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        rng = (C.c_int32 * 4)()
        self._lib.gecon_model_structure(C.byref(nnz), C.byref(table), C.byref(off), rng, C.byref(nlead), C.byref(lead))
        return (nnz.value, np.array([table[i] for i in range(nnz.value)], dtype=np.int32), np.array([off[i] for i in range(5)], dtype=np.int32),
                tuple(int(v) for v in rng), np.array([lead[i] for i in range(nlead.value)], dtype=np.int32))

    def jacobian_compact_device(self, theta, vals, status, stream, theta_stride=None) -> None:
        """theta [N, >= n_theta] (torch CUDA, rows may be wider than n_theta) -> vals [N, nnz]: structural non-zeros only."""
        rc = self._jac_compact(theta.data_ptr(), int(theta_stride or theta.shape[1]), theta.shape[0], vals.data_ptr(), None,
                               status.data_ptr() if status is not None else None, C.c_void_p(stream))
        self.launches += 1
        if rc != 0:
            raise L.GeconLibraryError(f"gecon_model_jacobian_compact({self.name}) failed with CUDA error {rc}")

    def vjp_device(self, theta, A_bar, B_bar, C_bar, D_bar, xss_bar, theta_bar, stream) -> None:
        """theta_bar[N, n_theta] = <(A_bar, B_bar, C_bar, D_bar[, xss_bar]), d(A, B, C, D[, x_ss])/dtheta> on device pointers: the
        generated reverse-mode kernel (codegen.vjp_body), last stage of the gradient path."""
        rc = self._vjp_dev(
            theta.data_ptr(), theta.shape[0], A_bar.data_ptr(), B_bar.data_ptr(), C_bar.data_ptr(), D_bar.data_ptr(),
            xss_bar.data_ptr() if xss_bar is not None else None, theta_bar.data_ptr(), C.c_void_p(stream),
        )  # fmt: skip
        self.launches += 1
        if rc != 0:
            raise L.GeconLibraryError(f"gecon_model_vjp_batched({self.name}) failed with CUDA error {rc}")

    def jacobian(self, theta):
        """theta[N, n_theta] (numpy, host path) -> A, B, C, D, xss, status in the reference's permuted solver order
        (rows eq_order, columns var_order; gEconpy/model/perturbation.py:130-158)."""
        L.require_device()
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        if th.shape[1] != self.n_theta:
            raise ValueError(f"theta has {th.shape[1]} columns, model {self.name} has {self.n_theta} free parameters")
        N, n, k = th.shape[0], self.n, self.k
        A, B, Cm = (np.empty((N, n, n)) for _ in range(3))
        D = np.empty((N, n, k))
        xss = np.empty((N, n))
        st = np.empty(N, dtype=np.int32)
        rc = self._jac_host(th.ctypes.data, N, A.ctypes.data, B.ctypes.data, Cm.ctypes.data, D.ctypes.data, xss.ctypes.data, st.ctypes.data)
        if rc != 0:
            raise L.GeconLibraryError(f"gecon_model_jacobian_host({self.name}) failed with CUDA error {rc}")
        return A, B, Cm, D, xss, st


# draws whose Jacobian is NaN never reach the BK kernel: flag them as BK failures here so that the gate sees them
GATE_MASK = L.ST_CR_NOT_CONVERGED | L.ST_CR_NAN | L.ST_SINGULAR | L.ST_RESID | L.ST_BK | L.ST_BK_INCONCLUSIVE | L.ST_JAC_NONFINITE


class BatchedStateSpace:
    """Population-batched equivalent of ``DSGEStateSpace.configure`` + ``build_statespace_graph`` + compiled logp.

    Parameter vector per draw (the reference's parameter names, statespace.py:1093-1137):
    ``[free parameters..., sigma_<shock>..., error_sigma_<state>...]``.
    """

    def __init__(self, model: CompiledModel):
        self.model = model
        self.configured = False
        self._grad_ws = None

    def configure(
        self,
        observed_states,
        measurement_error=None,
        solver: str = "cycle_reduction",
        tol: float = 1e-8,
        max_iter: int = 1000,
        solver_tol: float = 1e-8,
        cov_jitter: float = 1e-8,
        missing_fill_value: float = -9999.0,
        mvn_const: str = "per_obs",
        check_bk: bool = True,
        bk_on_rejected_draws: bool = True,
        chunk: int = 65536,
        reduce_state: bool = True,
        n_streams: int | None = None,
        temporal_aggregation: dict | None = None,
        aggregation_period: int = 4,
        ss_obs_intercept: list | None = None,
        observation_equations: dict | None = None,
        full_shock_covariance: bool = False,
        mask_intercept: bool = False,
        constant_params=None,
        fused: bool | None = None,
        mode=None,
        verbose: bool = True,
        use_adjoint_gradients: bool = True,
        use_direct_lyapunov: bool = False,
        add_bk_check: bool | None = None,
        add_solver_success_check: bool = True,
        specialize: bool | None = None,
    ):
        """Same meaning as ``DSGEStateSpace.configure`` (gEconpy/model/statespace.py:822-1090) for the arguments it
        shares: ``temporal_aggregation`` {"sum" | "mean" | "first" | "last"} with ``aggregation_period`` adds cumulator
        states (statespace.py:598-650), ``ss_obs_intercept`` puts log x_ss(theta) / x_ss(theta) of the listed observed
        states into the observation intercept d (statespace.py:363-388), ``observation_equations`` {observed series: GCN-syntax
        expression in model variables (``v[]``, ``v[-1]``, ``v[ss]``) and parameters} is linearised around the steady state into a
        parameter-dependent design-matrix row, an intercept and observation-lag states (statespace.py:390-556,652-694);
        ``full_shock_covariance`` replaces the k ``sigma_<shock>`` entries of the parameter vector by the k x k matrix
        ``state_cov`` (row-major, used as Q directly: statespace.py:245-249).  ``mask_intercept``: False = the observation
        intercept is not masked at missing entries (SURVEY A.5's restatement of the upstream filter: a missing entry then
        contributes -(d_i^2 / jitter + log jitter) / 2, which matters as soon as ``ss_obs_intercept`` meets missing data);
        True = missing entries contribute nothing.  ``tests/golden/make_kalman_goldens.py`` records which one the installed
        pymc_extras follows.

        ``solver`` (statespace.py:197-222): ``"cycle_reduction"``; ``"gensys"`` (the same kernel: T from cycle reduction,
        determinacy from the Blanchard-Kahn count); ``"scan_cycle_reduction"`` (the scan twin's conventions:
        only ||A0||_1 is tested and T is always solved for; the per-draw step count is what the reference publishes as
        ``n_cycle_steps``, here ``out_n_iter``); ``"backward_direct"`` (models without leads: T = -B^-1 A, R = -B^-1 D, no
        iteration, no Blanchard-Kahn check).  ``constant_params`` (list, or ``"auto"`` = every parameter without a prior):
        held at their defaults and dropped from the parameter vector (statespace.py:741-753).  ``mode`` (a pytensor
        compilation mode), ``use_adjoint_gradients`` (the gradient path is always adjoint-based) and ``use_direct_lyapunov``
        (P0 comes from Smith doubling, which agrees with both of the reference's Lyapunov solvers to rounding) are accepted for
        signature compatibility.  ``add_bk_check`` / ``add_solver_success_check``: the two ``pm.Potential(-inf)`` gates of
        ``build_statespace_graph`` (statespace.py:1206-1215).  The reference defaults both to False; here both default to
        True because a population of prior draws always contains draws the solver rejects -- pass False, False to reproduce
        the reference's default graph (non-finite steady states and failed solves are gated either way: their logp is NaN in
        the reference, -inf here).  ``check_bk`` is the older name of ``add_bk_check``.  ``bk_on_rejected_draws=False`` (fused path only):
        draws the gate already rejects for another reason are not Blanchard-Kahn-counted -- same log-likelihoods, their BK status bit
        stays unset; on a population where half the draws fail the count is most of the step, and a sampler only consumes the gate.
        ``specialize``: the filter kernel compiled for this configuration's (filter dimension, observables) pair (``build.build_filter_spec``:
        3-7 s of nvcc once, cached on disk; same source as the generic kernel, identical results, ~12 % faster) -- True: build it now;
        False: never use it; None (default): use it when it is already on disk, and build it the first time a population of at least
        ``SPEC_MIN_DRAWS`` draws arrives.  (The solver's counterpart needs no switch: it is part of the model library.)"""
        m = self.model
        if solver not in ("cycle_reduction", "gensys", "scan_cycle_reduction", "backward_direct"):
            raise NotImplementedError(f"solver={solver!r}: expected cycle_reduction, gensys, scan_cycle_reduction or backward_direct")
        if solver == "backward_direct" and len(m.permuted_lead_var_idx):
            raise ValueError("solver='backward_direct' needs a model without forward-looking variables (backward_looking.py:8-60)")
        self.solver = solver
        if add_bk_check is not None:
            check_bk = add_bk_check
        if solver == "backward_direct":
            check_bk = False
        self.add_solver_success_check = bool(add_solver_success_check)
        self.mode, self.verbose = mode, verbose
        self.use_adjoint_gradients, self.use_direct_lyapunov = bool(use_adjoint_gradients), bool(use_direct_lyapunov)
        observation_equations = dict(observation_equations or {})
        unknown_keys = [k_ for k_ in observation_equations if k_ not in observed_states]
        if unknown_keys:
            raise ValueError(f"The following observation_equations entries are not in observed_states: {', '.join(unknown_keys)}")
        overlap = set(observation_equations) & set(ss_obs_intercept or [])
        if overlap:
            raise ValueError(
                f"The following observed states appear in both observation_equations and ss_obs_intercept: {', '.join(sorted(overlap))}. "
                "An observation equation already determines its intercept; remove these names from one or the other."
            )
        unknown = [v for v in observed_states if v not in m.var_names and v not in observation_equations]
        if unknown:
            raise ValueError(f"unknown observed states {unknown}")
        # parse + linearise now so that errors surface at configure time (statespace.py:1027-1035)
        self._obs_eq = {}
        for name_, expr_ in observation_equations.items():
            sym_, refs_ = m.lin.parse_observation_equation(name_, expr_)
            self._obs_eq[name_] = m.lin.linearize_observation_equation(sym_, refs_)
        model_observed = [v for v in observed_states if v not in observation_equations]
        measurement_error = list(measurement_error or [])
        bad = [v for v in measurement_error if v not in observed_states]
        if bad:
            raise ValueError(f"measurement error on unobserved states {bad}")
        # stochastic-singularity guard (statespace.py:994-1005)
        if len(observed_states) > m.k + len(measurement_error):
            raise ValueError(
                f"stochastic singularity: {len(observed_states)} observed states but only {m.k} shocks + "
                f"{len(measurement_error)} measurement errors"
            )
        self.observed_states = list(observed_states)
        self.measurement_error = measurement_error
        self.p = len(observed_states)
        # filter runs in solver order: observed variable -> permuted position (T, R are not un-permuted in between)
        self.obs_idx = m.inv_var_order[[m.var_names.index(v) for v in model_observed]].astype(np.int32)
        # the reference puts the error variances at positions 0..len(error_states)-1 of diag(H), whatever the position of
        # those states among the observed ones (statespace.py:800-808, "mirror the previous semantics"): reproduced
        self.err_pos = np.arange(len(measurement_error), dtype=np.int64)
        # The likelihood only depends on the lagged (state) variables and the observed variables: every other column of
        # T is identically zero, so those variables never feed back into the recursion.  With reduce_state the solver
        # kernel hands the filter the exact sub-blocks T[U][:, U], R[U] for U = states + observed (solver order).
        state_pos = m.inv_var_order[m.lin.state_var_idx]
        referenced = {int(m.inv_var_order[m.var_names.index(v)]) for _, co in self._obs_eq.values() for (v, _lag) in co}
        # Order: the lagged (state) variables first, then the observed / referenced ones that are not states.  T = -A1hat^-1 A has
        # non-zero columns only at the lagged variables, so in this order only the first len(states) columns of the filter's T can be
        # non-zero: gecon_kalman_args.t_cols (large NK: 16 of 19 columns -> 4 instead of 5 k-steps in both products with T).
        states = sorted(set(state_pos.tolist()))
        others = sorted((set(self.obs_idx.tolist()) | referenced) - set(states))
        self.filter_vars = np.array(states + others, dtype=np.int32) if reduce_state else np.arange(m.n, dtype=np.int32)
        self.filter_t_cols = len(states) if (reduce_state and others and states) else 0
        self.n_filter = int(self.filter_vars.size)
        self.obs_idx_filter = np.array([int(np.flatnonzero(self.filter_vars == o)[0]) for o in self.obs_idx], dtype=np.int32)
        self.reduce_state = bool(reduce_state)
        self.n_streams = int(n_streams if n_streams is not None else os.environ.get("GECON_STREAMS", "1"))
        # ---- state augmentation (cumulators) and the observation intercept
        ss_obs_intercept = list(ss_obs_intercept or [])
        unknown = [v for v in ss_obs_intercept if v not in observed_states]
        if unknown:
            raise ValueError(f"The following ss_obs_intercept entries are not in observed_states: {', '.join(unknown)}")
        depths = StateAugmentation.required_obs_lag_depths(
            {k_: co.keys() for k_, (_ic, co) in self._obs_eq.items()}, temporal_aggregation, int(aggregation_period)
        )
        self.aug = StateAugmentation(
            [m.lin.vars_perm[int(u)] for u in self.filter_vars], list(observed_states), dict(temporal_aggregation or {}), int(aggregation_period),
            obs_equation_names=tuple(self._obs_eq), obs_lag_depths=depths,
        )  # fmt: skip
        self.n_aug = self.aug.k_states
        if self.n_aug > 64:
            raise NotImplementedError(f"augmented state dimension {self.n_aug} > 64")
        self.dense_Z = None if self.aug.is_selector() else np.ascontiguousarray(self.aug.design_matrix())
        self.ss_obs_intercept = ss_obs_intercept
        self._obs_lib = None
        if self._obs_eq:  # per-draw cells of Z and d: one generated kernel per configuration
            z_cells, d_cells = {}, {}
            for i_, name_ in enumerate(observed_states):
                if name_ not in self._obs_eq:
                    continue
                icpt, coeffs = self._obs_eq[name_]
                for col, terms in self.aug.design_cells(name_, coeffs).items():
                    z_cells[i_ * self.n_aug + col] = sum(w_ * coeffs[key_] for w_, key_ in terms)
                d_cells[i_] = icpt * (int(aggregation_period) if (temporal_aggregation or {}).get(name_) == "sum" else 1)
            tag = "obs_" + "_".join(sorted(self._obs_eq))
            src = m.lin.obs_source(z_cells, d_cells, tag)
            lib_path = build_model(f"{m.name}_{tag}"[:80], src)
            try:
                self._obs_lib = C.CDLL(str(lib_path))
            except OSError as e:
                raise L.GeconLibraryError(f"cannot load {lib_path}: {e}") from e
            self._obs_lib.gecon_obs_batched.restype = C.c_int
            self._obs_lib.gecon_obs_batched.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
            self._obs_lib.gecon_obs_vjp_batched.restype = C.c_int
            self._obs_lib.gecon_obs_vjp_batched.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        loglin = set(m.var_names) - set(m.lin.not_loglin_variables) if m.lin.log_linearize else set()
        self._d_pos = np.array([observed_states.index(v) for v in ss_obs_intercept], dtype=np.int64)
        self._d_var = np.array([m.var_names.index(v) for v in ss_obs_intercept], dtype=np.int64)
        self._d_loglin = np.array([v in loglin for v in ss_obs_intercept], dtype=bool)
        self._d_scale = self.aug.intercept_scale()[self._d_pos] if ss_obs_intercept else np.zeros(0)
        self.tol, self.max_iter, self.solver_tol = float(tol), int(max_iter), float(solver_tol)
        self.cov_jitter, self.missing_fill_value, self.mvn_const = float(cov_jitter), float(missing_fill_value), mvn_const
        self.check_bk = bool(check_bk)
        self.bk_on_rejected_draws = bool(bk_on_rejected_draws)
        self.specialize = specialize
        self.__dict__.pop("_kf_spec_cache", None)
        self.__dict__.pop("_kf_spec_paths", None)
        self.chunk = int(os.environ.get("GECON_CHUNK", chunk))
        self.full_covariance = bool(full_shock_covariance)
        self.mask_intercept = bool(mask_intercept)
        self._n_cov = m.k * m.k if self.full_covariance else m.k
        cov_names = [f"state_cov[{i},{j}]" for i in range(m.k) for j in range(m.k)] if self.full_covariance else [f"sigma_{s}" for s in m.shock_names]
        all_names = list(m.param_names) + cov_names + [f"error_sigma_{v}" for v in measurement_error]
        # constant parameters (statespace.py:741-753): held at the spec's defaults, not part of the parameter vector
        if constant_params == "auto":
            constant_params = [p_ for p_ in m.param_names if p_ not in m.lin.spec.get("bounds", {})]
        constant_params = list(constant_params or [])
        unknown = [p_ for p_ in constant_params if p_ not in m.param_names]
        if unknown:
            raise ValueError(f"unknown constant_params {unknown}; model parameters: {list(m.param_names)}")
        self.constant_params = constant_params
        self._n_param_full = len(all_names)
        self._free_cols = np.array([i for i, nm in enumerate(all_names) if nm not in constant_params], dtype=np.int64)
        self._const_row = np.zeros(self._n_param_full)
        self._const_row[: m.n_theta] = m.theta_vector()
        self.param_names = [all_names[i] for i in self._free_cols]
        self.n_param = len(self.param_names)
        # gate: the reference's two Potentials + what makes its logp NaN anyway
        self.gate_mask = L.ST_JAC_NONFINITE | L.ST_CR_NAN | L.ST_SINGULAR
        if check_bk:
            self.gate_mask |= L.ST_BK | L.ST_BK_INCONCLUSIVE
        if self.add_solver_success_check:
            self.gate_mask |= L.ST_RESID | L.ST_CR_NOT_CONVERGED
        # the fused C entry point (gecon_model_loglik) covers plain state spaces; everything else runs the Python pipeline
        eligible = (self.dense_Z is None and self._obs_lib is None and not self.ss_obs_intercept and not self.full_covariance
                    and self.n_aug == self.n_filter and self.solver != "backward_direct" and self.n_streams == 1)  # fmt: skip
        if fused and not eligible:
            raise ValueError("fused=True needs a plain state space: selector observations, diagonal shock covariance, no state "
                             "augmentation / observation equations / steady-state intercept, a forward-looking solver")
        self.fused = eligible if fused is None else bool(fused)
        self.fused = self.fused and os.environ.get("GECON_FUSED", "1") != "0"
        self.configured = True
        self._ws = None
        self._ws_extra = {}  # per-stream workspaces of the multi-stream pipeline: shapes depend on the configuration
        self._grad_ws = None
        return self

    # ------------------------------------------------------------------------------------------------ workspace
    def _workspace(self, device, nc, slot: int = 0):
        if slot:  # extra workspaces of the multi-stream pipeline
            cache = self.__dict__.setdefault("_ws_extra", {})
            main, self._ws = self._ws, cache.get(slot)
            try:
                cache[slot] = self._workspace(device, nc)
            finally:
                self._ws = main
            return cache[slot]
        if self._ws is not None and self._ws["nc"] >= nc and self._ws["device"] == device:
            return self._ws
        m = self.model
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        ws = dict(
            nc=nc, device=device,
            theta=torch.empty((nc, m.n_theta), **f64), sig=torch.empty((nc, m.k), **f64), herr=torch.zeros((nc, self.p), **f64),
            A=torch.empty((nc, m.n, m.n), **f64), B=torch.empty((nc, m.n, m.n), **f64), C=torch.empty((nc, m.n, m.n), **f64),
            D=torch.empty((nc, m.n, m.k), **f64), T=torch.zeros((nc, self.n_aug, self.n_aug), **f64),
            R=torch.zeros((nc, self.n_aug, m.k), **f64), subset=torch.as_tensor(self.filter_vars, **i32),
            status=torch.empty((nc,), **i32), n_iter=torch.empty((nc,), **i32), n_unstable=torch.empty((nc,), **i32),
            resid=torch.empty((nc,), **f64),
            lead=torch.as_tensor(m.permuted_lead_var_idx, **i32), obs=torch.as_tensor(self.obs_idx_filter, **i32),
            err_pos=torch.as_tensor(self.err_pos, device=device),
        )  # fmt: skip
        if self.n_aug > self.n_filter:
            # constant rows [F | kron(I, shift)] of the augmented transition: written once, the solver kernel only ever
            # writes the top-left n_filter x n_filter block (t_ld / t_stride) and the top n_filter rows of R
            ws["T"][:, self.n_filter :, :] = torch.as_tensor(self.aug.transition_rows(), **f64)
        if self.full_covariance:
            ws["Q"] = torch.empty((nc, m.k, m.k), **f64)
        if self.dense_Z is not None:
            ws["Z"] = torch.as_tensor(self.dense_Z, **f64)
            if self._obs_lib is not None:  # one design matrix per draw: constant rows now, equation rows by the obs kernel
                ws["Z"] = ws["Z"].unsqueeze(0).repeat(nc, 1, 1).contiguous()
        if self._obs_lib is not None or self.ss_obs_intercept:
            ws["d"] = torch.zeros((nc, self.p), **f64)
        if self.ss_obs_intercept:
            ws["xss"] = torch.empty((nc, m.n), **f64)
            ws["d_var"] = torch.as_tensor(self._d_var, device=device)
            ws["d_pos"] = torch.as_tensor(self._d_pos, device=device)
            ws["d_loglin"] = torch.as_tensor(self._d_loglin, device=device)
            ws["d_scale"] = torch.as_tensor(self._d_scale, **f64)
        self._ws = ws
        return ws

    def _expand_params(self, theta):
        """[N, n_param] -> [N, n_param_full]: constant parameters filled in at their defaults (no-op without any)."""
        if theta.shape[1] != self.n_param:
            raise ValueError(f"theta has {theta.shape[1]} columns, expected {self.n_param}: {self.param_names}")
        if not self.constant_params:
            return theta
        full = torch.as_tensor(self._const_row, dtype=torch.float64, device=theta.device).repeat(theta.shape[0], 1)
        full[:, torch.as_tensor(self._free_cols, device=theta.device)] = theta
        return full

    # ------------------------------------------------------------------------------------------------ evaluation
    def loglik_device(self, theta_full, Y, out_ll=None, out_status=None, out_n_iter=None, events=None):
        """theta_full: torch CUDA tensor [N, n_param]; Y: torch CUDA tensor [Tobs, p].  Returns (ll, status) on device.
        Four kernel launches (+4 memsets) per chunk of draws, all on the current stream; no host synchronisation."""
        if not self.configured:
            raise RuntimeError("call configure(...) first")
        if not (torch is not None and isinstance(theta_full, torch.Tensor) and theta_full.is_cuda):
            raise TypeError("loglik_device needs CUDA tensors; use loglik() for host arrays")
        m = self.model
        lib = L.load_library()
        dev = theta_full.device
        N = theta_full.shape[0]
        theta_full = self._expand_params(theta_full)
        Y = Y.to(torch.float64).contiguous().reshape(-1, self.p)
        Tobs = Y.shape[0]
        ll = out_ll if out_ll is not None else torch.empty((N,), dtype=torch.float64, device=dev)
        status = out_status if out_status is not None else torch.empty((N,), dtype=torch.int32, device=dev)
        if self.fused and max(1, int(getattr(self, "n_streams", 1))) == 1:
            return self._loglik_fused(theta_full, Y, ll, status, out_n_iter, events)
        nc = min(self.chunk, N)
        n_err = len(self.measurement_error)
        # (the staged path filters the augmented state at its full dimension; the augmentation adds rows AND columns to T, so the
        # promise about its zero columns is only made for the plain state space)
        t_cols = self.filter_t_cols if self.n_aug == self.n_filter else 0
        kf_spec = self._filter_spec_fn(self.n_aug, N, t_cols)
        n_streams = max(1, int(getattr(self, "n_streams", 1)))
        cur = torch.cuda.current_stream(dev)
        if n_streams > 1:
            # chunks alternate between streams (each with its own workspace) so that the solver kernel of one chunk and the
            # filter kernel of another share the SMs; GECON_CR_CTAS_PER_SM / GECON_KF_CTAS_PER_SM size their grids for that
            pool = self.__dict__.setdefault("_streams", {})
            streams = pool.setdefault(dev, [torch.cuda.Stream(dev) for _ in range(n_streams)])
            start = torch.cuda.Event()
            start.record(cur)
            for s_ in streams:
                s_.wait_event(start)
        else:
            streams = [cur]

        def mark(name):
            """bench.py hook: CUDA events on the launching stream around each kernel (per-kernel roofline)."""
            if events is None:
                return None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            events.append((name, e0, e1))
            return e1

        try:  # the current torch stream is restored even if a launch raises
            for ci, lo in enumerate(range(0, N, nc)):
                cnt = min(nc, N - lo)
                th = theta_full[lo : lo + cnt]
                slot = ci % len(streams)
                ws = self._workspace(dev, nc, slot)
                torch.cuda.set_stream(streams[slot])
                stream = streams[slot].cuda_stream
                # split the parameter vector (strided device copies; no arithmetic)
                ws["theta"][:cnt].copy_(th[:, : m.n_theta])
                if self.full_covariance:
                    ws["Q"][:cnt].copy_(th[:, m.n_theta : m.n_theta + self._n_cov].reshape(cnt, m.k, m.k))
                else:
                    ws["sig"][:cnt].copy_(th[:, m.n_theta : m.n_theta + m.k])
                if n_err:
                    ws["herr"][:cnt].index_copy_(1, ws["err_pos"], th[:, m.n_theta + self._n_cov :])
                st = ws["status"][:cnt]
                e = mark("jacobian")
                m.jacobian_device(ws["theta"][:cnt], ws["A"], ws["B"], ws["C"], ws["D"], ws.get("xss"), st, stream)
                e and e.record()
                if self.ss_obs_intercept:  # d = log x_ss / x_ss of the listed observed states (x aggregation period for "sum")
                    xs = ws["xss"][:cnt].index_select(1, ws["d_var"])
                    ws["d"][:cnt].index_copy_(1, ws["d_pos"], torch.where(ws["d_loglin"], xs.log(), xs) * ws["d_scale"])
                if self._obs_lib is not None:
                    rc = self._obs_lib.gecon_obs_batched(ws["theta"].data_ptr(), cnt, ws["Z"].data_ptr(), self.p * self.n_aug,
                                                         ws["d"].data_ptr(), self.p, C.c_void_p(stream))  # fmt: skip
                    m.launches += 1
                    if rc != 0:
                        raise L.GeconLibraryError(f"gecon_obs_batched failed with CUDA error {rc}")
                cr = L.CrArgs(
                    struct_size=C.sizeof(L.CrArgs), A=ws["A"].data_ptr(), B=ws["B"].data_ptr(),
                    C=(None if self.solver == "backward_direct" else ws["C"].data_ptr()), scan_semantics=int(self.solver == "scan_cycle_reduction"),
                    D=ws["D"].data_ptr(), N=cnt, n=m.n, k=m.k, max_iter=self.max_iter, accumulate=1, tol=self.tol,
                    resid_tol=self.solver_tol, unperm=ws["subset"].data_ptr(), T=ws["T"].data_ptr(), R=ws["R"].data_ptr(),
                    status=st.data_ptr(), n_iter=ws["n_iter"].data_ptr(), resid=ws["resid"].data_ptr(), norms=None,
                    n_out=self.n_filter, n_lead=(int(ws["lead"].numel()) if self.check_bk else 0),
                    lead_idx=(ws["lead"].data_ptr() if self.check_bk else None), n_unstable=ws["n_unstable"].data_ptr(),
                    t_stride=self.n_aug * self.n_aug, r_stride=self.n_aug * m.k, t_ld=self.n_aug,
                    lag_lo=m.col_ranges[0], lag_hi=m.col_ranges[1], lead_lo=m.col_ranges[2], lead_hi=m.col_ranges[3],
                )  # fmt: skip
                e = mark("cr_solve")
                L.check((m._cr_solve or lib.gecon_cr_solve_batched)(C.byref(cr), C.c_void_p(stream)), "gecon_cr_solve_batched")
                e and e.record()
                if self.check_bk:
                    bk = L.BkArgs(
                        struct_size=C.sizeof(L.BkArgs), A=ws["A"].data_ptr(), B=ws["B"].data_ptr(), C=ws["C"].data_ptr(), N=cnt,
                        n=m.n, n_lead=int(ws["lead"].numel()), lead_idx=ws["lead"].data_ptr(), accumulate=1, max_iter=0,
                        n_unstable=ws["n_unstable"].data_ptr(), status=st.data_ptr(),
                        skip_mask=L.ST_BK_CERTIFIED | L.ST_JAC_NONFINITE,
                    )  # fmt: skip
                    e = mark("bk_count")
                    L.check(lib.gecon_bk_count_batched(C.byref(bk), C.c_void_p(stream)), "gecon_bk_count_batched")
                    e and e.record()
                kf = L.KalmanArgs(
                    struct_size=C.sizeof(L.KalmanArgs), T=ws["T"].data_ptr(), R=ws["R"].data_ptr(),
                    qdiag=(None if self.full_covariance else ws["sig"].data_ptr()), q_stride=m.k,
                    qfull=(ws["Q"].data_ptr() if self.full_covariance else None), qfull_stride=m.k * m.k,
                    hdiag=ws["herr"].data_ptr() if n_err else None, h_stride=self.p,
                    Z=(ws["Z"].data_ptr() if self.dense_Z is not None else None),
                    z_stride=(self.p * self.n_aug if self._obs_lib is not None else 0),
                    obs_idx=(ws["obs"].data_ptr() if self.dense_Z is None else None),
                    d=(ws["d"].data_ptr() if "d" in ws else None), d_stride=(self.p if "d" in ws else 0),
                    Y=Y.data_ptr(), P0=None, N=cnt, n=self.n_aug, k=m.k, p=self.p,
                    Tobs=Tobs, jitter=self.cov_jitter, missing_fill=self.missing_fill_value,
                    mvn_const_mode=(0 if self.mvn_const == "per_obs" else 1), lyap_max_iter=0, status_in=st.data_ptr(),
                    gate_mask=self.gate_mask, sigma_inputs=1, ll=ll[lo : lo + cnt].data_ptr(), status=status[lo : lo + cnt].data_ptr(),
                    ll_t=None, mask_intercept=int(self.mask_intercept), t_cols=t_cols,
                )  # fmt: skip
                e = mark("kalman_ll")
                L.check((kf_spec or lib.gecon_kalman_ll_batched)(C.byref(kf), C.c_void_p(stream)), "gecon_kalman_ll_batched")
                e and e.record()
                if out_n_iter is not None:
                    out_n_iter[lo : lo + cnt].copy_(ws["n_iter"][:cnt])
        finally:
            torch.cuda.set_stream(cur)
        if n_streams > 1:
            for s_ in streams:
                done = torch.cuda.Event()
                done.record(s_)
                cur.wait_event(done)
        return ll, status

    def _loglik_fused(self, theta_full, Y, ll, status, out_n_iter, events):
        """One call of the generated library's ``gecon_model_loglik`` (include/gecon_b200.h, gecon_pipeline_args): the whole chunk
        loop runs in C on the current stream -- compact Jacobian, solver, Blanchard-Kahn count, filter -- with no torch kernel
        in between.  ``events`` (bench.py's per-kernel timing hook) turns on the pipeline's own CUDA-event timing."""
        m = self.model
        dev = theta_full.device
        theta_full = theta_full.contiguous()
        fv = np.ascontiguousarray(self.filter_vars, dtype=np.int32)
        obs = np.ascontiguousarray(self.obs_idx_filter, dtype=np.int32)
        lead = np.ascontiguousarray(m.permuted_lead_var_idx, dtype=np.int32)
        args = L.PipelineArgs(
            struct_size=C.sizeof(L.PipelineArgs), n_err=len(self.measurement_error), p=self.p, n_filter=self.n_filter,
            n_lead=(int(lead.size) if self.check_bk else 0), filter_vars=fv.ctypes.data, obs_idx=obs.ctypes.data,
            lead_idx=(lead.ctypes.data if self.check_bk and lead.size else None), theta=theta_full.data_ptr(),
            theta_stride=theta_full.shape[1], N=theta_full.shape[0], Y=Y.data_ptr(), Tobs=Y.shape[0], max_iter=self.max_iter, tol=self.tol,
            solver_tol=self.solver_tol, jitter=self.cov_jitter, missing_fill=self.missing_fill_value,
            mvn_const_mode=(0 if self.mvn_const == "per_obs" else 1), mask_intercept=int(self.mask_intercept), gate_mask=self.gate_mask,
            check_bk=(int(self.check_bk and lead.size > 0) * (1 if self.bk_on_rejected_draws else 2)), scan_semantics=int(self.solver == "scan_cycle_reduction"),
            timing=int(events is not None), chunk=self.chunk, ll=ll.data_ptr(), status=status.data_ptr(),
            n_iter=(out_n_iter.data_ptr() if out_n_iter is not None else None),
        )  # fmt: skip
        kf_spec = self._filter_spec_fn(self.n_filter, theta_full.shape[0], self.filter_t_cols)
        if kf_spec is not None:
            args.kalman_ll = C.cast(kf_spec, C.c_void_p)
        stream = torch.cuda.current_stream(dev).cuda_stream
        before = L.load_library().gecon_launch_count()
        rc = m._loglik(C.byref(args), C.c_void_p(stream))
        if rc != 0:
            L.check(rc, f"gecon_model_loglik({m.name})")
        if events is not None:
            ms = (C.c_float * 4)()
            L.load_library().gecon_pipeline_stage_ms(ms)
            events.append(("__fused_ms__", dict(zip(("jacobian", "cr_solve", "bk_count", "kalman_ll"), (float(v) for v in ms))), None))
        self.fused_launches = int(L.load_library().gecon_launch_count() - before)
        return ll, status

    SPEC_MIN_DRAWS = 4096  # populations this large pay for the per-configuration build of the filter within their first evaluation

    def _filter_spec_fn(self, n_filter: int, n_draws: int, t_cols: int = 0):
        """``gecon_kalman_ll_spec`` of the filter library built for (n_filter, shocks, observables) -- ``build.build_filter_spec`` -- or
        None: configure(specialize=False), ``GECON_KF_SPEC=0``, a configuration the warp-per-draw filter does not take (dense design
        matrix, full shock covariance, thread-per-draw or CTA-per-draw sizes), or not built yet and the population is small.  The entry
        point has the contract of ``gecon_kalman_ll_batched`` and calls it itself for arguments it was not built for."""
        if self.specialize is False or os.environ.get("GECON_KF_SPEC", "1") == "0" or self.dense_Z is not None or self.full_covariance:
            return None
        from .. import build

        key = (int(n_filter), self.model.k, self.p, int(t_cols))
        cache = self.__dict__.setdefault("_kf_spec_cache", {})
        if key not in cache:
            paths = self.__dict__.setdefault("_kf_spec_paths", {})  # (the digest in the file name reads the sources: once per key)
            if key not in paths:
                paths[key] = build.filter_spec_path(*key)
            path = paths[key]
            if path is None:
                cache[key] = None
            elif path.exists() or self.specialize or n_draws >= self.SPEC_MIN_DRAWS:
                L.load_library()
                lib = C.CDLL(str(build.build_filter_spec(*key)))
                lib.gecon_kalman_ll_spec.restype = C.c_int
                lib.gecon_kalman_ll_spec.argtypes = [C.POINTER(L.KalmanArgs), C.c_void_p]
                cache[key] = lib
            else:
                return None  # not cached: a later, larger population may still build it
        lib = cache[key]
        return None if lib is None else lib.gecon_kalman_ll_spec

    # ------------------------------------------------------------------------------------------------ gradient
    def _grad_spec_lib(self):
        """The Kalman adjoint kernel built for this configuration's (filter dimension, shocks, observables) -- ``build.build_grad_spec``,
        cached on disk like the generated model kernels -- or None when ``GECON_GRAD_SPEC=0`` (then the generic kernel of the core
        library runs: same source, run-time dimensions)."""
        import os

        if os.environ.get("GECON_GRAD_SPEC", "1") == "0":
            return None
        key = (self.n_aug, self.model.k, self.p)
        cache = self.__dict__.setdefault("_grad_spec_cache", {})
        if key not in cache:
            from .. import build

            L.load_library()
            lib = C.CDLL(str(build.build_grad_spec(*key)))
            lib.gecon_kalman_grad_spec.restype = C.c_int
            lib.gecon_kalman_grad_spec.argtypes = [C.POINTER(L.KalmanGradArgs), C.c_void_p]
            cache[key] = lib
        return cache[key]

    def loglik_and_grad_device(self, theta_full, Y, events=None):
        """theta_full [N, n_param] (CUDA) -> (ll [N], grad [N, n_param], status [N]): the log-likelihood and its gradient
        with respect to every entry of the parameter vector (free parameters, sigma_<shock>, error_sigma_<state>) --
        what ``pm.Model.dlogp`` delivers to NUTS in the reference, as seven launches per chunk:
        Jacobian -> cycle reduction (full T, R) -> BK count -> Kalman forward + reverse sweep -> policy / selection
        adjoint -> generated vector-Jacobian product.  Gated draws (ll = -inf) get a zero gradient."""
        if not self.configured:
            raise RuntimeError("call configure(...) first")
        if not (torch is not None and isinstance(theta_full, torch.Tensor) and theta_full.is_cuda):
            raise TypeError("loglik_and_grad_device needs CUDA tensors; use loglik_and_grad() for host arrays")
        m = self.model
        lib = L.load_library()
        dev = theta_full.device
        N = theta_full.shape[0]
        theta_full = self._expand_params(theta_full)
        Y = Y.to(torch.float64).contiguous().reshape(-1, self.p)
        Tobs = Y.shape[0]
        f64 = dict(dtype=torch.float64, device=dev)
        ll = torch.empty((N,), **f64)
        status = torch.empty((N,), dtype=torch.int32, device=dev)
        grad = torch.zeros((N, self._n_param_full), **f64)
        nc = min(self.chunk, N)  # (full-size adjoints per draw: ~50 KB per draw at n = 24, 3 GB per 65,536-draw chunk)
        ws = self._workspace(dev, nc)
        g = self._grad_ws
        if g is None or g["nc"] < nc or g["device"] != dev:
            na, n, k, p = self.n_aug, m.n, m.k, self.p
            g = self._grad_ws = dict(
                nc=nc, device=dev, Tfull=torch.empty((nc, n, n), **f64), Rfull=torch.empty((nc, n, k), **f64),
                Tb_f=torch.empty((nc, na, na), **f64), Rb_f=torch.empty((nc, na, k), **f64), qb=torch.empty((nc, self._n_cov), **f64),
                hb=torch.empty((nc, p), **f64), db=torch.empty((nc, p), **f64), Tb=torch.zeros((nc, n, n), **f64),
                Rb=torch.zeros((nc, n, k), **f64), Ab=torch.empty((nc, n, n), **f64), Bb=torch.empty((nc, n, n), **f64),
                Cb=torch.empty((nc, n, n), **f64), Db=torch.empty((nc, n, k), **f64), thb=torch.empty((nc, m.n_theta), **f64),
                xssb=torch.zeros((nc, n), **f64), U=torch.as_tensor(self.filter_vars.astype(np.int64), device=dev),
                ll=torch.empty((nc,), **f64), st2=torch.empty((nc,), dtype=torch.int32, device=dev),
                Zb=(torch.empty((nc, p, na), **f64) if self._obs_lib is not None else None),
                rows=torch.empty((nc, self.n_filter, n), **f64), err_pos=torch.as_tensor(self.err_pos, device=dev),
            )  # fmt: skip
        stream = torch.cuda.current_stream(dev).cuda_stream
        n_err, nf, U = len(self.measurement_error), self.n_filter, g["U"]

        def mark(name):
            if events is None:
                return None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            events.append((name, e0, e1))
            return e1

        for lo in range(0, N, nc):
            cnt = min(nc, N - lo)
            th = theta_full[lo : lo + cnt]
            ws["theta"][:cnt].copy_(th[:, : m.n_theta])
            if self.full_covariance:
                ws["Q"][:cnt].copy_(th[:, m.n_theta : m.n_theta + self._n_cov].reshape(cnt, m.k, m.k))
            else:
                ws["sig"][:cnt].copy_(th[:, m.n_theta : m.n_theta + m.k])
            if n_err:
                ws["herr"][:cnt].index_copy_(1, ws["err_pos"], th[:, m.n_theta + self._n_cov :])
            st = ws["status"][:cnt]
            e = mark("jacobian")
            m.jacobian_device(ws["theta"][:cnt], ws["A"], ws["B"], ws["C"], ws["D"], ws.get("xss"), st, stream)
            e and e.record()
            if self.ss_obs_intercept:
                xs = ws["xss"][:cnt].index_select(1, ws["d_var"])
                ws["d"][:cnt].index_copy_(1, ws["d_pos"], torch.where(ws["d_loglin"], xs.log(), xs) * ws["d_scale"])
            if self._obs_lib is not None:
                rc = self._obs_lib.gecon_obs_batched(ws["theta"].data_ptr(), cnt, ws["Z"].data_ptr(), self.p * self.n_aug,
                                                     ws["d"].data_ptr(), self.p, C.c_void_p(stream))  # fmt: skip
                m.launches += 1
                if rc != 0:
                    raise L.GeconLibraryError(f"gecon_obs_batched failed with CUDA error {rc}")
            cr = L.CrArgs(
                struct_size=C.sizeof(L.CrArgs), A=ws["A"].data_ptr(), B=ws["B"].data_ptr(),
                C=(None if self.solver == "backward_direct" else ws["C"].data_ptr()), scan_semantics=int(self.solver == "scan_cycle_reduction"),
                D=ws["D"].data_ptr(), N=cnt, n=m.n, k=m.k, max_iter=self.max_iter, accumulate=1, tol=self.tol,
                resid_tol=self.solver_tol, unperm=None, T=g["Tfull"].data_ptr(), R=g["Rfull"].data_ptr(),
                status=st.data_ptr(), n_iter=ws["n_iter"].data_ptr(), resid=ws["resid"].data_ptr(), norms=None,
                n_out=0, n_lead=(int(ws["lead"].numel()) if self.check_bk else 0),
                lead_idx=(ws["lead"].data_ptr() if self.check_bk else None), n_unstable=ws["n_unstable"].data_ptr(),
                lag_lo=m.col_ranges[0], lag_hi=m.col_ranges[1], lead_lo=m.col_ranges[2], lead_hi=m.col_ranges[3],
            )  # fmt: skip
            e = mark("cr_solve")
            L.check((m._cr_solve or lib.gecon_cr_solve_batched)(C.byref(cr), C.c_void_p(stream)), "gecon_cr_solve_batched")
            e and e.record()
            if self.check_bk:
                bk = L.BkArgs(
                    struct_size=C.sizeof(L.BkArgs), A=ws["A"].data_ptr(), B=ws["B"].data_ptr(), C=ws["C"].data_ptr(), N=cnt,
                    n=m.n, n_lead=int(ws["lead"].numel()), lead_idx=ws["lead"].data_ptr(), accumulate=1, max_iter=0,
                    n_unstable=ws["n_unstable"].data_ptr(), status=st.data_ptr(),
                    skip_mask=L.ST_BK_CERTIFIED | L.ST_JAC_NONFINITE,
                )  # fmt: skip
                L.check(lib.gecon_bk_count_batched(C.byref(bk), C.c_void_p(stream)), "gecon_bk_count_batched")
            # the filter's (exactly reduced, possibly augmented) transition and selection blocks
            ws["T"][:cnt, :nf, :nf] = g["Tfull"][:cnt].index_select(1, U).index_select(2, U)
            ws["R"][:cnt, :nf] = g["Rfull"][:cnt].index_select(1, U)
            kg = L.KalmanGradArgs(
                struct_size=C.sizeof(L.KalmanGradArgs), T=ws["T"].data_ptr(), R=ws["R"].data_ptr(),
                qdiag=(None if self.full_covariance else ws["sig"].data_ptr()), q_stride=m.k,
                qfull=(ws["Q"].data_ptr() if self.full_covariance else None), qfull_stride=m.k * m.k,
                qfull_bar=(g["qb"].data_ptr() if self.full_covariance else None),
                hdiag=ws["herr"].data_ptr() if n_err else None, h_stride=self.p,
                Z=(ws["Z"].data_ptr() if self.dense_Z is not None else None),
                obs_idx=(ws["obs"].data_ptr() if self.dense_Z is None else None),
                d=(ws["d"].data_ptr() if "d" in ws else None), d_stride=(self.p if "d" in ws else 0),
                z_stride=(self.p * self.n_aug if self._obs_lib is not None else 0),
                Z_bar=(g["Zb"].data_ptr() if self._obs_lib is not None else None),
                Y=Y.data_ptr(), N=cnt, n=self.n_aug, k=m.k, p=self.p, Tobs=Tobs, jitter=self.cov_jitter,
                missing_fill=self.missing_fill_value, mvn_const_mode=(0 if self.mvn_const == "per_obs" else 1), lyap_max_iter=0,
                status_in=st.data_ptr(), gate_mask=self.gate_mask, sigma_inputs=1, ll=ll[lo : lo + cnt].data_ptr(),
                status=status[lo : lo + cnt].data_ptr(), T_bar=g["Tb_f"].data_ptr(), R_bar=g["Rb_f"].data_ptr(),
                q_bar=(None if self.full_covariance else g["qb"].data_ptr()), h_bar=g["hb"].data_ptr(), d_bar=g["db"].data_ptr(),
                mask_intercept=int(self.mask_intercept),
            )  # fmt: skip
            e = mark("kalman_grad")
            spec = self._grad_spec_lib()
            if spec is not None:  # the same kernel source compiled for this configuration's (n, k, p): csrc/grad_spec.cu
                L.check(spec.gecon_kalman_grad_spec(C.byref(kg), C.c_void_p(stream)), "gecon_kalman_grad_spec")
                m.launches += 1
            else:
                L.check(lib.gecon_kalman_grad_batched(C.byref(kg), C.c_void_p(stream)), "gecon_kalman_grad_batched")
            e and e.record()
            # scatter the filter-block adjoints back into solver order (the augmentation rows are constants)
            Tb, Rb = g["Tb"][:cnt], g["Rb"][:cnt]
            Tb.zero_()
            Rb.zero_()
            rows = g["rows"][:cnt]
            rows.zero_()
            rows.index_copy_(2, U, g["Tb_f"][:cnt, :nf, :nf])
            Tb.index_copy_(1, U, rows)
            Rb.index_copy_(1, U, g["Rb_f"][:cnt, :nf])
            pa = L.PolicyAdjointArgs(
                struct_size=C.sizeof(L.PolicyAdjointArgs), A=ws["A"].data_ptr(), B=ws["B"].data_ptr(), C=ws["C"].data_ptr(),
                D=ws["D"].data_ptr(), T=g["Tfull"].data_ptr(), R=g["Rfull"].data_ptr(), T_bar=Tb.data_ptr(), R_bar=Rb.data_ptr(),
                N=cnt, n=m.n, k=m.k, max_iter=0, A_bar=g["Ab"].data_ptr(), B_bar=g["Bb"].data_ptr(), C_bar=g["Cb"].data_ptr(),
                D_bar=g["Db"].data_ptr(), status=g["st2"].data_ptr(),
            )  # fmt: skip
            e = mark("policy_adjoint")
            L.check(lib.gecon_policy_adjoint_batched(C.byref(pa), C.c_void_p(stream)), "gecon_policy_adjoint_batched")
            e and e.record()
            xssb = None
            if self.ss_obs_intercept:  # d = scale * log x_ss (or scale * x_ss): adjoint of the steady state
                xssb = g["xssb"][:cnt]
                xssb.zero_()
                xs = ws["xss"][:cnt].index_select(1, ws["d_var"])
                dbar = g["db"][:cnt].index_select(1, ws["d_pos"]) * ws["d_scale"]
                xssb.index_add_(1, ws["d_var"], torch.where(ws["d_loglin"], dbar / xs, dbar))
            e = mark("jacobian_vjp")
            m.vjp_device(ws["theta"][:cnt], g["Ab"], g["Bb"], g["Cb"], g["Db"], xssb, g["thb"], stream)
            if self._obs_lib is not None:  # the observation equations' share: theta_bar += <(Z_bar, d_bar), d(Z, d)/dtheta>
                rc = self._obs_lib.gecon_obs_vjp_batched(ws["theta"].data_ptr(), cnt, g["Zb"].data_ptr(), self.p * self.n_aug,
                                                         g["db"].data_ptr(), self.p, g["thb"].data_ptr(), C.c_void_p(stream))  # fmt: skip
                m.launches += 1
                if rc != 0:
                    raise L.GeconLibraryError(f"gecon_obs_vjp_batched failed with CUDA error {rc}")
            e and e.record()
            out = grad[lo : lo + cnt]
            out[:, : m.n_theta] = g["thb"][:cnt]
            out[:, m.n_theta : m.n_theta + self._n_cov] = g["qb"][:cnt]  # d/d sigma_<shock>, or the symmetrised d/d state_cov[i,j]
            if n_err:
                out[:, m.n_theta + self._n_cov :] = g["hb"][:cnt].index_select(1, g["err_pos"])
            bad = (status[lo : lo + cnt] != 0) | (g["st2"][:cnt] != 0)
            out.masked_fill_(bad[:, None], 0.0)
        if self.constant_params:
            grad = grad.index_select(1, torch.as_tensor(self._free_cols, device=dev))
        return ll, grad, status

    def loglik_and_grad(self, theta_full, Y, device="cuda:0"):
        """HOST arrays in, HOST arrays out: (ll, grad, status)."""
        L.require_device()
        th = np.ascontiguousarray(np.atleast_2d(theta_full), dtype=np.float64)
        dev = torch.device(device)
        th_d = torch.from_numpy(th).pin_memory().to(dev, non_blocking=True)
        Y_d = torch.as_tensor(np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, self.p)).to(dev)
        ll, grad, st = self.loglik_and_grad_device(th_d, Y_d)
        return ll.cpu().numpy(), grad.cpu().numpy(), st.cpu().numpy()

    def loglik(self, theta_full, Y, device="cuda:0"):
        """HOST arrays in, HOST arrays out (the end-to-end call a sampler makes): pinned staging, one H2D copy of the
        parameter population, the kernels, one D2H copy of (ll, status)."""
        L.require_device()
        th = np.ascontiguousarray(np.atleast_2d(theta_full), dtype=np.float64)
        dev = torch.device(device)
        th_d = torch.from_numpy(th).pin_memory().to(dev, non_blocking=True)
        Y_d = torch.as_tensor(np.ascontiguousarray(Y, dtype=np.float64).reshape(-1, self.p)).to(dev)
        ll, st = self.loglik_device(th_d, Y_d)
        return ll.cpu().numpy(), st.cpu().numpy()

    def sample_autocorrelation_matrices(self, theta_full, n_lags: int = 10, observed: bool = False, lag_step: int = 1):
        """``DSGEStateSpace.sample_autocorrelation_matrices`` (statespace.py:1217-1303) for a population of parameter draws:
        see ``geconpy_b200.model.posterior.sample_autocorrelation_matrices``."""
        from .posterior import sample_autocorrelation_matrices

        return sample_autocorrelation_matrices(self, theta_full, n_lags=n_lags, observed=observed, lag_step=lag_step)

    def solve(self, theta, device="cuda:0"):
        """theta[N, n_theta] -> dict(T, R, status, n_iter, resid, n_unstable) with T, R un-permuted to variable order
        (statespace.py:217-220), host arrays."""
        from .. import batched

        A, B, Cm, D, _xss, st_j = self.model.jacobian(theta)
        res = batched.cr_solve(
            A, B, Cm, D, max_iter=self.max_iter if self.configured else 1000, tol=self.tol if self.configured else 1e-8,
            resid_tol=self.solver_tol if self.configured else 1e-8, unperm=self.model.inv_var_order,
        )  # fmt: skip
        nu, st_b = batched.bk_count(A, B, Cm, self.model.permuted_lead_var_idx)
        return dict(T=res.T, R=res.R, status=res.status | st_j | st_b, n_iter=res.n_iter, resid=res.resid, n_unstable=nu)
