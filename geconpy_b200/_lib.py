"""ctypes binding of ``libgecon_b200.so`` (the C ABI declared in ``include/gecon_b200.h``).

The product has no CPU fallback: if the shared library is missing or cannot be loaded, every entry point raises
``GeconLibraryError`` -- it never routes to numpy/scipy.
"""

from __future__ import annotations

import ctypes as C
import os

from pathlib import Path

PKG = Path(__file__).resolve().parent
CORE_LIB = PKG / "_lib" / "libgecon_b200.so"
ABI_VERSION = 3  # include/gecon_b200.h: GECON_ABI_VERSION

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

# per-draw status bits (include/gecon_b200.h)
ST_CR_NOT_CONVERGED = 0x001
ST_CR_NAN = 0x002
ST_SINGULAR = 0x004
ST_RESID = 0x008
ST_BK = 0x010
ST_BK_INCONCLUSIVE = 0x020
ST_LYAP = 0x040
ST_NOT_PD = 0x080
ST_LL_NONFINITE = 0x100
ST_JAC_NONFINITE = 0x200
ST_SKIPPED = 0x400
ST_BK_CERTIFIED = 0x800  # informational, not a failure
ST_FAILURE_MASK = 0x7FF

STATUS_NAMES = {
    ST_CR_NOT_CONVERGED: "cycle_reduction_not_converged",
    ST_CR_NAN: "cycle_reduction_nan",
    ST_SINGULAR: "singular_solve",
    ST_RESID: "policy_residual_above_tol",
    ST_BK: "blanchard_kahn_violated",
    ST_BK_INCONCLUSIVE: "blanchard_kahn_inconclusive",
    ST_LYAP: "lyapunov_not_converged",
    ST_NOT_PD: "innovation_cov_not_pd",
    ST_LL_NONFINITE: "loglik_nonfinite",
    ST_JAC_NONFINITE: "jacobian_nonfinite",
    ST_SKIPPED: "skipped",
    ST_BK_CERTIFIED: "bk_certified_by_solver",
}


class GeconLibraryError(RuntimeError):
    pass


class CompactJac(C.Structure):
    _fields_ = [("vals", C.c_void_p), ("stride", C.c_int64), ("table", C.c_void_p), ("off", C.c_int32 * 5), ("reserved", C.c_int32)]


class CrArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("A", C.c_void_p),
        ("B", C.c_void_p),
        ("C", C.c_void_p),
        ("D", C.c_void_p),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("max_iter", C.c_int32),
        ("accumulate", C.c_int32),
        ("tol", C.c_double),
        ("resid_tol", C.c_double),
        ("unperm", C.c_void_p),
        ("T", C.c_void_p),
        ("R", C.c_void_p),
        ("status", C.c_void_p),
        ("n_iter", C.c_void_p),
        ("resid", C.c_void_p),
        ("norms", C.c_void_p),
        ("n_out", C.c_int32),
        ("n_lead", C.c_int32),
        ("lead_idx", C.c_void_p),
        ("n_unstable", C.c_void_p),
        ("solv_norms", C.c_void_p),
        ("trunc_tol", C.c_double),
        ("t_stride", C.c_int64),
        ("r_stride", C.c_int64),
        ("t_ld", C.c_int32),
        ("lag_lo", C.c_int32),
        ("lag_hi", C.c_int32),
        ("lead_lo", C.c_int32),
        ("lead_hi", C.c_int32),
        ("scan_semantics", C.c_int32),
        ("compact", C.c_void_p),
    ]


class PipelineArgs(C.Structure):
    """gecon_pipeline_args (the fused theta -> log-likelihood entry point)."""

    _fields_ = [
        ("struct_size", C.c_size_t),
        ("jacobian", C.c_void_p),
        ("nz_table", C.c_void_p),
        ("nz_off", C.c_void_p),
        ("nnz", C.c_int32),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("n_theta", C.c_int32),
        ("n_err", C.c_int32),
        ("p", C.c_int32),
        ("n_filter", C.c_int32),
        ("n_lead", C.c_int32),
        ("filter_vars", C.c_void_p),
        ("obs_idx", C.c_void_p),
        ("lead_idx", C.c_void_p),
        ("col_ranges", C.c_int32 * 4),
        ("theta", C.c_void_p),
        ("theta_stride", C.c_int64),
        ("N", C.c_int64),
        ("Y", C.c_void_p),
        ("Tobs", C.c_int32),
        ("max_iter", C.c_int32),
        ("tol", C.c_double),
        ("solver_tol", C.c_double),
        ("jitter", C.c_double),
        ("missing_fill", C.c_double),
        ("mvn_const_mode", C.c_int32),
        ("mask_intercept", C.c_int32),
        ("gate_mask", C.c_int32),
        ("check_bk", C.c_int32),
        ("scan_semantics", C.c_int32),
        ("timing", C.c_int32),
        ("chunk", C.c_int64),
        ("ll", C.c_void_p),
        ("status", C.c_void_p),
        ("n_iter", C.c_void_p),
        ("cr_solve", C.c_void_p),
        ("kalman_ll", C.c_void_p),
    ]


class BkArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("A", C.c_void_p),
        ("B", C.c_void_p),
        ("C", C.c_void_p),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("n_lead", C.c_int32),
        ("lead_idx", C.c_void_p),
        ("accumulate", C.c_int32),
        ("max_iter", C.c_int32),
        ("n_unstable", C.c_void_p),
        ("status", C.c_void_p),
        ("skip_mask", C.c_int32),
        ("reserved0", C.c_int32),
        ("compact", C.c_void_p),
    ]


class DlyapArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("T", C.c_void_p),
        ("R", C.c_void_p),
        ("qdiag", C.c_void_p),
        ("q_stride", C.c_int64),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("max_iter", C.c_int32),
        ("accumulate", C.c_int32),
        ("P", C.c_void_p),
        ("status", C.c_void_p),
        ("n_iter", C.c_void_p),
    ]


class KalmanArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("T", C.c_void_p),
        ("R", C.c_void_p),
        ("qdiag", C.c_void_p),
        ("q_stride", C.c_int64),
        ("hdiag", C.c_void_p),
        ("h_stride", C.c_int64),
        ("Z", C.c_void_p),
        ("obs_idx", C.c_void_p),
        ("d", C.c_void_p),
        ("d_stride", C.c_int64),
        ("Y", C.c_void_p),
        ("P0", C.c_void_p),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("p", C.c_int32),
        ("Tobs", C.c_int32),
        ("jitter", C.c_double),
        ("missing_fill", C.c_double),
        ("mvn_const_mode", C.c_int32),
        ("lyap_max_iter", C.c_int32),
        ("status_in", C.c_void_p),
        ("gate_mask", C.c_int32),
        ("sigma_inputs", C.c_int32),
        ("ll", C.c_void_p),
        ("status", C.c_void_p),
        ("ll_t", C.c_void_p),
        ("z_stride", C.c_int64),
        ("qfull", C.c_void_p),
        ("qfull_stride", C.c_int64),
        ("h_count", C.c_int32),
        ("t_cols", C.c_int32),
        ("mask_intercept", C.c_int32),
        ("reserved2", C.c_int32),
    ]


class KalmanGradArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("T", C.c_void_p),
        ("R", C.c_void_p),
        ("qdiag", C.c_void_p),
        ("q_stride", C.c_int64),
        ("hdiag", C.c_void_p),
        ("h_stride", C.c_int64),
        ("Z", C.c_void_p),
        ("obs_idx", C.c_void_p),
        ("d", C.c_void_p),
        ("d_stride", C.c_int64),
        ("Y", C.c_void_p),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("p", C.c_int32),
        ("Tobs", C.c_int32),
        ("jitter", C.c_double),
        ("missing_fill", C.c_double),
        ("mvn_const_mode", C.c_int32),
        ("lyap_max_iter", C.c_int32),
        ("status_in", C.c_void_p),
        ("gate_mask", C.c_int32),
        ("sigma_inputs", C.c_int32),
        ("ll", C.c_void_p),
        ("status", C.c_void_p),
        ("T_bar", C.c_void_p),
        ("R_bar", C.c_void_p),
        ("q_bar", C.c_void_p),
        ("h_bar", C.c_void_p),
        ("d_bar", C.c_void_p),
        ("z_stride", C.c_int64),
        ("Z_bar", C.c_void_p),
        ("mask_intercept", C.c_int32),
        ("reserved2", C.c_int32),
        ("qfull", C.c_void_p),
        ("qfull_stride", C.c_int64),
        ("qfull_bar", C.c_void_p),
    ]


class PolicyAdjointArgs(C.Structure):
    _fields_ = (
        [("struct_size", C.c_size_t)]
        + [(f, C.c_void_p) for f in ("A", "B", "C", "D", "T", "R", "T_bar", "R_bar")]
        + [("N", C.c_int64), ("n", C.c_int32), ("k", C.c_int32), ("max_iter", C.c_int32), ("reserved0", C.c_int32)]
        + [(f, C.c_void_p) for f in ("A_bar", "B_bar", "C_bar", "D_bar", "status")]
    )


class PropagateArgs(C.Structure):
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("T", C.c_void_p),
        ("R", C.c_void_p),
        ("X0", C.c_void_p),
        ("E", C.c_void_p),
        ("e_stride", C.c_int64),
        ("N", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("m", C.c_int32),
        ("L", C.c_int32),
        ("start_at_x0", C.c_int32),
        ("reserved0", C.c_int32),
        ("out", C.c_void_p),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "gecon_abi_version": (C.c_int, []),
    "gecon_fp64_peak": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gecon_device_count": (C.c_int, []),
    "gecon_get_last_error": (C.c_char_p, []),
    "gecon_launch_count": (C.c_int64, []),
    "gecon_kernel_info": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p]),
    "gecon_cr_solve_batched": (C.c_int, [C.POINTER(CrArgs), C.c_void_p]),
    "gecon_cr_check_args": (C.c_int, [C.POINTER(CrArgs)]),
    "gecon_kalman_check_args": (C.c_int, [C.POINTER(KalmanArgs)]),
    "gecon_kalman_warp_np": (C.c_int, [C.POINTER(KalmanArgs)]),
    "gecon_cr_solve_host": (C.c_int, [C.POINTER(CrArgs)]),
    "gecon_bk_count_batched": (C.c_int, [C.POINTER(BkArgs), C.c_void_p]),
    "gecon_bk_count_host": (C.c_int, [C.POINTER(BkArgs)]),
    "gecon_dlyap_batched": (C.c_int, [C.POINTER(DlyapArgs), C.c_void_p]),
    "gecon_dlyap_host": (C.c_int, [C.POINTER(DlyapArgs)]),
    "gecon_kalman_ll_batched": (C.c_int, [C.POINTER(KalmanArgs), C.c_void_p]),
    "gecon_kalman_ll_host": (C.c_int, [C.POINTER(KalmanArgs)]),
    "gecon_propagate_batched": (C.c_int, [C.POINTER(PropagateArgs), C.c_void_p]),
    "gecon_propagate_host": (C.c_int, [C.POINTER(PropagateArgs)]),
    "gecon_kalman_grad_batched": (C.c_int, [C.POINTER(KalmanGradArgs), C.c_void_p]),
    "gecon_kalman_grad_host": (C.c_int, [C.POINTER(KalmanGradArgs)]),
    "gecon_policy_adjoint_batched": (C.c_int, [C.POINTER(PolicyAdjointArgs), C.c_void_p]),
    "gecon_policy_adjoint_host": (C.c_int, [C.POINTER(PolicyAdjointArgs)]),
    "gecon_solve_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gecon_solve_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "gecon_gemm_batched": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p],
    ),
    "gecon_gemm_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p]),
    "gecon_loglik_pipeline": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gecon_pipeline_stage_ms": (C.c_int, [C.c_void_p]),
    "gecon_real_eig_batched": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gecon_real_eig_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library(path: os.PathLike | str | None = None) -> C.CDLL:
    """Load ``libgecon_b200.so`` (built in-tree by ``geconpy_b200.build.build_core``).  No fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else CORE_LIB
    if not p.exists():
        raise GeconLibraryError(
            f"{p} not found: build it with `python -m geconpy_b200.build` (needs nvcc). "
            "geconpy_b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(str(p))
    except OSError as e:
        raise GeconLibraryError(f"cannot load {p}: {e}") from e
    for name, (res, args) in EXPORTS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise GeconLibraryError(f"{p} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.gecon_abi_version() != ABI_VERSION:
        raise GeconLibraryError(f"{p}: ABI version {lib.gecon_abi_version()} != {ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load_library().gecon_get_last_error().decode(errors="replace")
        raise GeconLibraryError(f"{what or 'libgecon_b200'} failed with code {rc}: {msg}")


def require_device() -> None:
    if load_library().gecon_device_count() < 1:
        raise GeconLibraryError("no CUDA device visible: geconpy_b200 runs on B200 (sm_100a) only and has no CPU fallback")


def decode_status(word: int) -> list[str]:
    return [name for bit, name in STATUS_NAMES.items() if word & bit]
