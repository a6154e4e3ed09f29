"""CPU tests of the host side: C-ABI surface, code generation, package hygiene, draw sharding (gloo, world_size 2)."""

from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import sys

from pathlib import Path

import numpy as np
import pytest

from helpers import draws, jacobian_batch, model

ROOT = Path(__file__).resolve().parent.parent


# ------------------------------------------------------------------------------------------- C ABI
def test_library_loads_and_exports_every_declared_symbol():
    """Every function declared in include/gecon_b200.h must resolve in the built library (no compute calls here)."""
    from geconpy_b200 import _lib

    lib = _lib.load_library()
    header = (ROOT / "include" / "gecon_b200.h").read_text()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(gecon_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert {"gecon_cr_solve_batched", "gecon_kalman_ll_host", "gecon_bk_count_batched", "gecon_dlyap_host"} <= declared
    for name in declared:
        assert hasattr(lib, name), name
    assert declared <= set(_lib.EXPORTS), declared - set(_lib.EXPORTS)
    assert lib.gecon_abi_version() == _lib.ABI_VERSION == 3


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof() of the ctypes mirrors must equal what the C compiler lays out for include/gecon_b200.h."""
    from geconpy_b200 import _lib

    src = tmp_path / "abi_sizes.c"
    exe = tmp_path / "abi_sizes"
    src.write_text(
        f'#include <stdio.h>\n#include "{ROOT}/include/gecon_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gecon_cr_args), sizeof(gecon_bk_args), '
        "sizeof(gecon_dlyap_args), sizeof(gecon_kalman_args), sizeof(gecon_kalman_grad_args), sizeof(gecon_policy_adjoint_args), "
        "sizeof(gecon_propagate_args));return 0;}\n"
    )
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True, capture_output=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    mirrors = [_lib.CrArgs, _lib.BkArgs, _lib.DlyapArgs, _lib.KalmanArgs, _lib.KalmanGradArgs, _lib.PolicyAdjointArgs, _lib.PropagateArgs]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_device_is_a_loud_error_not_a_fallback():
    """Without a GPU the numeric entry points must raise -- never compute on the CPU."""
    from geconpy_b200 import _lib, batched

    if _lib.load_library().gecon_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.GeconLibraryError):
        batched.cr_solve(np.eye(3), np.eye(3), np.eye(3), np.ones((3, 1)))
    with pytest.raises(_lib.GeconLibraryError):
        batched.kalman_loglik(np.eye(2) * 0.5, np.eye(2), np.ones(2), np.zeros((4, 1)), obs_idx=[0])


def test_bad_arguments_are_rejected_by_the_library():
    from geconpy_b200 import _lib

    lib = _lib.load_library()
    args = _lib.CrArgs(struct_size=3)
    assert lib.gecon_cr_solve_batched(C.byref(args), None) == -1
    assert b"struct_size" in lib.gecon_get_last_error()
    big = _lib.CrArgs(struct_size=C.sizeof(_lib.CrArgs), A=1, B=1, T=1, status=1, N=1, n=65, k=0)
    assert lib.gecon_cr_solve_batched(C.byref(big), None) == -2  # n > 64 unsupported


def test_product_never_imports_the_oracle_or_the_reference():
    bad = []
    for path in list((ROOT / "geconpy_b200").rglob("*.py")):
        text = path.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "/root/reference" in text:
            bad.append(str(path))
    assert not bad, bad


# ------------------------------------------------------------------------------------------- code generation
@pytest.mark.parametrize("name", ["rbc", "full_nk", "open_rbc", "rbc_linearized"])
def test_generated_code_matches_oracle_when_compiled_for_the_host(name, tmp_path):
    """The generated expression code (same text nvcc compiles) built with g++ -DGECON_HOST_CHECK, against the oracle's
    independent lambdify evaluation: pins the code generator without a GPU."""
    from geconpy_b200.model.codegen import LinearizedModel, load_spec

    lin = LinearizedModel(load_spec(ROOT / "geconpy_b200" / "model" / "specs" / f"{name}.json"))
    mod = model(name)
    assert lin.var_names == mod.var_names and np.array_equal(lin.var_order, mod.var_order) and np.array_equal(lin.eq_order, mod.eq_order)
    assert np.array_equal(lin.permuted_lead_var_idx, mod.permuted_lead_var_idx)
    src = tmp_path / "m.cpp"
    so = tmp_path / "m.so"
    src.write_text(lin.cuda_source())
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-DGECON_HOST_CHECK", "-o", str(so), str(src)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    th = draws(mod, 16, seed=41, width=0.04)
    N, n, k = len(th), mod.n, mod.k
    A, B, Cm = (np.zeros((N, n, n)) for _ in range(3))
    D = np.zeros((N, n, k))
    xss = np.zeros((N, n))
    st = np.zeros(N, dtype=np.int32)
    lib.gecon_model_eval_host_check.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    lib.gecon_model_eval_host_check(th.ctypes.data, N, A.ctypes.data, B.ctypes.data, Cm.ctypes.data, D.ctypes.data, xss.ctypes.data, st.ctypes.data)
    Ao, Bo, Co, Do = jacobian_batch(mod, th)
    for i in range(N):
        fin = all(np.isfinite(M[i]).all() for M in (Ao, Bo, Co, Do))
        assert (st[i] == 0) == fin
        if fin:
            for G, O in ((A, Ao), (B, Bo), (Cm, Co), (D, Do)):
                np.testing.assert_allclose(G[i], O[i], rtol=1e-10, atol=1e-10)


def test_generated_source_is_deterministic():
    code = (
        "import sys; sys.path.insert(0, %r); from geconpy_b200.model.codegen import LinearizedModel, load_spec; import hashlib; "
        "print(hashlib.sha256(LinearizedModel(load_spec(%r)).cuda_source().encode()).hexdigest())"
    ) % (str(ROOT), str(ROOT / "geconpy_b200" / "model" / "specs" / "full_nk.json"))
    hashes = set()
    for seed in ("1", "2"):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        hashes.add(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.strip())
    assert len(hashes) == 1, "generated CUDA source depends on hash randomisation: prebuilt model libraries would not be found"


def test_codegen_rejects_models_outside_the_estimation_path():
    from geconpy_b200.model.codegen import LinearizedModel, load_spec

    spec = load_spec(ROOT / "geconpy_b200" / "model" / "specs" / "basic_rbc.json")
    with pytest.raises(NotImplementedError):  # no analytic steady state (build.py:658-659)
        LinearizedModel(spec)


# ------------------------------------------------------------------------------------------- host mirrors
def test_gensys_helpers_match_reference_test_vectors():
    """tests/solvers/test_gensys.py:12-70 of the reference."""
    from geconpy_b200.solvers import gensys as g

    alpha = np.array([-2.0123 - 0.5490j, -1.7594 + 0.4800j, 0.9347 - 0.1598j, 0.9237 + 0.1579j, 1.0847 + 0.0j])
    beta = np.array([2.2056, 1.9284, 2.4670, 2.4382, 1.3904], dtype=complex)
    assert g.determine_n_unstable(alpha, beta, 1.01, realsmall=1e-6) == (1.01, 5, False)
    div, nu, zxz = g.determine_n_unstable(np.array([1.0 + 0j, 1.0 + 0j]), np.array([1.005 + 0j, 2.0 + 0j]), None, 1e-6)
    assert abs(div - 0.5 * (1 + 1.005)) < 1e-15 and nu == 2 and zxz is False
    Q = np.arange(25, dtype=float).reshape(5, 5)
    Q1, Q2 = g.split_matrix_on_eigen_stability(Q, 3)
    assert Q1.shape == (2, 5) and Q2.shape == (3, 5) and np.array_equal(Q1, Q[:2])
    assert g.interpret_gensys_output([1, 0, 2]) == (
        "Gensys return codes: 1 0 2, with the following meaning:\nSolution exists, but is not unique."
    )  # tests/model/test_model.py:516-518


def test_gensys_setup_matches_oracle_pencil():
    from geconpy_b200.solvers.gensys import _gensys_setup
    from oracle import solvers as osol

    mod = model("full_nk")
    A, B, Cm, D = mod.jacobians(mod.theta_vector())
    for mine, ref in zip(_gensys_setup(A, B, Cm, D), osol.gensys_setup(A, B, Cm, D)):
        assert np.array_equal(mine, ref)


# ------------------------------------------------------------------------------------------- sharding (gloo)
def test_draw_sharding_and_allgather_world_size_2():
    """Two CPU processes over gloo: contiguous shards cover the population exactly once and the all-gather of per-draw
    values reassembles it in draw order (the SMC-stage exchange)."""
    script = ROOT / "tests" / "_gloo_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29577", str(script)],
        env=env, capture_output=True, text=True, timeout=300,
    )  # fmt: skip
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "gloo sharding ok" in r.stdout


def test_tempered_smc_stage_world_size_2():
    """TemperedSMC.stage under two gloo ranks with a stub likelihood: pack / all-gather of (weight, ll, status) / sorted
    ancestors / all_to_all of the surviving rows, both exchange modes identical, status resampled with its particle."""
    script = ROOT / "tests" / "_gloo_smc_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29579")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29579", str(script)],
        env=env, capture_output=True, text=True, timeout=300,
    )  # fmt: skip
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "gloo smc ok" in r.stdout


def test_gradient_entry_points_reject_bad_arguments_and_accept_empty_batches():
    """Argument checks of gecon_kalman_grad_* / gecon_policy_adjoint_* run before anything touches a device."""
    from geconpy_b200 import _lib

    lib = _lib.load_library()
    kg = _lib.KalmanGradArgs(struct_size=1)
    assert lib.gecon_kalman_grad_batched(C.byref(kg), None) == -1 and b"struct_size" in lib.gecon_get_last_error()
    ptr = dict(T=1, R=1, qdiag=1, Y=1, ll=1, status=1, T_bar=1, R_bar=1, q_bar=1, obs_idx=1)
    big = _lib.KalmanGradArgs(struct_size=C.sizeof(_lib.KalmanGradArgs), N=1, n=49, k=1, p=1, Tobs=1, **ptr)
    assert lib.gecon_kalman_grad_batched(C.byref(big), None) == -2  # n > 48 unsupported on the gradient path
    nosel = _lib.KalmanGradArgs(struct_size=C.sizeof(_lib.KalmanGradArgs), N=1, n=4, k=1, p=1, Tobs=1, **{**ptr, "obs_idx": None})
    assert lib.gecon_kalman_grad_batched(C.byref(nosel), None) == -1  # neither Z nor obs_idx
    empty = _lib.KalmanGradArgs(struct_size=C.sizeof(_lib.KalmanGradArgs), N=0, n=4, k=1, p=1, Tobs=1, **ptr)
    assert lib.gecon_kalman_grad_batched(C.byref(empty), None) == 0
    pa = _lib.PolicyAdjointArgs(struct_size=C.sizeof(_lib.PolicyAdjointArgs), B=1, C=1, T=1, T_bar=1, A_bar=1, B_bar=1, C_bar=1, N=0, n=4, k=0)
    assert lib.gecon_policy_adjoint_batched(C.byref(pa), None) == 0
    pa.n, pa.N = 65, 1
    assert lib.gecon_policy_adjoint_batched(C.byref(pa), None) == -2
    pa.n, pa.T_bar = 4, None
    assert lib.gecon_policy_adjoint_batched(C.byref(pa), None) == -1


def test_strided_solver_outputs_are_validated():
    from geconpy_b200 import _lib

    lib = _lib.load_library()
    a = _lib.CrArgs(struct_size=C.sizeof(_lib.CrArgs), A=1, B=1, T=1, status=1, N=1, n=8, k=0, t_ld=4)
    assert lib.gecon_cr_solve_batched(C.byref(a), None) == -1 and b"t_ld" in lib.gecon_get_last_error()
    a.t_ld, a.t_stride = 12, 12 * 12
    assert lib.gecon_cr_solve_host(C.byref(a)) == -1  # strided outputs are a device-entry-point feature
    kf = _lib.KalmanArgs(struct_size=C.sizeof(_lib.KalmanArgs), T=1, R=1, qdiag=1, Y=1, ll=1, status=1, Z=1, N=1, n=4, k=1, p=2, Tobs=1, z_stride=5)
    assert lib.gecon_kalman_ll_batched(C.byref(kf), None) == -1 and b"z_stride" in lib.gecon_get_last_error()


def test_grad_spec_source_compiles_for_any_dimensions_and_exports_its_entry_points(tmp_path):
    """csrc/grad_spec.cu (the per-configuration build of the Kalman adjoint kernel) cross-compiles for sm_100a with arbitrary
    compile-time (n, k, p), links against the core library and exports both entry points; its dimension check is a loud error."""
    import ctypes as C

    from geconpy_b200 import _lib as L
    from geconpy_b200.build import build_grad_spec

    lib = C.CDLL(str(build_grad_spec(5, 2, 2)))
    nkp = (C.c_int32 * 3)()
    assert lib.gecon_kalman_grad_spec_dims(nkp) == 0 and list(nkp) == [5, 2, 2]
    lib.gecon_kalman_grad_spec.argtypes = [C.POINTER(L.KalmanGradArgs), C.c_void_p]
    a = L.KalmanGradArgs(struct_size=C.sizeof(L.KalmanGradArgs), n=6, k=2, p=2)
    assert lib.gecon_kalman_grad_spec(C.byref(a), None) == -1  # GECON_E_BADARG
    assert b"built for (n, k, p) = (5, 2, 2)" in L.load_library().gecon_get_last_error()
    a = L.KalmanGradArgs(struct_size=C.sizeof(L.KalmanGradArgs) - 8, n=5, k=2, p=2)
    assert lib.gecon_kalman_grad_spec(C.byref(a), None) == -1  # GECON_E_BADARG


def test_filter_and_solver_spec_sources_compile_and_export_their_entry_points():
    """csrc/kalman_spec.cu (per-configuration filter) and csrc/cr_warp_spec.cu (per-model solver, inside the generated model library)
    cross-compile for sm_100a, link against the core library and export entry points with the contracts of the generic ones: the same
    argument validation runs first (no compute call without a GPU)."""
    import ctypes as C

    from geconpy_b200 import _lib as L
    from geconpy_b200.build import build_filter_spec, filter_spec_np, filter_spec_path
    from geconpy_b200.model.compiled import CompiledModel

    assert filter_spec_np(3, 1, 1) == 0 and filter_spec_path(3, 1, 1) is None  # thread-per-draw sizes have no such build
    assert filter_spec_np(40, 4, 3) == 0 and filter_spec_np(10, 4, 9) == 0      # CTA-per-draw sizes, p > 8
    assert filter_spec_np(10, 4, 3) == 16 and filter_spec_np(7, 4, 3) == 8 and filter_spec_np(5, 12, 3) == 16 and filter_spec_np(31, 9, 7) == 32
    lib = C.CDLL(str(build_filter_spec(7, 2, 3)))
    dims = (C.c_int32 * 3)()
    assert lib.gecon_kalman_ll_spec_dims(C.byref(dims, 0), C.byref(dims, 4), C.byref(dims, 8)) == 0 and list(dims) == [7, 8, 3]
    lib.gecon_kalman_ll_spec.argtypes = [C.POINTER(L.KalmanArgs), C.c_void_p]
    kf = L.KalmanArgs(struct_size=C.sizeof(L.KalmanArgs), T=1, R=1, qdiag=1, Y=1, ll=1, status=1, Z=1, N=1, n=7, k=2, p=3, Tobs=1, z_stride=5)
    assert lib.gecon_kalman_ll_spec(C.byref(kf), None) == -1 and b"z_stride" in L.load_library().gecon_get_last_error()
    kf = L.KalmanArgs(struct_size=C.sizeof(L.KalmanArgs) - 8)
    assert lib.gecon_kalman_ll_spec(C.byref(kf), None) == -1

    cm = CompiledModel("rbc")
    assert cm._cr_solve is not None
    a = L.CrArgs(struct_size=C.sizeof(L.CrArgs), A=1, B=1, T=1, status=1, N=1, n=cm.n, k=0, max_iter=10, tol=1e-8, t_ld=3)
    assert cm._cr_solve(C.byref(a), None) == -1
    assert CompiledModel("nk_rbc_composite")._cr_solve is None  # n = 45: beyond the warp-per-draw solver, generic kernels only


def test_posterior_helpers_validate_their_arguments_before_touching_the_device():
    from types import SimpleNamespace

    from geconpy_b200.model import posterior

    ss = SimpleNamespace(configured=False)
    with pytest.raises(RuntimeError, match="configure"):
        posterior.sample_autocorrelation_matrices(ss, np.zeros((1, 3)))
    with pytest.raises(RuntimeError, match="configure"):
        posterior.data_from_prior(ss)
    ss = SimpleNamespace(configured=True)
    with pytest.raises(ValueError, match="n_lags"):
        posterior.sample_autocorrelation_matrices(ss, np.zeros((1, 3)), n_lags=-1)
    with pytest.raises(ValueError, match="lag_step"):
        posterior.sample_autocorrelation_matrices(ss, np.zeros((1, 3)), lag_step=0)
    with pytest.raises(ValueError, match="pct_missing"):
        posterior.data_from_prior(ss, pct_missing=1.0)
