"""Build the sm_100a shared libraries in-tree with nvcc (no JIT cache: the built .so files travel with the repo).

``build_core()``  -> geconpy_b200/_lib/libgecon_b200.so   (cycle reduction, BK count, dlyap, Kalman, solve, gemm)
``build_model()`` -> geconpy_b200/_lib/models/libgecon_model_<name>_<hash>.so  (generated Jacobian kernels)
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
MODEL_LIBDIR = LIBDIR / "models"
CORE_LIB = LIBDIR / "libgecon_b200.so"
CORE_SOURCES = ["capi.cu", "cr_solve.cu", "kalman.cu", "bk_count.cu", "propagate.cu", "grad.cu", "eig.cu", "pipeline.cu", "policy_adjoint.cu"]
# the Kalman kernel is instantiated for (NP, p) in 8 x 8 combinations: one object per padded dimension NP, built in parallel
KALMAN_INST = "kalman_inst.cu"
KALMAN_NPS = [8, 16, 24, 32, 40, 48, 56, 64]
# the one-warp-per-draw cycle-reduction kernel: one object per padded dimension (every packed width C = 1 .. NP / 8)
CR_WARP_INST = "cr_warp_inst.cu"
CR_WARP_NPS = [8, 16, 24, 32]

# what a generated model source pulls in when it carries its own build of the solver (model/codegen.py: cr_spec_block)
MODEL_SOLVER_DEPS = ["cr_warp_spec.cu", "cr_warp.cuh", "linalg.cuh", "common.cuh"]

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "--expt-relaxed-constexpr",
]
# experiment hook: extra nvcc flags (e.g. GECON_NVCC_EXTRA="-DGECON_CW_EXPERIMENT_WPC=13"); part of the build digest
NVCC_FLAGS += os.environ.get("GECON_NVCC_EXTRA", "").split()


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA libraries of geconpy_b200 cannot be built")


def _digest(paths, extra=()) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(Path(p).read_bytes())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def build_core(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu -> _lib/libgecon_b200.so (skipped when the sources are unchanged)."""
    LIBDIR.mkdir(exist_ok=True)
    deps = [CSRC / s for s in CORE_SOURCES + [KALMAN_INST, CR_WARP_INST]] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "gecon_b200.h"]
    stamp = LIBDIR / "libgecon_b200.stamp"
    dig = _digest(deps, NVCC_FLAGS)
    if not force and CORE_LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return CORE_LIB
    nvcc = find_nvcc()
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(job):
        src, defines, stem = job
        obj = objdir / (stem + ".o")
        out = _run([nvcc, *NVCC_FLAGS, *defines, *(["-Xptxas", "-v"] if verbose else []), "-c", str(CSRC / src), "-o", str(obj)])
        if verbose:
            print(out)
        return obj

    jobs = [(s, [], Path(s).stem) for s in CORE_SOURCES]
    jobs += [(KALMAN_INST, [f"-DGECON_KF_NP={np_}"], f"kalman_inst_np{np_}") for np_ in KALMAN_NPS]
    jobs += [(CR_WARP_INST, [f"-DGECON_CW_NP={np_}"], f"cr_warp_inst_np{np_}") for np_ in CR_WARP_NPS]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, jobs))
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(CORE_LIB), *map(str, objs), "-lcudart"])
    stamp.write_text(dig)
    return CORE_LIB


def build_model(name: str, source: str, force: bool = False) -> Path:
    """Compile one generated model source (a string of CUDA C++) into its own shared library."""
    MODEL_LIBDIR.mkdir(parents=True, exist_ok=True)
    # the generated source includes the public header (gecon_pipeline_args) and, for models the warp-per-draw solver covers, the
    # per-model solver build (cr_warp_spec.cu and what it includes, found through -I csrc): all part of the digest
    incl = [PKG.parent / "include" / "gecon_b200.h"]
    if "cr_warp_spec.cu" in source:
        incl += [CSRC / f for f in MODEL_SOLVER_DEPS]
    dig = hashlib.sha256((source + "".join(f.read_text() for f in incl) + " ".join(NVCC_FLAGS)).encode()).hexdigest()[:16]
    lib = MODEL_LIBDIR / f"libgecon_model_{name}_{dig}.so"
    if lib.exists() and not force:
        return lib
    build_core()  # the model library links against it
    # several ranks (one process per GPU) may build the same model at once: write source and library under process-private
    # names and move them into place atomically, so that nobody ever dlopens a half-written file
    tag = f".{os.getpid()}.tmp"
    src = MODEL_LIBDIR / f"gecon_model_{name}_{dig}.cu"
    tmp_src, tmp_lib = src.with_name(src.name[:-3] + tag + ".cu"), lib.with_name(lib.name + tag)
    tmp_src.write_text(source)
    nvcc = find_nvcc()
    try:
        # (the fused entry point gecon_model_loglik calls gecon_loglik_pipeline of the core library: link it, found at run time
        # next door through $ORIGIN)
        link = ["-L", str(LIBDIR), "-lgecon_b200", "-Xlinker", "-rpath=$ORIGIN/.."]
        _run([nvcc, *NVCC_FLAGS, "-shared", "-I", str(PKG.parent / "include"), "-I", str(CSRC), "-o", str(tmp_lib), str(tmp_src), *link, "-lcudart"])
        os.replace(tmp_src, src)
        os.replace(tmp_lib, lib)
        import re

        stale = re.compile(rf"^(lib)?gecon_model_{re.escape(name)}_[0-9a-f]{{16}}\.(so|cu)$")  # builds of older sources of the same model
        for f in MODEL_LIBDIR.iterdir():
            if stale.match(f.name) and f not in (lib, src):
                f.unlink(missing_ok=True)
    finally:
        for f in (tmp_src, tmp_lib):
            if f.exists():
                f.unlink()
    return lib


def build_grad_spec(n: int, k: int, p: int, force: bool = False) -> Path:
    """The Kalman adjoint kernel compiled for ONE (filter dimension, shocks, observables) triple (csrc/grad_spec.cu): same source as
    the generic kernel, dimensions as compile-time constants.  Cached by dimensions + source digest next to the model libraries."""
    deps = [CSRC / f for f in ("grad_spec.cu", "grad.cuh", "grad_args.h", "common.cuh")] + [PKG.parent / "include" / "gecon_b200.h"]
    dig = _digest(deps, NVCC_FLAGS)[:16]
    MODEL_LIBDIR.mkdir(parents=True, exist_ok=True)
    lib = MODEL_LIBDIR / f"libgecon_grad_n{n}_k{k}_p{p}_{dig}.so"
    if lib.exists() and not force:
        return lib
    build_core()
    tmp_lib = lib.with_name(lib.name + f".{os.getpid()}.tmp")
    try:
        link = ["-L", str(LIBDIR), "-lgecon_b200", "-Xlinker", "-rpath=$ORIGIN/.."]
        _run([find_nvcc(), *NVCC_FLAGS, "-shared", f"-DGECON_GRAD_CN={n}", f"-DGECON_GRAD_CK={k}", f"-DGECON_GRAD_CP={p}", "-o", str(tmp_lib),
              str(CSRC / "grad_spec.cu"), *link, "-lcudart"])
        os.replace(tmp_lib, lib)
        for stale in MODEL_LIBDIR.glob(f"libgecon_grad_n{n}_k{k}_p{p}_*.so"):  # builds of older sources of the same configuration
            if stale != lib:
                stale.unlink(missing_ok=True)
    finally:
        if tmp_lib.exists():
            tmp_lib.unlink()
    return lib


def filter_spec_np(n: int, k: int, p: int) -> int:
    """Padded dimension at which the warp-per-draw filter runs an (n filter variables, k shocks, p observables) configuration, or 0
    when another kernel takes it (csrc/kalman.cu: thread per draw for n <= 4 and p <= 2, CTA per draw beyond 32)."""
    np_ = -(-max(n + 1, k) // 8) * 8
    if np_ > 32 or not (1 <= p <= 8) or p > n or (n <= 4 and p <= 2):
        return 0
    return np_


def filter_spec_path(n: int, k: int, p: int, t_cols: int = 0):
    """Where build_filter_spec puts (or would put) the library of this configuration; None when the configuration has no such build."""
    np_ = filter_spec_np(n, k, p)
    if not np_:
        return None
    t_cols = t_cols if 0 < t_cols < n else 0
    deps = [CSRC / f for f in ("kalman_spec.cu", "kalman_warp_launch.cuh", "kalman_warp.cuh", "kalman.cuh", "linalg.cuh", "common.cuh")]
    dig = _digest(deps + [PKG.parent / "include" / "gecon_b200.h"], NVCC_FLAGS)[:16]
    return MODEL_LIBDIR / f"libgecon_kf_n{n}_np{np_}_p{p}_tc{t_cols}_{dig}.so"


def build_filter_spec(n: int, k: int, p: int, t_cols: int = 0, force: bool = False):
    """The warp-per-draw Kalman filter compiled for ONE (filter dimension, padded dimension, observables, non-zero columns of T) tuple
    (csrc/kalman_spec.cu): same source as the generic kernel, the dimensions compile-time constants.  ``t_cols``: the leading columns of
    T that can be non-zero (``gecon_kalman_args.t_cols``; 0 = dense).  Cached next to the model libraries."""
    lib = filter_spec_path(n, k, p, t_cols)
    if lib is None or (lib.exists() and not force):
        return lib
    np_ = filter_spec_np(n, k, p)
    t_cols = t_cols if 0 < t_cols < n else 0
    MODEL_LIBDIR.mkdir(parents=True, exist_ok=True)
    build_core()
    tmp_lib = lib.with_name(lib.name + f".{os.getpid()}.tmp")
    try:
        link = ["-L", str(LIBDIR), "-lgecon_b200", "-Xlinker", "-rpath=$ORIGIN/.."]
        _run([find_nvcc(), *NVCC_FLAGS, "-shared", f"-DGECON_KW_SPEC_N={n}", f"-DGECON_KW_SPEC_NP={np_}", f"-DGECON_KW_SPEC_P={p}", f"-DGECON_KW_SPEC_TC={t_cols}", "-o", str(tmp_lib),
              str(CSRC / "kalman_spec.cu"), *link, "-lcudart"])
        os.replace(tmp_lib, lib)
        for stale in MODEL_LIBDIR.glob(f"libgecon_kf_n{n}_np{np_}_p{p}_tc{t_cols}_*.so"):  # builds of older sources of the same configuration
            if stale != lib:
                stale.unlink(missing_ok=True)
    finally:
        if tmp_lib.exists():
            tmp_lib.unlink()
    return lib


if __name__ == "__main__":
    import sys

    print(build_core(force="--force" in sys.argv, verbose="-v" in sys.argv))
