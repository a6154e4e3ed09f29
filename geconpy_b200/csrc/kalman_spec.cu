// Per-configuration build of the one-warp-per-draw filter (kalman_warp.cuh): compile with
//     -DGECON_KW_SPEC_N=<filter dimension> -DGECON_KW_SPEC_NP=<padded dimension: round_up8(max(n + 1, shocks))> -DGECON_KW_SPEC_P=<observables>
//     [-DGECON_KW_SPEC_TC=<leading non-zero columns of T, gecon_kalman_args.t_cols>]
// (geconpy_b200/build.py: build_filter_spec).  Exports  int gecon_kalman_ll_spec(const gecon_kalman_args*, void* stream)  with the
// contract of gecon_kalman_ll_batched: arguments this build does not cover (another n / p / padded dimension, a dense design matrix,
// a full shock covariance, samples too long for the warp kernel's shared memory, GECON_KF_SPEC=0) go to the generic entry point of
// the core library, so the function is a drop-in everywhere; the fused pipeline takes it as gecon_pipeline_args.kalman_ll.
#if !defined(GECON_KW_SPEC_N) || !defined(GECON_KW_SPEC_NP) || !defined(GECON_KW_SPEC_P)
#error "compile with -DGECON_KW_SPEC_N=<n> -DGECON_KW_SPEC_NP=<NP> -DGECON_KW_SPEC_P=<p>"
#endif
#ifndef GECON_KW_SPEC_TC
#define GECON_KW_SPEC_TC 0  // leading columns of T that can be non-zero (gecon_kalman_args.t_cols); 0: dense T
#endif
#include <cstdlib>

#include "kalman_warp_launch.cuh"

extern "C" int gecon_kalman_check_args(const gecon_kalman_args* args);
extern "C" int gecon_kalman_warp_np(const gecon_kalman_args* args);

static_assert(GECON_KW_SPEC_NP % 8 == 0 && GECON_KW_SPEC_NP <= 32 && GECON_KW_SPEC_N + 1 <= GECON_KW_SPEC_NP, "warp-per-draw filter: n + 1 <= NP <= 32");
static_assert(GECON_KW_SPEC_P >= 1 && GECON_KW_SPEC_P <= 8, "1..8 observables");

static_assert(GECON_KW_SPEC_TC >= 0 && GECON_KW_SPEC_TC < GECON_KW_SPEC_N, "t_cols: 0 (dense) or 1 .. n - 1");

extern "C" int gecon_kalman_ll_spec_t_cols(void) { return GECON_KW_SPEC_TC; }

extern "C" int gecon_kalman_ll_spec_dims(int32_t* n, int32_t* np, int32_t* p) {
    if (n) *n = GECON_KW_SPEC_N;
    if (np) *np = GECON_KW_SPEC_NP;
    if (p) *p = GECON_KW_SPEC_P;
    return 0;
}

extern "C" int gecon_kalman_ll_spec(const gecon_kalman_args* a, void* stream) {
    using namespace gecon;
    int rc = gecon_kalman_check_args(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const char* env = getenv("GECON_KF_SPEC");  // read per call: tests toggle it
    const bool enabled = !(env && atoi(env) == 0);
    const int tc = (a->t_cols > 0 && a->t_cols < a->n) ? a->t_cols : 0;
    if (!enabled || a->n != GECON_KW_SPEC_N || a->p != GECON_KW_SPEC_P || tc != GECON_KW_SPEC_TC || gecon_kalman_warp_np(a) != GECON_KW_SPEC_NP)
        return gecon_kalman_ll_batched(a, stream);
    return launch_one_warp<GECON_KW_SPEC_NP, GECON_KW_SPEC_P>(*a, (cudaStream_t)stream, nullptr);
}
