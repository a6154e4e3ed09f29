// Per-model build of the one-warp-per-draw solver (cr_warp.cuh), #included at the end of every generated model source whose system
// fits the warp kernel (n <= 32, a forward-looking block):
//     #define GECON_CW_SPEC_N / _O0 / _W0 / _O2 / _W2 / _C     (variables, packed lag / lead column ranges, packed tiles)
//     #include "cr_warp_spec.cu"
// Exports  int gecon_model_cr_solve(const gecon_cr_args*, void* stream)  with the contract of gecon_cr_solve_batched: arguments this
// build does not cover (another n, diagnostics norms, a wider packed block than it was built for, GECON_CR_SPEC=0) go to the generic
// entry point of the core library, so the function is a drop-in everywhere; gecon_model_loglik hands it to the fused pipeline
// (gecon_pipeline_args.cr_solve).
#if !defined(GECON_CW_SPEC_N) || !defined(GECON_CW_SPEC_C)
#error "define GECON_CW_SPEC_N, _O0, _W0, _O2, _W2, _C before including cr_warp_spec.cu"
#endif
#include <cstdlib>

#include "common.cuh"
#include "cr_warp.cuh"

extern "C" int gecon_cr_check_args(const gecon_cr_args* args);  // core library: the argument validation of gecon_cr_solve_*

extern "C" int gecon_model_cr_solve(const gecon_cr_args* a, void* stream) {
    using namespace gecon;
    constexpr int NP = (GECON_CW_SPEC_N + 7) / 8 * 8, C = GECON_CW_SPEC_C, WPC = 4;
    static_assert(NP <= 32 && 8 * C <= NP, "the warp-per-draw solver covers n <= 32 and packed blocks no wider than the matrix");
    int rc = gecon_cr_check_args(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const char* env = getenv("GECON_CR_SPEC");  // read per call (one call per chunk of draws): tests toggle it
    const bool enabled = !(env && atoi(env) == 0);
    const bool has_c = a->compact ? (a->compact->off[3] > a->compact->off[2]) : (a->C != nullptr);
    const int kd = ((a->D || a->compact) && a->R) ? a->k : 0, nl = a->lead_idx ? a->n_lead : 0;
    if (!enabled || a->n != GECON_CW_SPEC_N || !has_c || a->solv_norms || kd > 8 * C || nl > 8 * C) return gecon_cr_solve_batched(a, stream);
    const size_t smem = CwCfg<NP, C>::bytes(WPC);
    int grid = 0;
    rc = persistent_grid(cr_warp_kernel<NP, C, WPC>, WPC * 32, smem, (a->N + WPC - 1) / WPC, &grid, nullptr, "GECON_CR_CTAS_PER_SM");
    if (rc) return rc;
    gecon_compact_jac cj{};
    if (a->compact) cj = *a->compact;
    cw_ranges rg{GECON_CW_SPEC_O0, GECON_CW_SPEC_W0, GECON_CW_SPEC_O2, GECON_CW_SPEC_W2};
    cr_warp_kernel<NP, C, WPC><<<grid, WPC * 32, smem, (cudaStream_t)stream>>>(*a, rg, cj);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}
