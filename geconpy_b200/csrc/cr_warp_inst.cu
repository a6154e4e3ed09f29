// Instantiations of the one-warp-per-draw cycle-reduction kernel (cr_warp.cuh) for ONE padded dimension
// (compile with -DGECON_CW_NP=8|16|24|32) and every packed width C = 1 .. NP / 8.
#include <cstdlib>

#include "common.cuh"
#include "cr_warp.cuh"

#ifndef GECON_CW_NP
#error "compile with -DGECON_CW_NP=<padded dimension>"
#endif

namespace gecon {

// WPC = warps (draws in flight) per CTA.  The kernel has no CTA-wide synchronisation after its set-up, so the CTA shape only
// decides the granularity at which shared memory and registers are handed out: GECON_CW_WPC=13 (experiment hook) runs one
// CTA of 13 warps per SM where four-warp CTAs fit three times (12 warps).
template <int NP, int C, int WPC>
static int launch_cw_w(const gecon_cr_args& a, const cw_ranges& rg, cudaStream_t st, int* info) {
    const size_t smem = CwCfg<NP, C>::bytes(WPC);
    int grid = 0, per_sm = 0;
    int rc = persistent_grid(cr_warp_kernel<NP, C, WPC>, WPC * 32, smem, (a.N + WPC - 1) / WPC, &grid, &per_sm, "GECON_CR_CTAS_PER_SM");
    if (rc) return rc;
    if (info) {
        info[0] = per_sm;
        info[1] = (int)smem;
        info[2] = WPC * 32;
        return 0;
    }
    gecon_compact_jac cj{};
    if (a.compact) cj = *a.compact;
    cr_warp_kernel<NP, C, WPC><<<grid, WPC * 32, smem, st>>>(a, rg, cj);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

template <int NP, int C>
static int launch_cw(const gecon_cr_args& a, const cw_ranges& rg, cudaStream_t st, int* info) {
#ifdef GECON_CW_EXPERIMENT_WPC
    if (const char* e = getenv("GECON_CW_WPC")) {
        if (atoi(e) == GECON_CW_EXPERIMENT_WPC) {
            if constexpr (CwCfg<NP, C>::bytes(GECON_CW_EXPERIMENT_WPC) <= 227 * 1024) return launch_cw_w<NP, C, GECON_CW_EXPERIMENT_WPC>(a, rg, st, info);
        }
    }
#endif
    return launch_cw_w<NP, C, 4>(a, rg, st, info);
}

#define GECON_CW_CASE(c)                                             \
    case c:                                                          \
        if constexpr (8 * c <= GECON_CW_NP) return launch_cw<GECON_CW_NP, c>(a, rg, st, info); \
        break;

#define GECON_CW_CAT2(a, b) a##b
#define GECON_CW_CAT(a, b) GECON_CW_CAT2(a, b)

// defined once per NP: cr_warp_launch_np8, ..._np32
int GECON_CW_CAT(cr_warp_launch_np, GECON_CW_NP)(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info) {
    switch (c) {
        GECON_CW_CASE(1)
        GECON_CW_CASE(2)
        GECON_CW_CASE(3)
        GECON_CW_CASE(4)
        default:
            break;
    }
    set_last_error("cr_warp: no instantiation for NP = %d, C = %d", GECON_CW_NP, c);
    return GECON_E_UNSUPPORTED_SIZE;
}

}  // namespace gecon
