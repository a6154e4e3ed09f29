// Instantiations of the one-warp-per-draw cycle-reduction kernel (cr_warp.cuh) for ONE padded dimension
// (compile with -DGECON_CW_NP=8|16|24|32) and every packed width C = 1 .. NP / 8.
#include <cstdlib>

#include "common.cuh"
#include "cr_warp.cuh"

#ifndef GECON_CW_NP
#error "compile with -DGECON_CW_NP=<padded dimension>"
#endif

namespace gecon {

template <int NP, int C>
static int launch_cw(const gecon_cr_args& a, const cw_ranges& rg, cudaStream_t st, int* info) {
    constexpr int WPC = 4;
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
    cr_warp_kernel<NP, C, WPC><<<grid, WPC * 32, smem, st>>>(a, rg);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
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
