// Launch configuration of the one-warp-per-draw filter kernel (kalman_warp.cuh) per padded dimension: shared by the generic
// instantiations of the core library (kalman_inst.cu) and the per-configuration build (kalman_spec.cu).
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "kalman_warp.cuh"

namespace gecon {

// one warp per draw (kalman_warp.cuh): selector Z, n + 1 <= NP <= 32
template <int NP, int PT, int MINB, int WPC_ = 4>
static int launch_one_warp_b(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    using S = KwSmem<NP, PT, WPC_>;
    const size_t smem = S::bytes(a.Tobs);
    if (smem > 227 * 1024) {
        set_last_error("observation matrix does not fit in shared memory (%zu bytes needed)", smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    int grid = 0, per_sm = 0;
    int rc = persistent_grid(kalman_ll_warp_kernel<NP, PT, MINB, WPC_>, S::WPC * 32, smem, (a.N + S::WPC - 1) / S::WPC, &grid, &per_sm, "GECON_KF_CTAS_PER_SM");
    if (rc) return rc;
    if (info) {
        info[0] = per_sm;
        info[1] = (int)smem;
        info[2] = S::WPC * 32;
        return 0;
    }
    kalman_ll_warp_kernel<NP, PT, MINB, WPC_><<<grid, S::WPC * 32, smem, st>>>(a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

template <int NP, int PT>
static int launch_one_warp(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    // resident CTAs the register allocator must leave room for: 4 x 4 warps per SM up to NP = 16 (measured: 128 registers
    // with the constant term in shared memory beats 168 registers and 3 CTAs), 2 CTAs at NP = 24
    if constexpr (NP == 16) {
        // one CTA of 16 warps per SM (same 16 warps per SM as 4 CTAs of 4, but one staged copy of Y and one set-up per SM):
        // measured 45.6 -> 43.9 ms on the medium NK model; 20 warps (96 registers, spills) 50.5 ms, 12 warps 46.6 ms.
        // Falls back to 4-warp CTAs when 16 private tile sets + Y do not fit in shared memory.
        static const int wpc_try = std::getenv("GECON_KW16_WPC") ? std::atoi(std::getenv("GECON_KW16_WPC")) : 16;  // experiment hook
        if (wpc_try == 20 && KwSmem<NP, PT, 20>::bytes(a.Tobs) <= 227 * 1024) return launch_one_warp_b<NP, PT, 1, 20>(a, st, info);
        if (KwSmem<NP, PT, 16>::bytes(a.Tobs) <= 227 * 1024) return launch_one_warp_b<NP, PT, 1, 16>(a, st, info);
    }
    if constexpr (NP == 32) {
        // filter dimensions 24..31 (the 45-variable composite of BASELINE config 4b runs at u = 26): T fragments, P and W accumulators
        // fill the register file (255 registers), four 9 KB tiles per warp: one CTA of 5 warps per SM when Y leaves room, else 4.
        // Still 4-5x the CTA-per-draw kernel, which spends its time in five CTA barriers per step.
        if (KwSmem<NP, PT, 5>::bytes(a.Tobs) <= 227 * 1024) return launch_one_warp_b<NP, PT, 1, 5>(a, st, info);
        return launch_one_warp_b<NP, PT, 1, 4>(a, st, info);
    }
    // (NP = 24: one CTA of 8 warps instead of two of 4 measured the same, 99.8 vs 99.5 ms on the large NK model: not built)
    // NP = 8 needs 94 registers: five 4-warp CTAs per SM fit (2.57 -> 2.47 ms on the RBC workload)
    return launch_one_warp_b<NP, PT, (NP <= 8 ? 5 : NP <= 16 ? 4 : 2)>(a, st, info);
}

}  // namespace gecon
