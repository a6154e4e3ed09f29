// Instantiations of the Kalman kernel for ONE padded dimension (compile with -DGECON_KF_NP=8|16|...|56) and every
// number of observables p = 1..8.
#include <cstdlib>

#include "kalman.cuh"
#include "kalman_warp.cuh"
#include "kalman_warp_launch.cuh"

#ifndef GECON_KF_NP
#error "compile with -DGECON_KF_NP=<padded dimension>"
#endif

namespace gecon {

template <int NP, int PT>
static int launch_one(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    const size_t smem = KfSmem<NP>::bytes(a.Tobs, a.p);
    if (smem > 227 * 1024) {
        set_last_error("observation matrix does not fit in shared memory (%zu bytes needed)", smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    int grid = 0, per_sm = 0;
    int rc = persistent_grid(kalman_ll_kernel<NP, PT>, Cfg<NP>::NT, smem, a.N, &grid, &per_sm);
    if (rc) return rc;
    if (info) {
        info[0] = per_sm;
        info[1] = (int)smem;
        info[2] = Cfg<NP>::NT;
        return 0;
    }
    kalman_ll_kernel<NP, PT><<<grid, Cfg<NP>::NT, smem, st>>>(a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

#define GECON_CAT2(a, b) a##b
#define GECON_CAT(a, b) GECON_CAT2(a, b)

int GECON_CAT(launch_kf_, GECON_KF_NP)(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    constexpr int NP = GECON_KF_NP;
    switch (a.p) {
        case 1: return launch_one<NP, 1>(a, st, info);
        case 2: return launch_one<NP, 2>(a, st, info);
        case 3: return launch_one<NP, 3>(a, st, info);
        case 4: return launch_one<NP, 4>(a, st, info);
        case 5: return launch_one<NP, 5>(a, st, info);
        case 6: return launch_one<NP, 6>(a, st, info);
        case 7: return launch_one<NP, 7>(a, st, info);
        case 8: return launch_one<NP, 8>(a, st, info);
    }
    set_last_error("unsupported number of observables p = %d (1..8)", a.p);
    return GECON_E_UNSUPPORTED_SIZE;
}

#if GECON_KF_NP <= 32
int GECON_CAT(launch_kw_, GECON_KF_NP)(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    constexpr int NP = GECON_KF_NP;
    switch (a.p) {
        case 1: return launch_one_warp<NP, 1>(a, st, info);
        case 2: return launch_one_warp<NP, 2>(a, st, info);
        case 3: return launch_one_warp<NP, 3>(a, st, info);
        case 4: return launch_one_warp<NP, 4>(a, st, info);
        case 5: return launch_one_warp<NP, 5>(a, st, info);
        case 6: return launch_one_warp<NP, 6>(a, st, info);
        case 7: return launch_one_warp<NP, 7>(a, st, info);
        case 8: return launch_one_warp<NP, 8>(a, st, info);
    }
    set_last_error("unsupported number of observables p = %d (1..8)", a.p);
    return GECON_E_UNSUPPORTED_SIZE;
}
#endif

}  // namespace gecon
