// Per-configuration build of the Kalman adjoint kernel: the SAME source as the generic kernel (grad.cuh) compiled with the filter
// dimension, the number of shocks and the number of observables of ONE state-space configuration as compile-time constants
//     -DGECON_GRAD_CN=<n> -DGECON_GRAD_CK=<k> -DGECON_GRAD_CP=<p>
// (geconpy_b200/build.py: build_grad_spec, cached by dimensions + source hash, like the generated Jacobian / observation kernels).
// Why: ncu on the generic kernel (profiles/r02_kalman_grad_*.txt) shows 6,760 warp-instructions per filter step of which 10 % are
// DFMA -- the rest is loop control and index arithmetic on run-time n, p, k (IMAD 28 %, BRA 9 %, ISETP 8 %).  With constant
// dimensions every inner loop unrolls and every shared-memory access gets an immediate offset.
// Exports  int gecon_kalman_grad_spec(const gecon_kalman_grad_args*, void* stream)  and  gecon_kalman_grad_spec_dims(int32_t[3]).
#if !defined(GECON_GRAD_CN) || !defined(GECON_GRAD_CK) || !defined(GECON_GRAD_CP)
#error "compile with -DGECON_GRAD_CN=<n> -DGECON_GRAD_CK=<k> -DGECON_GRAD_CP=<p>"
#endif
#include "common.cuh"
#include "grad_args.h"

namespace gecon {

__global__ void __launch_bounds__(GECON_GRAD_CN <= 12 ? 32 : (GECON_GRAD_CN <= 32 ? 128 : 256)) kalman_grad_spec_kernel(const gecon_grad::KalmanGradArgs g) {
    extern __shared__ __align__(16) double sm_grad[];
    for (long long draw = blockIdx.x; draw < g.N; draw += gridDim.x) {
        gecon_grad::kalman_grad_draw(g, draw, (int)blockIdx.x, sm_grad);
        __syncthreads();
    }
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_kalman_grad_spec_dims(int32_t* nkp) {
    nkp[0] = GECON_GRAD_CN;
    nkp[1] = GECON_GRAD_CK;
    nkp[2] = GECON_GRAD_CP;
    return 0;
}

extern "C" int gecon_kalman_grad_spec(const gecon_kalman_grad_args* a, void* stream) {
    if (!a || a->struct_size != sizeof(gecon_kalman_grad_args)) {
        set_last_error("gecon_kalman_grad_spec: bad struct_size");
        return GECON_E_BADARG;
    }
    if (a->n != GECON_GRAD_CN || a->k != GECON_GRAD_CK || a->p != GECON_GRAD_CP) {
        set_last_error("gecon_kalman_grad_spec: built for (n, k, p) = (%d, %d, %d), called with (%d, %d, %d)", GECON_GRAD_CN, GECON_GRAD_CK,
                       GECON_GRAD_CP, a->n, a->k, a->p);
        return GECON_E_BADARG;
    }
    if (!a->T || !a->R || (!a->qfull && (!a->qdiag || !a->q_bar)) || (a->qfull && !a->qfull_bar) || !a->Y || !a->ll || !a->status || !a->T_bar ||
        !a->R_bar || a->N < 0 || a->Tobs < 0 || (!a->Z && !a->obs_idx)) {
        set_last_error("gecon_kalman_grad_spec: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    gecon_grad::KalmanGradArgs g = gecon_grad::to_internal(*a);
    constexpr int nt = GECON_GRAD_CN <= 12 ? 32 : (GECON_GRAD_CN <= 32 ? 128 : 256);
    const size_t smem = sizeof(double) * gecon_grad::kalman_grad_smem_doubles(a->n, a->k, a->p, nt);
    int grid = 0;
    int rc = persistent_grid(kalman_grad_spec_kernel, nt, smem, a->N, &grid, nullptr);
    if (rc) return rc;
    const size_t traj_cta = (size_t)a->Tobs * gecon_grad::kalman_grad_traj_stride(a->n, a->p);
    const size_t per_cta = traj_cta + 2 * (size_t)a->n * a->n;
    double* ws = nullptr;
    keep_mempool();
    GECON_CUDA(cudaMallocAsync((void**)&ws, sizeof(double) * per_cta * grid, st));
    g.traj = ws;
    g.c0bar_ws = ws + (size_t)grid * traj_cta;
    kalman_grad_spec_kernel<<<grid, nt, smem, st>>>(g);
    g_launch_count++;
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, st);
    GECON_CUDA(le);
    return 0;
}
