// C ABI of the gradient path: gecon_kalman_grad_*, gecon_policy_adjoint_*.  The per-draw routines are
// in grad.cuh (shared with the CPU host-check build); this file holds the persistent-grid kernels and the launchers.
#include <cstdlib>

#include "common.cuh"
#include "grad_args.h"

namespace gecon {

int policy_adjoint_dmma_launch(const gecon_grad::PolicyAdjointArgs& g, cudaStream_t st);  // policy_adjoint.cu

__global__ void kalman_grad_kernel(const gecon_grad::KalmanGradArgs g) {
    extern __shared__ __align__(16) double sm_grad[];
    for (long long draw = blockIdx.x; draw < g.N; draw += gridDim.x) {
        gecon_grad::kalman_grad_draw(g, draw, (int)blockIdx.x, sm_grad);
        __syncthreads();
    }
}

__global__ void policy_adjoint_kernel(const gecon_grad::PolicyAdjointArgs g) {
    extern __shared__ __align__(16) double sm_grad[];
    __shared__ int s_int[4];
    for (long long draw = blockIdx.x; draw < g.N; draw += gridDim.x) {
        gecon_grad::policy_adjoint_draw(g, draw, sm_grad, s_int);
        __syncthreads();
    }
}

static int grad_threads(int n) {
    // experiment hook: GECON_GRAD_THREADS overrides the CTA size (a multiple of 32)
    if (const char* e = getenv("GECON_GRAD_THREADS")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 1024 && v % 32 == 0) return v;
    }
    // measured on B200 (medium NK: filter dimension 10, solver dimension 24) with the register-blocked products: one warp per
    // draw is best up to n = 12 (128 / 165 / 233 ms for 32 / 64 / 128 threads), 64-128 threads at n = 24, 256 is 1.6x slower
    return n <= 12 ? 32 : (n <= 32 ? 128 : 256);
}

static int check_kg(const gecon_kalman_grad_args* a) {
    if (!a || a->struct_size != sizeof(gecon_kalman_grad_args)) {
        set_last_error("gecon_kalman_grad_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (!a->T || !a->R || (!a->qfull && (!a->qdiag || !a->q_bar)) || (a->qfull && !a->qfull_bar) || !a->Y || !a->ll || !a->status || !a->T_bar || !a->R_bar || a->N < 0 || a->n < 1 || a->k < 1 ||
        a->p < 1 || a->Tobs < 0 || (!a->Z && !a->obs_idx)) {
        set_last_error("gecon_kalman_grad_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if ((a->Z && a->z_stride && a->z_stride != (int64_t)a->p * a->n) || (a->Z_bar && !a->Z)) {
        set_last_error("gecon_kalman_grad_args: z_stride must be 0 or p * n, and Z_bar needs a dense Z");
        return GECON_E_BADARG;
    }
    if (a->n > 48 || a->k > a->n || a->p > gecon_grad::PMAXG || a->p > a->n) {
        set_last_error("gecon_kalman_grad_args: unsupported size n = %d (max 48), k = %d (max n), p = %d (max %d)", a->n, a->k, a->p, gecon_grad::PMAXG);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

static int check_pa(const gecon_policy_adjoint_args* a) {
    if (!a || a->struct_size != sizeof(gecon_policy_adjoint_args)) {
        set_last_error("gecon_policy_adjoint_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (!a->B || !a->C || !a->T || !a->T_bar || !a->A_bar || !a->B_bar || !a->C_bar || a->N < 0 || a->n < 1 || a->k < 0) {
        set_last_error("gecon_policy_adjoint_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (a->n > 64 || a->k > a->n) {
        set_last_error("gecon_policy_adjoint_args: unsupported size n = %d (max 64), k = %d (max n)", a->n, a->k);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_kalman_grad_batched(const gecon_kalman_grad_args* a, void* stream) {
    int rc = check_kg(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    gecon_grad::KalmanGradArgs g = gecon_grad::to_internal(*a);
    const int nt = grad_threads(a->n);
    const size_t smem = sizeof(double) * gecon_grad::kalman_grad_smem_doubles(a->n, a->k, a->p, nt);
    int grid = 0;
    rc = persistent_grid(kalman_grad_kernel, nt, smem, a->N, &grid, nullptr);
    if (rc) return rc;
    const size_t traj_cta = (size_t)a->Tobs * gecon_grad::kalman_grad_traj_stride(a->n, a->p);
    const size_t per_cta = traj_cta + 2 * (size_t)a->n * a->n;
    double* ws = nullptr;
    GECON_CUDA(cudaMallocAsync((void**)&ws, sizeof(double) * per_cta * grid, st));
    g.traj = ws;
    g.c0bar_ws = ws + (size_t)grid * traj_cta;
    kalman_grad_kernel<<<grid, nt, smem, st>>>(g);
    g_launch_count++;
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, st);
    GECON_CUDA(le);
    return 0;
}

extern "C" int gecon_policy_adjoint_batched(const gecon_policy_adjoint_args* a, void* stream) {
    int rc = check_pa(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const gecon_grad::PolicyAdjointArgs g = gecon_grad::to_internal(*a);
    // tensor-path kernel (policy_adjoint.cu); GECON_PA_DFMA=1 keeps the host-checkable DFMA version below (same arithmetic, one source
    // with the CPU check build)
    static const bool use_dfma = getenv("GECON_PA_DFMA") && atoi(getenv("GECON_PA_DFMA")) != 0;
    if (!use_dfma) return policy_adjoint_dmma_launch(g, (cudaStream_t)stream);
    const int nt = grad_threads(a->n);
    const size_t smem = sizeof(double) * gecon_grad::policy_adjoint_smem_doubles(a->n, nt);
    int grid = 0;
    rc = persistent_grid(policy_adjoint_kernel, nt, smem, a->N, &grid, nullptr);
    if (rc) return rc;
    policy_adjoint_kernel<<<grid, nt, smem, (cudaStream_t)stream>>>(g);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

#define G_H2D(buf, field, type, count)                                                            \
    if (a->field) {                                                                               \
        GECON_CUDA(buf.alloc((count) * sizeof(type)));                                            \
        GECON_CUDA(cudaMemcpy(buf.p, a->field, (count) * sizeof(type), cudaMemcpyHostToDevice)); \
        d.field = buf.as<type>();                                                                 \
    }
#define G_OUT(buf, field, type, count)                 \
    if (a->field) {                                    \
        GECON_CUDA(buf.alloc((count) * sizeof(type))); \
        d.field = buf.as<type>();                      \
    }
#define G_D2H(field, type, count) \
    if (a->field) GECON_CUDA(cudaMemcpy(a->field, d.field, (count) * sizeof(type), cudaMemcpyDeviceToHost));

extern "C" int gecon_kalman_grad_host(const gecon_kalman_grad_args* a) {
    int rc = check_kg(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, k = a->k, p = a->p, Tobs = a->Tobs;
    DevBuf bT, bR, bq, bh, bZ, bo, bd, bY, bS, oll, ost, oT, oR, oq, oh, od, oZ;
    gecon_kalman_grad_args d = *a;
    G_H2D(bT, T, double, N * n * n)
    G_H2D(bR, R, double, N * n * k)
    G_H2D(bq, qdiag, double, (a->q_stride ? N * k : k))
    G_H2D(bh, hdiag, double, (a->h_stride ? N * p : p))
    G_H2D(bZ, Z, double, (a->z_stride ? N * p * n : p * n))
    G_H2D(bo, obs_idx, int32_t, p)
    G_H2D(bd, d, double, (a->d_stride ? N * p : p))
    G_H2D(bY, Y, double, Tobs * p)
    G_H2D(bS, status_in, int32_t, N)
    G_OUT(oll, ll, double, N)
    G_OUT(ost, status, int32_t, N)
    G_OUT(oT, T_bar, double, N * n * n)
    G_OUT(oR, R_bar, double, N * n * k)
    G_OUT(oq, q_bar, double, N * k)
    G_OUT(oh, h_bar, double, N * p)
    G_OUT(od, d_bar, double, N * p)
    G_OUT(oZ, Z_bar, double, N * p * n)
    rc = gecon_kalman_grad_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaDeviceSynchronize());
    G_D2H(ll, double, N)
    G_D2H(status, int32_t, N)
    G_D2H(T_bar, double, N * n * n)
    G_D2H(R_bar, double, N * n * k)
    G_D2H(q_bar, double, N * k)
    G_D2H(h_bar, double, N * p)
    G_D2H(d_bar, double, N * p)
    G_D2H(Z_bar, double, N * p * n)
    return 0;
}

extern "C" int gecon_policy_adjoint_host(const gecon_policy_adjoint_args* a) {
    int rc = check_pa(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, k = a->k;
    DevBuf bA, bB, bC, bD, bT, bR, bTb, bRb, oA, oB, oC, oD, oS;
    gecon_policy_adjoint_args d = *a;
    G_H2D(bA, A, double, N * n * n)
    G_H2D(bB, B, double, N * n * n)
    G_H2D(bC, C, double, N * n * n)
    G_H2D(bD, D, double, N * n * k)
    G_H2D(bT, T, double, N * n * n)
    G_H2D(bR, R, double, N * n * k)
    G_H2D(bTb, T_bar, double, N * n * n)
    G_H2D(bRb, R_bar, double, N * n * k)
    G_OUT(oA, A_bar, double, N * n * n)
    G_OUT(oB, B_bar, double, N * n * n)
    G_OUT(oC, C_bar, double, N * n * n)
    G_OUT(oD, D_bar, double, N * n * k)
    G_OUT(oS, status, int32_t, N)
    rc = gecon_policy_adjoint_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaDeviceSynchronize());
    G_D2H(A_bar, double, N * n * n)
    G_D2H(B_bar, double, N * n * n)
    G_D2H(C_bar, double, N * n * n)
    G_D2H(D_bar, double, N * n * k)
    G_D2H(status, int32_t, N)
    return 0;
}
