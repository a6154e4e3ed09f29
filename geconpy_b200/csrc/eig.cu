// Eigenvalues of batched real general matrices, one WARP per matrix, the matrix resident in shared memory.
//
// Replaces the eigenvalue call behind gEconpy/pytensorf/real_eig.py:10-36 (RealEig.perform: numpy.linalg.eig, eigenvalues
// only on the forward path) and gEconpy/model/perturbation.py:448-505 (compute_bk_eigenvalues_pt: eigenvalues of
// M = (-Gamma0_sel + 1e-8 I)^-1 Gamma1_sel), SURVEY.md Appendix A.4 option (1).  Diagnostics, not the hot path: the
// likelihood pipeline only needs the COUNT of |lambda| > 1 (bk_count.cu); this kernel gives the per-eigenvalue table of
// check_bk_condition(return_value="dataframe") and lets the count be cross-checked.
//
// Algorithm (the LAPACK dgeev / EISPACK route for eigenvalues only):
//   1  balancing as dgebal (isolating permutation + Parlett-Reinsch scaling by powers of two; warp_real_eig, eig.cuh)
//   2  Householder reduction to upper Hessenberg form; reflectors applied with one lane per column / per row
//   3  Francis double-shift QR sweeps with deflation and the two exceptional shifts (EISPACK hqr): the scalar logic runs
//      redundantly in every lane on broadcast shared-memory reads, the row / column modifications of a sweep step are spread
//      over the lanes
// Output order is the deflation order; the host wrappers sort by modulus as RealEig does.
#include "eig.cuh"

namespace gecon {

__global__ void __launch_bounds__(32) real_eig_kernel(const double* __restrict__ M, long long N, int m, int balance, double* __restrict__ re,
                                                      double* __restrict__ im, int* __restrict__ status) {
    extern __shared__ __align__(16) double sm_eig[];
    const int ld = m | 1;  // odd leading dimension: rows and columns are both conflict-free
    double* H = sm_eig;
    double* ort = H + (size_t)m * ld;  // [m] Householder vector
    double* wr = ort + m;              // [m]
    double* wi = wr + m;               // [m]
    const int lane = threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    for (long long mat = blockIdx.x; mat < N; mat += gridDim.x) {
        const double* g = M + (size_t)mat * m * m;
        bool finite = true;
        for (int idx = lane; idx < m * m; idx += 32) {
            const int i = idx / m, j = idx - i * m;
            const double v = g[idx];
            finite = finite && (fabs(v) <= 1.7e308);
            H[i * ld + j] = v;
        }
        finite = __all_sync(0xffffffffu, finite);
        __syncwarp();
        int st = 0;
        if (!finite) {
            st = GECON_ST_LL_NONFINITE;
        } else {
            if (!warp_real_eig(H, ld, m, balance, ort, wr, wi, lane)) st = GECON_ST_BK_INCONCLUSIVE;
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            re[(size_t)mat * m + i] = st ? qnan : wr[i];
            im[(size_t)mat * m + i] = st ? qnan : wi[i];
        }
        if (lane == 0 && status) status[mat] = st;
        __syncwarp();
    }
}

static int check_eig(const double* M, long long N, int m, double* re, double* im) {
    if (!M || !re || !im || N < 0 || m < 1) {
        set_last_error("gecon_real_eig: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (m > 160) {
        set_last_error("gecon_real_eig: m = %d > 160 (the matrix lives in shared memory)", m);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_real_eig_batched(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status, void* stream) {
    int rc = check_eig(M, N, m, re, im);
    if (rc) return rc;
    if (N == 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)m * (m | 1) + 3 * (size_t)m);
    int grid = 0;
    rc = persistent_grid(real_eig_kernel, 32, smem, N, &grid, nullptr);
    if (rc) return rc;
    real_eig_kernel<<<grid, 32, smem, (cudaStream_t)stream>>>(M, N, m, balance, re, im, status);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gecon_real_eig_host(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status) {
    int rc = check_eig(M, N, m, re, im);
    if (rc) return rc;
    if (N == 0) return 0;
    DevBuf dM, dRe, dIm, dSt;
    const size_t bm = (size_t)N * m * m * sizeof(double), bv = (size_t)N * m * sizeof(double);
    GECON_CUDA(dM.alloc(bm));
    GECON_CUDA(dRe.alloc(bv));
    GECON_CUDA(dIm.alloc(bv));
    GECON_CUDA(dSt.alloc((size_t)N * sizeof(int32_t)));
    GECON_CUDA(cudaMemcpy(dM.p, M, bm, cudaMemcpyHostToDevice));
    rc = gecon_real_eig_batched(dM.as<double>(), N, m, balance, dRe.as<double>(), dIm.as<double>(), dSt.as<int32_t>(), nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(re, dRe.p, bv, cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(im, dIm.p, bv, cudaMemcpyDeviceToHost));
    if (status) GECON_CUDA(cudaMemcpy(status, dSt.p, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}
