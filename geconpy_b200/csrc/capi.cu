// Library-level C-ABI entry points: error reporting, device information, and the two batched building blocks
// (general solve, DMMA product) that the cycle-reduction / Kalman kernels are made of.
#include <mutex>

#include "common.cuh"
#include "linalg.cuh"

namespace gecon {

std::atomic<long long> g_launch_count{0};
static thread_local char t_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------ batched solve
template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT) solve_kernel(const double* __restrict__ M, const double* __restrict__ RHS, long long N, int n,
                                                            int m, double* __restrict__ X, int* __restrict__ status) {
    using C = Cfg<NP>;
    extern __shared__ __align__(16) double sm[];
    double* W = sm;
    double* Xt = W + C::TILE;
    double* s_inv = Xt + C::TILE;
    int* s_piv = reinterpret_cast<int*>(s_inv + NP);
    int* s_flag = s_piv + NP;  // [NP + 1]
    for (long long draw = blockIdx.x; draw < N; draw += gridDim.x) {
        tile_load<NP>(W, M + (size_t)draw * n * n, n, n, n);
        tile_load<NP>(Xt, RHS + (size_t)draw * n * m, n, m, m);
        __syncthreads();
        const bool ok = gj_solve_blocked<NP>(W, W, Xt, Xt, 0, (m + 7) >> 3, nullptr, nullptr, 0, 0, n, true, s_piv, s_flag);
        if (!ok) {
            tile_nanfill<NP>(Xt, n, m);
            __syncthreads();
        }
        tile_store<NP>(X + (size_t)draw * n * m, Xt, n, m, m, 1.0, nullptr, nullptr);
        if (threadIdx.x == 0 && status) status[draw] = ok ? 0 : GECON_ST_SINGULAR;
        __syncthreads();
    }
}

template <int NP>
struct SolveSmem {
    static constexpr size_t bytes = sizeof(double) * (2 * Cfg<NP>::TILE + NP) + sizeof(int) * (2 * NP + 8);
};

// ------------------------------------------------------------------------------------------------ batched product
template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT) gemm_kernel(const double* __restrict__ A, const double* __restrict__ B, long long N, int n,
                                                           int ta, int tb, double alpha, double* __restrict__ Cout) {
    using C = Cfg<NP>;
    extern __shared__ __align__(16) double sm[];
    double* At = sm;
    double* Bt = At + C::TILE;
    double* Ct = Bt + C::TILE;
    for (long long draw = blockIdx.x; draw < N; draw += gridDim.x) {
        tile_load<NP>(At, A + (size_t)draw * n * n, n, n, n);
        tile_load<NP>(Bt, B + (size_t)draw * n * n, n, n, n);
        __syncthreads();
        Acc<NP> acc;
        acc_zero(acc);
        if (!ta && !tb) gemm_acc<NP, false, false>(acc, At, Bt, alpha);
        else if (ta && !tb) gemm_acc<NP, true, false>(acc, At, Bt, alpha);
        else if (!ta && tb) gemm_acc<NP, false, true>(acc, At, Bt, alpha);
        else gemm_acc<NP, true, true>(acc, At, Bt, alpha);
        acc_store<NP>(acc, Ct);
        __syncthreads();
        tile_store<NP>(Cout + (size_t)draw * n * n, Ct, n, n, n, 1.0, nullptr, nullptr);
        __syncthreads();
    }
}

template <int NP>
static int launch_solve(const double* M, const double* RHS, long long N, int n, int m, double* X, int* status, cudaStream_t st) {
    int grid = 0;
    int rc = persistent_grid(solve_kernel<NP>, Cfg<NP>::NT, SolveSmem<NP>::bytes, N, &grid, nullptr);
    if (rc) return rc;
    solve_kernel<NP><<<grid, Cfg<NP>::NT, SolveSmem<NP>::bytes, st>>>(M, RHS, N, n, m, X, status);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

template <int NP>
static int launch_gemm(const double* A, const double* B, long long N, int n, int ta, int tb, double alpha, double* Cc, cudaStream_t st) {
    const size_t smem = sizeof(double) * 3 * Cfg<NP>::TILE;
    int grid = 0;
    int rc = persistent_grid(gemm_kernel<NP>, Cfg<NP>::NT, smem, N, &grid, nullptr);
    if (rc) return rc;
    gemm_kernel<NP><<<grid, Cfg<NP>::NT, smem, st>>>(A, B, N, n, ta, tb, alpha, Cc);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

int cr_kernel_info(int n, int* ctas, int* smem, int* threads);
int kf_kernel_info(int n, int p, int Tobs, int* ctas, int* smem, int* threads);
int lyap_kernel_info(int n, int* ctas, int* smem, int* threads);
int bk_kernel_info(int m, int* ctas, int* smem, int* threads);

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_abi_version(void) { return GECON_ABI_VERSION; }

extern "C" int gecon_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}

extern "C" const char* gecon_get_last_error(void) { return t_err; }

extern "C" int64_t gecon_launch_count(void) { return (int64_t)g_launch_count.load(); }

extern "C" int gecon_kernel_info(int32_t which, int32_t n, int32_t p, int32_t Tobs, int32_t* ctas_per_sm, int32_t* smem_bytes,
                                 int32_t* threads) {
    if (!ctas_per_sm || !smem_bytes || !threads || n < 1) {
        set_last_error("gecon_kernel_info: bad argument");
        return GECON_E_BADARG;
    }
    switch (which) {
        case 0: return cr_kernel_info(n, ctas_per_sm, smem_bytes, threads);
        case 1: return kf_kernel_info(n, p, Tobs, ctas_per_sm, smem_bytes, threads);
        case 2: return bk_kernel_info(n, ctas_per_sm, smem_bytes, threads);
        case 3: return lyap_kernel_info(n, ctas_per_sm, smem_bytes, threads);
    }
    set_last_error("gecon_kernel_info: unknown kernel %d", which);
    return GECON_E_BADARG;
}

extern "C" int gecon_solve_batched(const double* M, const double* RHS, int64_t N, int32_t n, int32_t m, double* X, int32_t* status,
                                   void* stream) {
    if (!M || !RHS || !X || N < 0 || n < 1 || m < 0) {
        set_last_error("gecon_solve_batched: bad argument");
        return GECON_E_BADARG;
    }
    if (m > round_up8(n)) {
        set_last_error("gecon_solve_batched: m = %d right-hand sides exceed the padded dimension", m);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    if (N == 0) return 0;
    const int np = round_up8(n);
    GECON_DISPATCH_NP_WIDE(np, return launch_solve<NP_>(M, RHS, N, n, m, X, status, (cudaStream_t)stream));
    return 0;
}

extern "C" int gecon_solve_host(const double* M, const double* RHS, int64_t N, int32_t n, int32_t m, double* X, int32_t* status) {
    if (!M || !RHS || !X || N < 0 || n < 1 || m < 0) {
        set_last_error("gecon_solve_host: bad argument");
        return GECON_E_BADARG;
    }
    if (N == 0) return 0;
    DevBuf dM, dR, dX, dS;
    const size_t bm = (size_t)N * n * n * 8, br = (size_t)N * n * m * 8;
    GECON_CUDA(dM.alloc(bm));
    GECON_CUDA(dR.alloc(br));
    GECON_CUDA(dX.alloc(br));
    GECON_CUDA(dS.alloc((size_t)N * 4));
    GECON_CUDA(cudaMemcpy(dM.p, M, bm, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dR.p, RHS, br, cudaMemcpyHostToDevice));
    int rc = gecon_solve_batched(dM.as<double>(), dR.as<double>(), N, n, m, dX.as<double>(), dS.as<int32_t>(), nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(X, dX.p, br, cudaMemcpyDeviceToHost));
    if (status) GECON_CUDA(cudaMemcpy(status, dS.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int gecon_gemm_batched(const double* A, const double* B, int64_t N, int32_t n, int32_t trans_a, int32_t trans_b, double alpha,
                                  double* C, void* stream) {
    if (!A || !B || !C || N < 0 || n < 1) {
        set_last_error("gecon_gemm_batched: bad argument");
        return GECON_E_BADARG;
    }
    if (N == 0) return 0;
    const int np = round_up8(n);
    GECON_DISPATCH_NP(np, return launch_gemm<NP_>(A, B, N, n, trans_a, trans_b, alpha, C, (cudaStream_t)stream));
    return 0;
}

extern "C" int gecon_gemm_host(const double* A, const double* B, int64_t N, int32_t n, int32_t trans_a, int32_t trans_b, double alpha,
                               double* C) {
    if (!A || !B || !C || N < 0 || n < 1) {
        set_last_error("gecon_gemm_host: bad argument");
        return GECON_E_BADARG;
    }
    if (N == 0) return 0;
    DevBuf dA, dB, dC;
    const size_t bm = (size_t)N * n * n * 8;
    GECON_CUDA(dA.alloc(bm));
    GECON_CUDA(dB.alloc(bm));
    GECON_CUDA(dC.alloc(bm));
    GECON_CUDA(cudaMemcpy(dA.p, A, bm, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dB.p, B, bm, cudaMemcpyHostToDevice));
    int rc = gecon_gemm_batched(dA.as<double>(), dB.as<double>(), N, n, trans_a, trans_b, alpha, dC.as<double>(), nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(C, dC.p, bm, cudaMemcpyDeviceToHost));
    return 0;
}
