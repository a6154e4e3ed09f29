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

// ------------------------------------------------------------------------------------------------ fp64 peak probe
// Register-resident DFMA and DMMA (mma.sync.m8n8k4.f64) chains with enough independent accumulators to fill the pipe: the
// denominators of the roofline fractions bench.py reports, measured on the device the run is on (MEASURED_PEAKS.json has no
// fp64 figure).  Same kernels as profiles/microbench/fp64_peak.cu.
constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
    for (int i = 0; i < PEAK_ITERS; ++i) {
        x0 = fma(x0, a, b);
        x1 = fma(x1, a, b);
        x2 = fma(x2, a, b);
        x3 = fma(x3, a, b);
        x4 = fma(x4, a, b);
        x5 = fma(x5, a, b);
        x6 = fma(x6, a, b);
        x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        c[j][0] = threadIdx.x + j;
        c[j][1] = j;
    }
    for (int i = 0; i < PEAK_ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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

extern "C" int gecon_fp64_peak(double* dfma_tflops, double* dmma_tflops) {
    const int sms = sm_count();
    if (sms < 1) return GECON_E_NO_DEVICE;
    const int blocks = sms * 8, threads = 256;
    double* out = nullptr;
    GECON_CUDA(cudaMalloc((void**)&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best[2] = {0.0, 0.0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 6; ++rep) {  // the first repetitions warm the clocks up; best of the rest
            cudaEventRecord(e0, nullptr);
            if (which == 0) peak_dfma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9);
            else peak_dmma_kernel<<<blocks, threads>>>(out, 0.999999, 1e-9);
            cudaEventRecord(e1, nullptr);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            // DFMA: 8 chains x 2 flops per thread and iteration; DMMA: 4 products of 8 x 8 x 4 x 2 flops per warp and iteration
            const double flops = which == 0 ? (double)blocks * threads * PEAK_ITERS * 16.0 : (double)blocks * (threads / 32) * PEAK_ITERS * 4.0 * 512.0;
            const double tf = flops / (ms * 1e-3) / 1e12;
            if (rep >= 2 && tf > best[which]) best[which] = tf;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const cudaError_t le = cudaGetLastError();
    cudaFree(out);
    GECON_CUDA(le);
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return 0;
}

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
