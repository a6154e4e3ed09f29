// fp64 micro-benchmarks for the B200 roofline denominators used by this repo.
//
//   1. DFMA peak             (vector fp64 pipe; the denominator of roofline.frac)
//   2. DMMA m8n8k4 peak      (fp64 tensor path, mma.sync)
//   3. DMMA m16n8k{4,8,16}   (sm_90+ fp64 shapes, if they assemble for sm_100a)
//   4. shared-memory LDS.64 throughput for the fragment access patterns the
//      kernels use (row-major, ld = 4 mod 8)
//
// Build : nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma884(c[j][0], c[j][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef TRY_BIG_SHAPES
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, double av, double bv) {
    double c[NACC][4];
    double a[4] = {av, av + 1, av + 2, av + 3};
    double b[2] = {bv, bv + 1};
#pragma unroll
    for (int j = 0; j < NACC; ++j) { c[j][0] = threadIdx.x + j; c[j][1] = j; c[j][2] = 1; c[j][3] = 2; }
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma1688(c[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

// smem -> DMMA: one warp multiplies NT x NT tiles (8x8 each) of two smem matrices, k = 8*NT,
// repeatedly. Models the in-kernel GEMM: per k-step of 4, NT A-frags + NT B-frags, NT*NT DMMAs.
template <int NT>
__global__ void __launch_bounds__(128) k_smem_gemm(double* out, int reps) {
    constexpr int N = 8 * NT;
    constexpr int LD = N + 4;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* A = sm + warp * 2 * N * LD;
    double* B = A + N * LD;
    for (int i = lane; i < N * LD; i += 32) { A[i] = 1e-3 * (i % 7); B[i] = 1e-3 * (i % 5); }
    __syncwarp();
    double c[NT][NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    const int r = lane >> 2, q = lane & 3;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int k0 = 0; k0 < N; k0 += 4) {
            double af[NT], bf[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) af[i] = A[(8 * i + r) * LD + k0 + q];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[j] = B[(k0 + q) * LD + 8 * j + r];
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma884(c[i][j][0], c[i][j][1], af[i], bf[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same GEMM with plain DFMA and a TRxTC register tile per thread (cyclic row/col ownership).
template <int N, int NTI, int NTJ>
__global__ void __launch_bounds__(128) k_smem_gemm_dfma(double* out, int reps) {
    constexpr int LD = N + 1;
    constexpr int TR = N / NTI, TC = N / NTJ;
    static_assert(NTI * NTJ == 32, "one warp");
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* A = sm + warp * 2 * N * LD;
    double* B = A + N * LD;
    for (int i = lane; i < N * LD; i += 32) { A[i] = 1e-3 * (i % 7); B[i] = 1e-3 * (i % 5); }
    __syncwarp();
    const int ti = lane / NTJ, tj = lane % NTJ;
    double c[TR][TC];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) c[i][j] = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            double af[TR], bf[TC];
#pragma unroll
            for (int i = 0; i < TR; ++i) af[i] = A[(ti + i * NTI) * LD + k];
#pragma unroll
            for (int j = 0; j < TC; ++j) bf[j] = B[k * LD + tj + j * NTJ];
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < TC; ++j) c[i][j] = fma(af[i], bf[j], c[i][j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 8));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_prop\": %d", prop.name, sms, prop.clockRate);

    {   // DFMA: 8 chains x 256 thr x (sms*8) blocks
        const int blocks = sms * 8;
        double ms = time_ms([&] { k_dfma<<<blocks, 256>>>(out, 1.0000001, 1e-9); });
        double flop = 2.0 * 8 * ITERS * 256.0 * blocks;
        printf(", \"dfma_tflops\": %.3f", flop / ms * 1e-9);
    }
    {
        const int blocks = sms * 8;
        double ms4 = time_ms([&] { k_dmma884<4><<<blocks, 256>>>(out, 1.0000001, 1e-9); });
        double ms8 = time_ms([&] { k_dmma884<8><<<blocks, 256>>>(out, 1.0000001, 1e-9); });
        double ms16 = time_ms([&] { k_dmma884<16><<<blocks, 256>>>(out, 1.0000001, 1e-9); });
        double f = 2.0 * 8 * 8 * 4 * ITERS * 8.0 * blocks;  // per warp per mma: 256 fma
        printf(", \"dmma884_tflops_acc4\": %.3f, \"dmma884_tflops_acc8\": %.3f, \"dmma884_tflops_acc16\": %.3f",
               4 * f / ms4 * 1e-9, 8 * f / ms8 * 1e-9, 16 * f / ms16 * 1e-9);
        // low occupancy: 1 warp per SMSP
        double msl = time_ms([&] { k_dmma884<8><<<sms, 128>>>(out, 1.0000001, 1e-9); });
        double fl = 2.0 * 256 * ITERS * 8.0 * 4 * sms;
        printf(", \"dmma884_tflops_1warp_per_smsp_acc8\": %.3f", fl / msl * 1e-9);
    }
#ifdef TRY_BIG_SHAPES
    {
        const int blocks = sms * 8;
        double ms = time_ms([&] { k_dmma1688<4><<<blocks, 256>>>(out, 1.0000001, 1e-9); });
        double f = 2.0 * 16 * 8 * 8 * ITERS * 4 * 8.0 * blocks;
        printf(", \"dmma1688_tflops_acc4\": %.3f", f / ms * 1e-9);
    }
#endif
    {   // smem-fed GEMMs, 4 warps per CTA, enough CTAs to fill the SM's smem
        const int reps = 200;
        auto run_dmma = [&](auto kern, int NT, const char* name) {
            int N = 8 * NT, LD = N + 4;
            size_t smem = 4 * 2 * N * LD * sizeof(double);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
            int blocks = sms * per_sm;
            double ms = time_ms([&] { kern<<<blocks, 128, smem>>>(out, reps); });
            double flop = 2.0 * N * N * N * reps * 4.0 * blocks;
            printf(", \"%s\": {\"tflops\": %.3f, \"ctas_per_sm\": %d}", name, flop / ms * 1e-9, per_sm);
        };
        run_dmma(k_smem_gemm<2>, 2, "smem_dmma_n16");
        run_dmma(k_smem_gemm<3>, 3, "smem_dmma_n24");
        run_dmma(k_smem_gemm<4>, 4, "smem_dmma_n32");
        auto run_dfma = [&](auto kern, int N, const char* name) {
            int LD = N + 1;
            size_t smem = 4 * 2 * N * LD * sizeof(double);
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
            int blocks = sms * per_sm;
            double ms = time_ms([&] { kern<<<blocks, 128, smem>>>(out, reps); });
            double flop = 2.0 * N * N * N * reps * 4.0 * blocks;
            printf(", \"%s\": {\"tflops\": %.3f, \"ctas_per_sm\": %d}", name, flop / ms * 1e-9, per_sm);
        };
        run_dfma(k_smem_gemm_dfma<24, 4, 8>, 24, "smem_dfma_n24_6x3");
        run_dfma(k_smem_gemm_dfma<24, 8, 4>, 24, "smem_dfma_n24_3x6");
        run_dfma(k_smem_gemm_dfma<32, 4, 8>, 32, "smem_dfma_n32_8x4");
        run_dfma(k_smem_gemm_dfma<32, 8, 4>, 32, "smem_dfma_n32_4x8");
        run_dfma(k_smem_gemm_dfma<16, 4, 8>, 16, "smem_dfma_n16_4x2");
    }
    printf("}\n");
    return 0;
}
