// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "../../include/gecon_b200.h"

namespace gecon {

void set_last_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;

inline int fail_cuda(cudaError_t e, const char* what) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

#define GECON_CUDA(call)                                   \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

inline int round_up8(int n) { return (n + 7) / 8 * 8; }

// The stream-ordered allocator gives freed memory back to the driver at the next synchronisation unless its pool is told to
// keep it: the per-call workspaces (cudaMallocAsync / cudaFreeAsync around every launch) would be re-mapped on every step.
// Called once per device before the first cudaMallocAsync.
inline void keep_mempool(void) {
    static std::atomic<unsigned> done{0u};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return;
    const unsigned bit = 1u << dev;
    if (done.load() & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done.fetch_or(bit);
}

// number of SMs of the current device (cached per device would be nicer; the query is cheap)
inline int sm_count() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms;
}

// Persistent-grid launch helper: sets the dynamic shared memory attribute, asks the occupancy calculator for the
// resident CTAs per SM and returns grid = min(N, SMs * CTAs/SM).
template <typename K>
inline int persistent_grid(K kernel, int threads, size_t smem, long long N, int* grid, int* ctas_per_sm, const char* cap_env = nullptr) {
    GECON_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // these kernels live in shared memory and barely touch L1: ask for the largest shared-memory carve-out
    GECON_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    GECON_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) {
        set_last_error("kernel does not fit on an SM (threads=%d, smem=%zu)", threads, smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    if (cap_env) {  // resident CTAs per SM capped from the environment: lets two kernels of different chunks share the SMs
        if (const char* e = getenv(cap_env)) {
            const int cap = atoi(e);
            if (cap > 0 && cap < per_sm) per_sm = cap;
        }
    }
    const int sms = sm_count();
    long long g = (long long)sms * per_sm;
    if (g > N) g = N;
    if (g < 1) g = 1;
    *grid = (int)g;
    if (ctas_per_sm) *ctas_per_sm = per_sm;
    return 0;
}

// RAII device buffer for the *_host entry points
struct DevBuf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() {
        return static_cast<T*>(p);
    }
};

#define GECON_DISPATCH_NP(np, ...)                  \
    switch (np) {                                   \
        case 8: { constexpr int NP_ = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int NP_ = 16; __VA_ARGS__; } break; \
        case 24: { constexpr int NP_ = 24; __VA_ARGS__; } break; \
        case 32: { constexpr int NP_ = 32; __VA_ARGS__; } break; \
        case 40: { constexpr int NP_ = 40; __VA_ARGS__; } break; \
        case 48: { constexpr int NP_ = 48; __VA_ARGS__; } break; \
        case 56: { constexpr int NP_ = 56; __VA_ARGS__; } break; \
        case 64: { constexpr int NP_ = 64; __VA_ARGS__; } break; \
        default:                                    \
            set_last_error("unsupported matrix dimension (padded %d > 64)", np); \
            return GECON_E_UNSUPPORTED_SIZE;        \
    }

// one-warp-per-draw cycle reduction (cr_warp.cuh, instantiated in cr_warp_inst.cu once per padded dimension): packed
// column ranges of the launch and the per-NP launchers (info != nullptr: only query {CTAs per SM, smem, threads})
struct cw_ranges {
    int o0, w0;  // lag columns: packed offset (even), packed width (hi - o0)
    int o2, w2;  // lead columns
};
int cr_warp_launch_np8(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info);
int cr_warp_launch_np16(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info);
int cr_warp_launch_np24(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info);
int cr_warp_launch_np32(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info);

// pencil dimensions of the Blanchard-Kahn count (n + n_lead) and the general solve: 3 / 2 tiles, up to 88
#define GECON_DISPATCH_NP_WIDE(np, ...)             \
    switch (np) {                                   \
        case 8: { constexpr int NP_ = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int NP_ = 16; __VA_ARGS__; } break; \
        case 24: { constexpr int NP_ = 24; __VA_ARGS__; } break; \
        case 32: { constexpr int NP_ = 32; __VA_ARGS__; } break; \
        case 40: { constexpr int NP_ = 40; __VA_ARGS__; } break; \
        case 48: { constexpr int NP_ = 48; __VA_ARGS__; } break; \
        case 56: { constexpr int NP_ = 56; __VA_ARGS__; } break; \
        case 64: { constexpr int NP_ = 64; __VA_ARGS__; } break; \
        case 72: { constexpr int NP_ = 72; __VA_ARGS__; } break; \
        case 80: { constexpr int NP_ = 80; __VA_ARGS__; } break; \
        case 88: { constexpr int NP_ = 88; __VA_ARGS__; } break; \
        default:                                    \
            set_last_error("unsupported matrix dimension (padded %d > 88)", np); \
            return GECON_E_UNSUPPORTED_SIZE;        \
    }

}  // namespace gecon
