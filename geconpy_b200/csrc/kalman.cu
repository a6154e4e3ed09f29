// Host entry points of the Kalman / Lyapunov kernels and the stand-alone dlyap kernel.  The Kalman kernel template
// (kalman.cuh) is instantiated in kalman_inst.cu, compiled once per padded dimension NP (build.py) so that the
// 7 x 8 (NP, p) instantiations build in parallel.
#include "kalman.cuh"
#include "kalman_thread.cuh"

namespace gecon {

// stand-alone dlyap kernel: 4 tiles (RQ->P, A, W, R staging shares W)
template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT) dlyap_kernel(const gecon_dlyap_args p) {
    using C = Cfg<NP>;
    extern __shared__ __align__(16) double sm[];
    double* P = sm;
    double* Aw = P + C::TILE;
    double* W = Aw + C::TILE;
    double* s_q = W + C::TILE;
    double* s_red = s_q + NP;
    int* s_i = reinterpret_cast<int*>(s_red + NP);
    const int n = p.n, k = p.k;
    const int cap = p.max_iter > 0 ? p.max_iter : 64;
    for (long long draw = blockIdx.x; draw < p.N; draw += gridDim.x) {
        tile_load<NP>(Aw, p.T + (size_t)draw * n * n, n, n, n);
        tile_load<NP>(W, p.R + (size_t)draw * n * k, n, k, k);
        tile_zero<NP>(P);
        if ((int)threadIdx.x < k) s_q[threadIdx.x] = p.qdiag[(size_t)draw * p.q_stride + threadIdx.x];
        __syncthreads();
        rqr_fill<NP>(P, W, s_q, n, k);
        int lo, hi;
        nonzero_col_range<NP>(Aw, n, s_i, lo, hi);
        bool bad = false;
        const int it = dlyap_doubling<NP>(P, Aw, W, n, lo & ~3, (hi + 3) & ~3, lo >> 3, (hi + 7) >> 3, cap, s_red, &bad);
        tile_symmetrize<NP>(P, n);
        __syncthreads();
        tile_store<NP>(p.P + (size_t)draw * n * n, P, n, n, n, 1.0, nullptr, nullptr);
        if (threadIdx.x == 0) {
            const int st = bad ? GECON_ST_LYAP : 0;
            p.status[draw] = p.accumulate ? (p.status[draw] | st) : st;
            if (p.n_iter) p.n_iter[draw] = it;
        }
        __syncthreads();
    }
}

template <int NP>
struct LyapSmem {
    static constexpr size_t bytes = sizeof(double) * (3 * Cfg<NP>::TILE + 2 * NP) + sizeof(int) * 4;
};

// ---------------------------------------------------------------------------------------------------------- host
static int check_kf_args(const gecon_kalman_args* a) {
    if (!a || a->struct_size != sizeof(gecon_kalman_args)) {
        set_last_error("gecon_kalman_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (!a->T || !a->R || (!a->qdiag && !a->qfull) || !a->Y || !a->ll || !a->status || a->N < 0 || a->n < 1 || a->k < 0 || a->p < 1 || a->Tobs < 0 ||
        (!a->Z && !a->obs_idx)) {
        set_last_error("gecon_kalman_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (a->Z && a->z_stride && a->z_stride != (int64_t)a->p * a->n) {
        set_last_error("gecon_kalman_args: z_stride must be 0 (shared Z) or p * n");
        return GECON_E_BADARG;
    }
    if (a->qfull && a->qfull_stride && a->qfull_stride != (int64_t)a->k * a->k) {
        set_last_error("gecon_kalman_args: qfull_stride must be 0 (shared Q) or k * k");
        return GECON_E_BADARG;
    }
    if (a->t_cols < 0 || a->t_cols > a->n) {
        set_last_error("gecon_kalman_args: t_cols = %d needs 0 <= t_cols <= n", a->t_cols);
        return GECON_E_BADARG;
    }
    if (a->p > PMAX || a->p > a->n) {
        set_last_error("gecon_kalman_args: unsupported p = %d (max %d, and p <= n)", a->p, PMAX);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

// defined in kalman_inst.cu (one translation unit per NP): launch, or only query occupancy when info != nullptr
#define GECON_KF_DECL(NPV) int launch_kf_##NPV(const gecon_kalman_args& a, cudaStream_t st, int* info);
GECON_KF_DECL(8) GECON_KF_DECL(16) GECON_KF_DECL(24) GECON_KF_DECL(32) GECON_KF_DECL(40) GECON_KF_DECL(48) GECON_KF_DECL(56) GECON_KF_DECL(64)
#undef GECON_KF_DECL
int launch_kw_8(const gecon_kalman_args& a, cudaStream_t st, int* info);
int launch_kw_16(const gecon_kalman_args& a, cudaStream_t st, int* info);
int launch_kw_24(const gecon_kalman_args& a, cudaStream_t st, int* info);
int launch_kw_32(const gecon_kalman_args& a, cudaStream_t st, int* info);

// padded dimension of the warp-per-draw kernel (needs a spare column for the mean), or 0 when the CTA kernel must run
static int warp_kernel_np(const gecon_kalman_args& a) {
    if (!a.obs_idx || a.Z || a.qfull) return 0;  // dense design matrices and full shock covariances stay on the CTA kernel
    const int np = round_up8((a.n + 1) > a.k ? (a.n + 1) : a.k);
    // NP = 32: four private 9 KB tiles per warp (149 KB for a 4-warp CTA); very long samples leave no room for them next to Y
    if (np == 32 && (size_t)a.Tobs * a.p * 8 + (size_t)a.Tobs * 4 > 70 * 1024) return 0;
    return np <= 32 ? np : 0;
}

// one thread per draw (kalman_thread.cuh): selector Z, diagonal Q, filter dimension <= 4, p <= 2 (GECON_KF_THREAD=0 disables it)
static int launch_kf_thread(const gecon_kalman_args& a, cudaStream_t st, int* info, bool* taken) {
    static const bool enabled = !(getenv("GECON_KF_THREAD") && atoi(getenv("GECON_KF_THREAD")) == 0);
    *taken = enabled && a.obs_idx && !a.Z && !a.qfull && a.qdiag && a.n >= 1 && a.n <= 4 && a.p >= 1 && a.p <= 2 && a.p <= a.n;
    if (!*taken) return 0;
#define GECON_KT_CASE(U_, P_) \
    if (a.n == U_ && a.p == P_) return launch_kalman_thread<U_, P_>(a, st, info);
    GECON_KT_CASE(1, 1) GECON_KT_CASE(2, 1) GECON_KT_CASE(3, 1) GECON_KT_CASE(4, 1)
    GECON_KT_CASE(2, 2) GECON_KT_CASE(3, 2) GECON_KT_CASE(4, 2)
#undef GECON_KT_CASE
    *taken = false;
    return 0;
}

static int launch_kf(int np, const gecon_kalman_args& a, cudaStream_t st, int* info) {
    {
        bool taken = false;
        const int rc = launch_kf_thread(a, st, info, &taken);
        if (taken) return rc;
    }
    switch (warp_kernel_np(a)) {
        case 8: return launch_kw_8(a, st, info);
        case 16: return launch_kw_16(a, st, info);
        case 24: return launch_kw_24(a, st, info);
        case 32: return launch_kw_32(a, st, info);
    }
    switch (np) {
        case 8: return launch_kf_8(a, st, info);
        case 16: return launch_kf_16(a, st, info);
        case 24: return launch_kf_24(a, st, info);
        case 32: return launch_kf_32(a, st, info);
        case 40: return launch_kf_40(a, st, info);
        case 48: return launch_kf_48(a, st, info);
        case 56: return launch_kf_56(a, st, info);
        case 64: return launch_kf_64(a, st, info);
    }
    set_last_error("unsupported matrix dimension (padded %d > 64)", np);
    return GECON_E_UNSUPPORTED_SIZE;
}

template <int NP>
static int launch_lyap(const gecon_dlyap_args& a, cudaStream_t st) {
    int grid = 0;
    int rc = persistent_grid(dlyap_kernel<NP>, Cfg<NP>::NT, LyapSmem<NP>::bytes, a.N, &grid, nullptr);
    if (rc) return rc;
    dlyap_kernel<NP><<<grid, Cfg<NP>::NT, LyapSmem<NP>::bytes, st>>>(a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

int kf_kernel_info(int n, int p, int Tobs, int* ctas, int* smem, int* threads) {
    if (p < 1 || p > PMAX) {
        set_last_error("kalman kernel info: p = %d out of range", p);
        return GECON_E_BADARG;
    }
    gecon_kalman_args a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.p = p;
    a.Tobs = Tobs;
    a.N = 1 << 30;
    static const int32_t dummy_obs = 0;
    a.obs_idx = &dummy_obs;  // info for the selector path (the one the pipeline uses)
    int info[3] = {0, 0, 0};
    int rc = launch_kf(round_up8(n), a, nullptr, info);
    if (rc) return rc;
    *ctas = info[0];
    *smem = info[1];
    *threads = info[2];
    return 0;
}

int lyap_kernel_info(int n, int* ctas, int* smem, int* threads) {
    const int np = round_up8(n);
    GECON_DISPATCH_NP(np, {
        int grid = 0;
        int rc = persistent_grid(dlyap_kernel<NP_>, Cfg<NP_>::NT, LyapSmem<NP_>::bytes, 1 << 30, &grid, ctas);
        if (rc) return rc;
        *smem = (int)LyapSmem<NP_>::bytes;
        *threads = Cfg<NP_>::NT;
    });
    return 0;
}

}  // namespace gecon

using namespace gecon;

// for the per-configuration builds (kalman_spec.cu): the argument validation of gecon_kalman_ll_*, and the padded dimension at which
// the generic entry point would run the warp-per-draw kernel on these arguments (0: it would not -- thread-per-draw or CTA-per-draw)
extern "C" int gecon_kalman_check_args(const gecon_kalman_args* args) { return check_kf_args(args); }
extern "C" int gecon_kalman_warp_np(const gecon_kalman_args* args) {
    bool taken = false;
    int info[3];
    if (launch_kf_thread(*args, nullptr, info, &taken) == 0 && taken) return 0;
    return warp_kernel_np(*args);
}

extern "C" int gecon_kalman_ll_batched(const gecon_kalman_args* args, void* stream) {
    int rc = check_kf_args(args);
    if (rc) return rc;
    if (args->N == 0) return 0;
    const int np = round_up8(args->n > args->k ? args->n : args->k);  // the R staging tile needs k columns
    return launch_kf(np, *args, (cudaStream_t)stream, nullptr);
}

extern "C" int gecon_dlyap_batched(const gecon_dlyap_args* a, void* stream) {
    if (!a || a->struct_size != sizeof(gecon_dlyap_args) || !a->T || !a->R || !a->qdiag || !a->P || !a->status || a->N < 0 ||
        a->n < 1 || a->k < 0) {
        set_last_error("gecon_dlyap_args: bad argument");
        return GECON_E_BADARG;
    }
    if (a->N == 0) return 0;
    const int np = round_up8(a->n > a->k ? a->n : a->k);
    GECON_DISPATCH_NP(np, return launch_lyap<NP_>(*a, (cudaStream_t)stream));
    return 0;
}

extern "C" int gecon_dlyap_host(const gecon_dlyap_args* a) {
    if (!a || a->struct_size != sizeof(gecon_dlyap_args)) {
        set_last_error("gecon_dlyap_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (a->N == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, k = a->k;
    DevBuf dT, dR, dq, dP, dSt, dIt;
    gecon_dlyap_args d = *a;
    const size_t nq = a->q_stride ? N * k : k;
    GECON_CUDA(dT.alloc(N * n * n * 8));
    GECON_CUDA(dR.alloc(N * n * k * 8));
    GECON_CUDA(dq.alloc(nq * 8));
    GECON_CUDA(dP.alloc(N * n * n * 8));
    GECON_CUDA(dSt.alloc(N * 4));
    GECON_CUDA(cudaMemcpy(dT.p, a->T, N * n * n * 8, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dR.p, a->R, N * n * k * 8, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dq.p, a->qdiag, nq * 8, cudaMemcpyHostToDevice));
    if (a->accumulate) GECON_CUDA(cudaMemcpy(dSt.p, a->status, N * 4, cudaMemcpyHostToDevice));
    d.T = dT.as<double>();
    d.R = dR.as<double>();
    d.qdiag = dq.as<double>();
    d.P = dP.as<double>();
    d.status = dSt.as<int32_t>();
    if (a->n_iter) {
        GECON_CUDA(dIt.alloc(N * 4));
        d.n_iter = dIt.as<int32_t>();
    }
    int rc = gecon_dlyap_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(a->P, d.P, N * n * n * 8, cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(a->status, d.status, N * 4, cudaMemcpyDeviceToHost));
    if (a->n_iter) GECON_CUDA(cudaMemcpy(a->n_iter, d.n_iter, N * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int gecon_kalman_ll_host(const gecon_kalman_args* a) {
    int rc = check_kf_args(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, k = a->k, p = a->p, Tobs = a->Tobs;
    DevBuf dT, dR, dq, dqf, dh, dZ, dobs, dd, dY, dP0, dSin, dll, dSt, dllt;
    gecon_kalman_args d = *a;
#define H2D(buf, field, type, count)                                                          \
    if (a->field) {                                                                           \
        GECON_CUDA(buf.alloc((count) * sizeof(type)));                                        \
        GECON_CUDA(cudaMemcpy(buf.p, a->field, (count) * sizeof(type), cudaMemcpyHostToDevice)); \
        d.field = buf.as<type>();                                                             \
    }
    H2D(dT, T, double, N * n * n)
    H2D(dR, R, double, N * n * k)
    H2D(dq, qdiag, double, (a->q_stride ? N * k : k))
    H2D(dqf, qfull, double, (a->qfull_stride ? N * k * k : k * k))
    H2D(dh, hdiag, double, (a->h_stride ? N * p : p))
    H2D(dZ, Z, double, (a->z_stride ? N * p * n : p * n))
    H2D(dobs, obs_idx, int32_t, p)
    H2D(dd, d, double, (a->d_stride ? N * p : p))
    H2D(dY, Y, double, Tobs * p)
    H2D(dP0, P0, double, N * n * n)
    H2D(dSin, status_in, int32_t, N)
#undef H2D
    GECON_CUDA(dll.alloc(N * 8));
    GECON_CUDA(dSt.alloc(N * 4));
    d.ll = dll.as<double>();
    d.status = dSt.as<int32_t>();
    if (a->ll_t) {
        GECON_CUDA(dllt.alloc(N * Tobs * 8));
        d.ll_t = dllt.as<double>();
    }
    rc = gecon_kalman_ll_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(a->ll, d.ll, N * 8, cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(a->status, d.status, N * 4, cudaMemcpyDeviceToHost));
    if (a->ll_t) GECON_CUDA(cudaMemcpy(a->ll_t, d.ll_t, N * Tobs * 8, cudaMemcpyDeviceToHost));
    return 0;
}
