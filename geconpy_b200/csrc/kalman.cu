// Batched Kalman-filter log-likelihood (and the discrete Lyapunov solve for P0), one CTA per parameter draw.
//
// Semantics: pymc_extras StandardFilter as called from gEconpy/model/statespace.py:1151-1157 (restated in
// oracle/statespace.py): update -> jitter -> predict, Joseph-form covariance update,
// missing observations masked out of Z and H, a0 = 0 (statespace.py:812), P0 = dlyap(T, R Q R') (statespace.py:814).
//
// Shared memory: 5 tiles (T, RQR', P, A/scratch, W) + the p x p observation algebra + the whole observation
// matrix Y, brought in once per CTA with a 1-D TMA bulk copy (cp.async.bulk + mbarrier) that overlaps the
// Lyapunov solve of the CTA's first draw.
//
// Work actually skipped (never approximated): the columns of T that are identically zero (variables that do not
// appear with a lag; exact zeros because T = -A1hat^-1 A inherits the zero columns of A) are detected per draw and
// the k-loops of T P T' run only over the range that contains the non-zero columns.
#include "common.cuh"
#include "linalg.cuh"

namespace gecon {

constexpr int PMAX = 8;       // observables
constexpr int PS = PMAX + 1;  // odd row stride of the n x p work arrays: conflict-free row-per-thread access

template <int NP>
struct KfSmem {
    static constexpr int TILES = 5;
    // doubles: tiles | PZt, K (+1 row for v), KG | Z | F, G | a, af, q, red | v, w, d, h, dinv (PMAX each)
    static constexpr int NDBL = TILES * Cfg<NP>::TILE + (3 * NP + 1) * PS + PMAX * NP + 2 * PMAX * PS + 4 * NP + 5 * PMAX + 2;
    static constexpr int NINT = PMAX + 8;
    static size_t bytes(int Tobs, int p) {
        size_t ny = ((size_t)Tobs * p + 1) & ~(size_t)1;  // keep 16-byte granularity
        return sizeof(double) * (NDBL + ny) + sizeof(int) * NINT + 16;
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// R Q R' with Q = diag(q): exactly symmetric by construction.  Rt: tile holding R (n x k).  Fills the n x n corner.
template <int NP>
__device__ __forceinline__ void rqr_fill(double* __restrict__ RQ, const double* __restrict__ Rt, const double* __restrict__ q, int n, int k) {
    constexpr int LD = Cfg<NP>::LD;
    for (int idx = threadIdx.x; idx < n * n; idx += Cfg<NP>::NT) {
        const int i = idx / n, j = idx - i * n;
        if (i <= j) {
            double s = 0.0;
            for (int c = 0; c < k; ++c) s = fma(Rt[i * LD + c] * q[c], Rt[j * LD + c], s);
            RQ[i * LD + j] = s;
            RQ[j * LD + i] = s;
        }
    }
}

// Smith doubling for P = A P A' + Q:  P <- P + A_j P A_j',  A_{j+1} = A_j^2.  On entry P = Q, Aw = A (both tiles are
// overwritten; W is scratch).  Stops when max|increment| <= 1e-16 max|P|.  Returns iterations; sets *flag when the
// cap was hit or a NaN appeared.  Contains barriers.
template <int NP>
__device__ int dlyap_doubling(double* __restrict__ P, double* __restrict__ Aw, double* __restrict__ W, int n, int klo, int khi,
                              int ctlo, int cthi, int max_iter, double* __restrict__ s_red, bool* flag) {
    int it = 0;
    bool done = false;
    while (it < max_iter) {
        ++it;
        {
            Acc<NP> w;
            acc_zero(w);
            gemm_acc<NP, false, false>(w, Aw, P, 1.0, klo, khi, ctlo, cthi);
            acc_store<NP>(w, W, ctlo, cthi);
        }
        __syncthreads();
        Acc<NP> d, a2, pp;
        acc_zero(d);
        gemm_acc<NP, false, true>(d, W, Aw, 1.0, klo, khi);
        acc_zero(a2);
        gemm_acc<NP, false, false>(a2, Aw, Aw, 1.0, klo, khi, ctlo, cthi);
        acc_load<NP>(pp, P);
        double dmax = 0.0, pmax = 0.0;
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                pp.v[ct][e] += d.v[ct][e];
                const double ad = fabs(d.v[ct][e]), ap = fabs(pp.v[ct][e]);
                if (ad > dmax || ad != ad) dmax = ad;
                if (ap > pmax || ap != ap) pmax = ap;
            }
        }
        acc_store<NP>(pp, P);
        dmax = block_max<NP>(dmax, s_red);  // barriers: every warp is done reading Aw and W
        pmax = block_max<NP>(pmax, s_red);
        if (dmax != dmax || pmax != pmax) break;
        if (dmax <= 1e-16 * pmax) {
            done = true;
            break;
        }
        acc_store<NP>(a2, Aw, ctlo, cthi);
        __syncthreads();
    }
    __syncthreads();
    *flag = !done;
    return it;
}

template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT) kalman_ll_kernel(const gecon_kalman_args p) {
    using C = Cfg<NP>;
    constexpr int LD = C::LD, NT = C::NT;
    extern __shared__ __align__(16) double sm[];
    double* Tm = sm;
    double* RQ = Tm + C::TILE;
    double* P = RQ + C::TILE;
    double* Aw = P + C::TILE;
    double* W = Aw + C::TILE;
    double* s_Y = W + C::TILE;          // [Tobs][p], 16-byte aligned (five even-sized tiles precede it)
    const size_t ny = ((size_t)p.Tobs * p.p + 1) & ~(size_t)1;
    double* s_PZt = s_Y + ny;           // [NP][PS]
    double* s_K = s_PZt + NP * PS;      // [NP+1][PS]  (row n: innovation v)
    double* s_KG = s_K + (NP + 1) * PS; // [NP][PS]
    double* s_Z = s_KG + NP * PS;       // [PMAX][NP]
    double* s_F = s_Z + PMAX * NP;      // [PMAX][PS]  LDL' factor of F
    double* s_G = s_F + PMAX * PS;      // [PMAX][PS]  Zm P Zm' + Hm
    double* s_a = s_G + PMAX * PS;      // [NP] predicted mean
    double* s_af = s_a + NP;            // [NP] filtered mean
    double* s_q = s_af + NP;            // [NP] shock variances
    double* s_red = s_q + NP;           // [NP]
    double* s_v = s_red + NP;           // [PMAX]
    double* s_w = s_v + PMAX;           // [PMAX] 1 = observed, 0 = missing
    double* s_d = s_w + PMAX;
    double* s_h = s_d + PMAX;
    double* s_dinv = s_h + PMAX;
    double* s_sc = s_dinv + PMAX;       // [2] det, notpd flag
    int* s_obs = reinterpret_cast<int*>(s_sc + 2);  // [PMAX]
    int* s_i = s_obs + PMAX;                          // [4]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_i + 4) + 7) & ~(uintptr_t)7);

    const int n = p.n, k = p.k, np = p.p, Tobs = p.Tobs;
    const bool sel = (p.obs_idx != nullptr);
    const int tid = threadIdx.x;

    // ---- stage the observations once per CTA: 1-D TMA bulk copy (16-byte granules) + plain tail
    const uint32_t ybytes = (uint32_t)((size_t)Tobs * np * sizeof(double));
    const uint32_t ybulk = ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) ? (ybytes & ~15u) : 0u;
    if (tid == 0) mbar_init(s_bar, 1);
    __syncthreads();
    if (tid == 0) {
        if (ybulk) {
            mbar_expect_tx(s_bar, ybulk);
            tma_bulk_g2s(s_Y, p.Y, ybulk, s_bar);
        }
    }
    for (uint32_t i = ybulk / 8 + tid; i < ybytes / 8; i += NT) s_Y[i] = p.Y[i];
    for (int i = tid; i < PMAX * NP; i += NT) s_Z[i] = 0.0;
    if (tid < np) {
        s_obs[tid] = sel ? p.obs_idx[tid] : 0;
    }
    __syncthreads();
    if (!sel) {
        for (int i = tid; i < np * n; i += NT) {
            const int a = i / n, j = i - a * n;
            s_Z[a * NP + j] = p.Z[i];
        }
    }
    bool y_ready = (ybulk == 0);

    const double LOG2PI = 1.8378770664093453;
    const double ll_const = (p.mvn_const_mode == 0) ? np * LOG2PI : LOG2PI;
    const int lyap_cap = p.lyap_max_iter > 0 ? p.lyap_max_iter : 64;

    for (long long draw = blockIdx.x; draw < p.N; draw += gridDim.x) {
        // GECON_ST_BK_CERTIFIED is informational (set by the solver kernel): a clean draw leaves this kernel with status 0
        int status = p.status_in ? (p.status_in[draw] & ~GECON_ST_BK_CERTIFIED) : 0;
        if (status & p.gate_mask) {  // uniform: same value for every thread
            if (tid == 0) {
                p.ll[draw] = -INFINITY;
                p.status[draw] = status | GECON_ST_SKIPPED;
            }
            if (p.ll_t) {
                for (int t = tid; t < Tobs; t += NT) p.ll_t[(size_t)draw * Tobs + t] = -INFINITY;
            }
            continue;
        }
        // ---- load T, R, variances; RQR'
        tile_load<NP>(Tm, p.T + (size_t)draw * n * n, n, n, n);
        tile_load<NP>(W, p.R + (size_t)draw * n * k, n, k, k);
        tile_zero<NP>(RQ);
        if (tid < k) {
            const double qv = p.qdiag[(size_t)draw * p.q_stride + tid];
            s_q[tid] = p.sigma_inputs ? qv * qv : qv;
        }
        if (tid < np) {
            const double hv = p.hdiag ? p.hdiag[(size_t)draw * p.h_stride + tid] : 0.0;
            s_h[tid] = p.sigma_inputs ? hv * hv : hv;
            s_d[tid] = p.d ? p.d[(size_t)draw * p.d_stride + tid] : 0.0;
        }
        if (tid < NP) s_a[tid] = 0.0;
        __syncthreads();
        rqr_fill<NP>(RQ, W, s_q, n, k);
        int lo, hi;
        nonzero_col_range<NP>(Tm, n, s_i, lo, hi);  // barriers inside also publish RQ
        const int klo = lo & ~3, khi = (hi + 3) & ~3, ctlo = lo >> 3, cthi = (hi + 7) >> 3;

        // ---- P0
        if (p.P0) {
            tile_load<NP>(P, p.P0 + (size_t)draw * n * n, n, n, n);
            __syncthreads();
        } else {
            tile_copy<NP>(P, RQ);
            tile_copy<NP>(Aw, Tm);
            __syncthreads();
            bool bad = false;
            dlyap_doubling<NP>(P, Aw, W, n, klo, khi, ctlo, cthi, lyap_cap, s_red, &bad);
            if (bad) status |= GECON_ST_LYAP;
            tile_symmetrize<NP>(P, n);
            __syncthreads();
        }
        if (!y_ready) {
            mbar_wait(s_bar, 0);
            y_ready = true;
        }

        double ll_acc = 0.0;     // meaningful in thread n only
        bool notpd = false;      // thread n only
        for (int t = 0; t < Tobs; ++t) {
            const double* y = s_Y + (size_t)t * np;
            // ---- phase A: masks, PZt = P Zm', innovation v
            if (tid < np) {
                const double yv = y[tid];
                s_w[tid] = (yv != yv || yv == p.missing_fill) ? 0.0 : 1.0;
            }
            if (tid < n) {
                for (int a = 0; a < np; ++a) {
                    const double yv = y[a];
                    const double w = (yv != yv || yv == p.missing_fill) ? 0.0 : 1.0;
                    double s;
                    if (sel) {
                        s = P[s_obs[a] * LD + tid];
                    } else {
                        s = 0.0;
                        for (int j = 0; j < n; ++j) s = fma(P[j * LD + tid], s_Z[a * NP + j], s);
                    }
                    s_PZt[tid * PS + a] = w * s;
                    s_K[tid * PS + a] = w * s;
                }
            } else if (tid < n + np) {
                const int a = tid - n;
                const double yv = y[a];
                const bool miss = (yv != yv || yv == p.missing_fill);
                double za;
                if (sel) {
                    za = s_a[s_obs[a]];
                } else {
                    za = 0.0;
                    for (int j = 0; j < n; ++j) za = fma(s_Z[a * NP + j], s_a[j], za);
                }
                const double v = (miss ? 0.0 : yv) - (s_d[a] + (miss ? 0.0 : za));
                s_v[a] = v;
                s_K[n * PS + a] = v;
            }
            __syncthreads();
            // ---- phase B (warp 0): G = Zm PZt + Hm, F = G + jitter I, LDL' of F in place
            if (tid < 32) {
                for (int idx = tid; idx < PMAX * PMAX; idx += 32) {
                    const int a = idx >> 3, b = idx & 7;
                    if (b <= a && a < np) {
                        double s;
                        if (sel) {
                            s = s_PZt[s_obs[a] * PS + b];
                        } else {
                            s = 0.0;
                            for (int j = 0; j < n; ++j) s = fma(s_Z[a * NP + j], s_PZt[j * PS + b], s);
                        }
                        s *= s_w[a];
                        if (a == b) s += s_w[a] * s_h[a];
                        s_G[a * PS + b] = s;
                        s_G[b * PS + a] = s;
                        s_F[a * PS + b] = (a == b) ? s + p.jitter : s;
                    }
                }
                __syncwarp();
                double det = 1.0;
                bool bad = false;
                for (int c = 0; c < np; ++c) {
                    const double dc = s_F[c * PS + c];
                    if (!(dc > 0.0)) bad = true;
                    det *= dc;
                    const double inv = 1.0 / dc;
                    for (int idx = tid; idx < PMAX * PMAX; idx += 32) {
                        const int a = idx >> 3, b = idx & 7;
                        if (c < b && b <= a && a < np) s_F[a * PS + b] = fma(-s_F[a * PS + c] * inv, s_F[b * PS + c], s_F[a * PS + b]);
                    }
                    __syncwarp();
                    if (tid > c && tid < np) s_F[tid * PS + c] *= inv;
                    if (tid == 0) s_dinv[c] = inv;
                    __syncwarp();
                }
                if (tid == 0) {
                    s_sc[0] = det;
                    s_sc[1] = bad ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            // ---- phase C: rows of K = PZt F^-1 (and the innovation row) by forward / backward substitution
            if (tid <= n) {
                double* x = s_K + tid * PS;
                for (int a = 1; a < np; ++a) {
                    double s = x[a];
                    for (int b = 0; b < a; ++b) s = fma(-s_F[a * PS + b], x[b], s);
                    x[a] = s;
                }
                if (tid == n) {
                    double quad = 0.0;
                    for (int a = 0; a < np; ++a) quad = fma(x[a] * s_dinv[a], x[a], quad);
                    bool all_missing = true;
                    for (int a = 0; a < np; ++a) all_missing = all_missing && (s_w[a] == 0.0);
                    const double llt = all_missing ? 0.0 : -0.5 * (ll_const + log(s_sc[0]) + quad);
                    ll_acc += llt;
                    if (s_sc[1] != 0.0) notpd = true;
                    if (p.ll_t) p.ll_t[(size_t)draw * Tobs + t] = llt;
                } else {
                    for (int a = 0; a < np; ++a) x[a] *= s_dinv[a];
                    for (int a = np - 2; a >= 0; --a) {
                        double s = x[a];
                        for (int b = a + 1; b < np; ++b) s = fma(-s_F[b * PS + a], x[b], s);
                        x[a] = s;
                    }
                    for (int a = 0; a < np; ++a) {
                        double s = 0.0;
                        for (int b = 0; b < np; ++b) s = fma(x[b], s_G[b * PS + a], s);
                        s_KG[tid * PS + a] = s;
                    }
                }
            }
            __syncthreads();
            // ---- phase D: filtered mean and Joseph-form covariance, expanded:
            //      (I-KZ) P (I-KZ)' + K H K' = P - K PZt' - PZt K' + K (Z P Z' + H) K'   (+ jitter I)
            if (tid < n) {
                double s = s_a[tid];
                for (int a = 0; a < np; ++a) s = fma(s_K[tid * PS + a], s_v[a], s);
                s_af[tid] = s;
            }
            for (int idx = tid; idx < n * n; idx += NT) {
                const int i = idx / n, j = idx - i * n;
                double s1 = 0.0, s2 = 0.0;
                for (int a = 0; a < np; ++a) {
                    s1 += s_K[i * PS + a] * s_PZt[j * PS + a] + s_PZt[i * PS + a] * s_K[j * PS + a];
                    s2 += s_KG[i * PS + a] * s_K[j * PS + a] + s_KG[j * PS + a] * s_K[i * PS + a];
                }
                double v = P[i * LD + j] - s1 + 0.5 * s2;
                if (i == j) v += p.jitter;
                P[i * LD + j] = v;
            }
            __syncthreads();
            // ---- phase E: predict  a = T af,  P = T P T' + R Q R'
            if (tid < n) {
                double s = 0.0;
                for (int j = lo; j < hi; ++j) s = fma(Tm[tid * LD + j], s_af[j], s);
                s_a[tid] = s;
            }
            {
                Acc<NP> w;
                acc_zero(w);
                gemm_acc<NP, false, false>(w, Tm, P, 1.0, klo, khi, ctlo, cthi);
                acc_store<NP>(w, W, ctlo, cthi);
            }
            __syncthreads();
            {
                Acc<NP> pn;
                acc_load<NP>(pn, RQ);
                gemm_acc<NP, false, true>(pn, W, Tm, 1.0, klo, khi);
                acc_store<NP>(pn, P);
            }
            __syncthreads();
            tile_symmetrize<NP>(P, n);
            __syncthreads();
        }
        if (tid == n) {
            if (notpd) status |= GECON_ST_NOT_PD;
            if (!(fabs(ll_acc) <= 1.7e308)) status |= GECON_ST_LL_NONFINITE;
            p.ll[draw] = ll_acc;
            p.status[draw] = status;
        }
        __syncthreads();
    }
}

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
    if (!a->T || !a->R || !a->qdiag || !a->Y || !a->ll || !a->status || a->N < 0 || a->n < 1 || a->k < 0 || a->p < 1 || a->Tobs < 0 ||
        (!a->Z && !a->obs_idx)) {
        set_last_error("gecon_kalman_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (a->p > PMAX || a->p > a->n) {
        set_last_error("gecon_kalman_args: unsupported p = %d (max %d, and p <= n)", a->p, PMAX);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

template <int NP>
static int launch_kf(const gecon_kalman_args& a, cudaStream_t st) {
    const size_t smem = KfSmem<NP>::bytes(a.Tobs, a.p);
    if (smem > 227 * 1024) {
        set_last_error("observation matrix does not fit in shared memory (%zu bytes needed)", smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    int grid = 0;
    int rc = persistent_grid(kalman_ll_kernel<NP>, Cfg<NP>::NT, smem, a.N, &grid, nullptr);
    if (rc) return rc;
    kalman_ll_kernel<NP><<<grid, Cfg<NP>::NT, smem, st>>>(a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
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
    const int np = round_up8(n);
    GECON_DISPATCH_NP(np, {
        int grid = 0;
        const size_t sm = KfSmem<NP_>::bytes(Tobs, p);
        int rc = persistent_grid(kalman_ll_kernel<NP_>, Cfg<NP_>::NT, sm, 1 << 30, &grid, ctas);
        if (rc) return rc;
        *smem = (int)sm;
        *threads = Cfg<NP_>::NT;
    });
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

extern "C" int gecon_kalman_ll_batched(const gecon_kalman_args* args, void* stream) {
    int rc = check_kf_args(args);
    if (rc) return rc;
    if (args->N == 0) return 0;
    const int np = round_up8(args->n > args->k ? args->n : args->k);  // the R staging tile needs k columns
    GECON_DISPATCH_NP(np, return launch_kf<NP_>(*args, (cudaStream_t)stream));
    return 0;
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
    DevBuf dT, dR, dq, dh, dZ, dobs, dd, dY, dP0, dSin, dll, dSt, dllt;
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
    H2D(dh, hdiag, double, (a->h_stride ? N * p : p))
    H2D(dZ, Z, double, p * n)
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
