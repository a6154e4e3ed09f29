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
#pragma once
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
        return sizeof(double) * (NDBL + ny) + sizeof(int) * (NINT + Tobs) + 16;
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

// One CTA per draw.  PT = number of observables (compile time: the p x p algebra lives in registers, fully unrolled).
//
// Step t (update -> jitter -> predict), 5 barriers:
//   1  [threads < n] PZt row (selector: column pick of sym(P));  [threads n..n+PT) innovation v;
//      [warp 0] G = Zm P Zm' + Hm, F = G + jitter I, LDL' of F in registers, published to shared memory
//   2  [threads <= n] row of K = PZt F^-1 by forward/backward substitution (thread n: innovation row -> v'F^-1 v, ll_t);
//      filtered mean; M = KG/2 - PZt with KG = K G
//   3  Joseph form, expanded and symmetric by construction, one (i <= j) pair per thread:
//        P+_ij = sym(P)_ij + K_j . M_i + K_i . M_j (+ jitter on the diagonal);   predicted mean a = T a+
//   4  W = T P+   (DMMA, non-zero column range of T only)
//   5  P = R Q R' + W T'   (symmetrised lazily: every reader takes (P_ij + P_ji) / 2)
// resident CTAs per SM the register allocator must leave room for (shared memory allows about this many)
template <int NP>
constexpr int kf_min_ctas() {
    return NP <= 16 ? 8 : NP <= 24 ? 4 : NP <= 32 ? 3 : NP <= 40 ? 2 : 1;
}

template <int NP, int PT>
__global__ void __launch_bounds__(Cfg<NP>::NT, kf_min_ctas<NP>()) kalman_ll_kernel(const gecon_kalman_args p) {
    using C = Cfg<NP>;
    constexpr int LD = C::LD, NT = C::NT;
    extern __shared__ __align__(16) double sm[];
    double* Tm = sm;
    double* RQ = Tm + C::TILE;
    double* P = RQ + C::TILE;
    double* Aw = P + C::TILE;
    double* W = Aw + C::TILE;
    double* s_Y = W + C::TILE;          // [Tobs][p], 16-byte aligned (five even-sized tiles precede it)
    const size_t ny = ((size_t)p.Tobs * PT + 1) & ~(size_t)1;
    double* s_PZt = s_Y + ny;           // [NP][PS]
    double* s_K = s_PZt + NP * PS;      // [NP+1][PS]
    double* s_M = s_K + (NP + 1) * PS;  // [NP][PS]   K G / 2 - PZt
    double* s_Z = s_M + NP * PS;        // [PMAX][NP] dense design matrix (unused for selectors)
    double* s_F = s_Z + PMAX * NP;      // [PMAX][PS] strict lower part: L of F = L D L'
    double* s_G = s_F + PMAX * PS;      // [PMAX][PS] Zm P Zm' + Hm
    double* s_a = s_G + PMAX * PS;      // [NP] predicted mean
    double* s_af = s_a + NP;            // [NP] filtered mean
    double* s_q = s_af + NP;            // [NP] shock variances
    double* s_red = s_q + NP;           // [NP]
    double* s_v = s_red + NP;           // [PMAX] innovation
    double* s_w = s_v + PMAX;           // [PMAX] (spare)
    double* s_d = s_w + PMAX;
    double* s_h = s_d + PMAX;
    double* s_dinv = s_h + PMAX;        // 1 / D
    double* s_sc = s_dinv + PMAX;       // [2] det F, not-PD flag
    int* s_obs = reinterpret_cast<int*>(s_sc + 2);  // [PMAX]
    int* s_i = s_obs + PMAX;                          // [4]
    int* s_wb = s_i + 4;                              // [Tobs] bit a set <=> y[t][a] observed
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_wb + p.Tobs) + 7) & ~(uintptr_t)7);

    const int n = p.n, k = p.k, Tobs = p.Tobs;
    const bool sel = (p.obs_idx != nullptr);
    const int tid = threadIdx.x;

    // ---- stage the observations once per CTA: 1-D TMA bulk copy (16-byte granules) + plain tail
    const uint32_t ybytes = (uint32_t)((size_t)Tobs * PT * sizeof(double));
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
    if (tid < PT) s_obs[tid] = sel ? p.obs_idx[tid] : 0;
    __syncthreads();
    if (!sel) {
        for (int i = tid; i < PT * n; i += NT) {
            const int a = i / n, j = i - a * n;
            s_Z[a * NP + j] = p.Z[i];
        }
    }
    bool y_ready = false;
    int obs_r[PT];
#pragma unroll
    for (int a = 0; a < PT; ++a) obs_r[a] = s_obs[a];

    // ---- this thread's (i <= j) pairs of the symmetric covariance update: fixed for the launch
    constexpr int MAXP = (NP * (NP + 1) / 2 + NT - 1) / NT;
    int pr_i[MAXP], pr_j[MAXP];
    {
        const int npairs = n * (n + 1) / 2;
#pragma unroll
        for (int e = 0; e < MAXP; ++e) {
            const int idx = tid + e * NT;
            pr_i[e] = -1;
            pr_j[e] = 0;
            if (idx < npairs) {
                int i = 0, rem = idx;
                while (rem >= n - i) {
                    rem -= n - i;
                    ++i;
                }
                pr_i[e] = i;
                pr_j[e] = i + rem;
            }
        }
    }

    const double LOG2PI = 1.8378770664093453;
    const double ll_const = (p.mvn_const_mode == 0) ? PT * LOG2PI : LOG2PI;
    const int lyap_cap = p.lyap_max_iter > 0 ? p.lyap_max_iter : 64;
    const double jitter = p.jitter;

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
        if (tid < k && !p.qfull) {
            const double qv = p.qdiag[(size_t)draw * p.q_stride + tid];
            s_q[tid] = p.sigma_inputs ? qv * qv : qv;
        }
        if (tid < PT) {
            const double hv = (p.hdiag && (p.h_count <= 0 || (int)tid < p.h_count)) ? p.hdiag[(size_t)draw * p.h_stride + tid] : 0.0;
            s_h[tid] = p.sigma_inputs ? hv * hv : hv;
            s_d[tid] = p.d ? p.d[(size_t)draw * p.d_stride + tid] : 0.0;
        }
        if (tid < NP) s_a[tid] = 0.0;
        if (!sel && p.z_stride) {  // one design matrix per draw (the zero padding of s_Z is never overwritten)
            const double* gZ = p.Z + (size_t)draw * (size_t)p.z_stride;
            for (int i = tid; i < PT * n; i += NT) {
                const int a = i / n, j = i - a * n;
                s_Z[a * NP + j] = gZ[i];
            }
        }
        __syncthreads();
        if (p.qfull) {  // R Q R' with a full shock covariance: two DMMA products through the tiles that are still free
            tile_load<NP>(P, p.qfull + (size_t)draw * (size_t)p.qfull_stride, k, k, k);
            __syncthreads();
            Acc<NP> acc;
            acc_zero(acc);
            gemm_acc<NP, false, false>(acc, W, P, 1.0);
            acc_store<NP>(acc, Aw);
            __syncthreads();
            acc_zero(acc);
            gemm_acc<NP, false, true>(acc, Aw, W, 1.0);
            acc_store<NP>(acc, RQ);
            __syncthreads();
            tile_symmetrize<NP>(RQ, n);
        } else {
            rqr_fill<NP>(RQ, W, s_q, n, k);
        }
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
            __syncthreads();
        }
        if (!y_ready) {  // first draw of this CTA: the bulk copy has had the whole Lyapunov solve to land
            if (ybulk) mbar_wait(s_bar, 0);
            __syncthreads();
            for (int t = tid; t < Tobs; t += NT) {
                int bits = 0;
#pragma unroll
                for (int a = 0; a < PT; ++a) {
                    const double yv = s_Y[(size_t)t * PT + a];
                    if (!(yv != yv || yv == p.missing_fill)) bits |= 1 << a;
                }
                s_wb[t] = bits;
            }
            y_ready = true;
            __syncthreads();
        }

        // The predicted mean rides along in the spare column n of the P / W tiles when there is one (n < NP): phase 2
        // stores a+ in P[:, n], the W = T P+ product then leaves T a+ in W[:, n] -- no separate matrix-vector loop.
        const bool aug = (n < NP);
        const int wlo = aug ? min(ctlo, n >> 3) : ctlo, whi = aug ? max(cthi, (n >> 3) + 1) : cthi;
        if (aug && tid < NP) W[tid * LD + n] = 0.0;  // a_0 = 0
        // T's fragments are the same in all T_obs steps: keep them in registers (small NP only: register budget)
        constexpr bool CACHE = (NP <= 24);
        constexpr int KS = NP / 4, CTN = NP / 8;
        double tA[CACHE ? KS : 1], tB[CACHE ? CTN : 1][CACHE ? KS : 1], rq[CACHE ? CTN : 1][2];
        if constexpr (CACHE) {
            const int g = (tid & 31) >> 2, q = tid & 3, r = (tid >> 5) * 8 + g;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                tA[ks] = Tm[r * LD + 4 * ks + q];
#pragma unroll
                for (int ct = 0; ct < CTN; ++ct) tB[ct][ks] = Tm[(ct * 8 + g) * LD + 4 * ks + q];
            }
#pragma unroll
            for (int ct = 0; ct < CTN; ++ct) {
                rq[ct][0] = RQ[r * LD + ct * 8 + 2 * q];
                rq[ct][1] = RQ[r * LD + ct * 8 + 2 * q + 1];
            }
        }
        __syncthreads();

        // thread n accumulates sum_t (log det F_t + v' F^-1 v): the determinants are multiplied up (mantissa and
        // exponent kept apart) and ONE logarithm is taken at the end, unless per-step values are requested
        double ll_acc = 0.0, quad_acc = 0.0, detprod = 1.0;
        long long det_exp = 0;
        int n_ll_steps = 0;
        bool notpd = false;
        for (int t = 0; t < Tobs; ++t) {
            const double* y = s_Y + (size_t)t * PT;
            const int wb = s_wb[t];
            double pz[PT];
            // ---- phase 1a: PZt = sym(P) Zm' (one row per thread) and the innovation
            if (tid < n) {
#pragma unroll
                for (int a = 0; a < PT; ++a) {
                    double s;
                    if (sel) {
                        s = 0.5 * (P[obs_r[a] * LD + tid] + P[tid * LD + obs_r[a]]);
                    } else {
                        s = 0.0;
                        for (int j = 0; j < n; ++j) s = fma(0.5 * (P[j * LD + tid] + P[tid * LD + j]), s_Z[a * NP + j], s);
                    }
                    pz[a] = ((wb >> a) & 1) ? s : 0.0;
                    s_PZt[tid * PS + a] = pz[a];
                }
            } else if (tid < n + PT) {
                const int a = tid - n;
                const bool obs = (wb >> a) & 1;
                double za;
                if (sel) {
                    za = aug ? W[s_obs[a] * LD + n] : s_a[s_obs[a]];
                } else {
                    za = 0.0;
                    for (int j = 0; j < n; ++j) za = fma(s_Z[a * NP + j], aug ? W[j * LD + n] : s_a[j], za);
                }
                s_v[a] = (obs ? y[a] : 0.0) - (((obs || !p.mask_intercept) ? s_d[a] : 0.0) + (obs ? za : 0.0));
            }
            if (!sel) __syncthreads();  // dense Z: G below is formed from PZt
            // ---- phase 1b (warp 0): G, F = G + jitter I and its L D L' factorisation, all in registers
            if (tid < 32) {
                double f[PT][PT];
#pragma unroll
                for (int a = 0; a < PT; ++a) {
#pragma unroll
                    for (int b = 0; b <= a; ++b) {
                        double s;
                        if (sel) {
                            s = 0.5 * (P[obs_r[a] * LD + obs_r[b]] + P[obs_r[b] * LD + obs_r[a]]);
                            s = ((wb >> b) & 1) ? s : 0.0;
                        } else {
                            s = 0.0;
                            for (int j = 0; j < n; ++j) s = fma(s_Z[a * NP + j], s_PZt[j * PS + b], s);
                        }
                        s = ((wb >> a) & 1) ? s : 0.0;
                        if (a == b) s += ((wb >> a) & 1) ? s_h[a] : 0.0;
                        if (tid == 0) {
                            s_G[a * PS + b] = s;
                            s_G[b * PS + a] = s;
                        }
                        f[a][b] = (a == b) ? s + jitter : s;
                    }
                }
                double det = 1.0;
                bool bad = false;
#pragma unroll
                for (int c = 0; c < PT; ++c) {
                    const double dc = f[c][c];
                    bad = bad || !(dc > 0.0);
                    det *= dc;
                    const double inv = 1.0 / dc;
                    if (tid == 0) s_dinv[c] = inv;
                    double u[PT];
#pragma unroll
                    for (int a = c + 1; a < PT; ++a) u[a] = f[a][c];
#pragma unroll
                    for (int a = c + 1; a < PT; ++a) {
                        const double l = u[a] * inv;
#pragma unroll
                        for (int b = c + 1; b <= a; ++b) f[a][b] = fma(-l, u[b], f[a][b]);
                        f[a][c] = l;
                        if (tid == 0) s_F[a * PS + c] = l;
                    }
                }
                if (tid == 0) {
                    s_sc[0] = det;
                    s_sc[1] = bad ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            // ---- phase 2: rows of K = PZt F^-1 (thread n: the innovation row), filtered mean, M = K G / 2 - PZt
            if (tid <= n) {
                double x[PT], dv[PT];
#pragma unroll
                for (int a = 0; a < PT; ++a) {
                    x[a] = (tid < n) ? pz[a] : s_v[a];
                    dv[a] = s_dinv[a];
                }
#pragma unroll
                for (int a = 1; a < PT; ++a) {
#pragma unroll
                    for (int b = 0; b < a; ++b) x[a] = fma(-s_F[a * PS + b], x[b], x[a]);
                }
                if (tid == n) {
                    double quad = 0.0;
#pragma unroll
                    for (int a = 0; a < PT; ++a) quad = fma(x[a] * dv[a], x[a], quad);
                    if (s_sc[1] != 0.0) notpd = true;
                    if (p.ll_t) {
                        const double llt = (wb == 0) ? 0.0 : -0.5 * (ll_const + log(s_sc[0]) + quad);
                        ll_acc += llt;
                        p.ll_t[(size_t)draw * Tobs + t] = llt;
                    } else if (wb != 0) {
                        int e;
                        detprod *= frexp(s_sc[0], &e);
                        det_exp += e;
                        quad_acc += quad;
                        ++n_ll_steps;
                        if ((t & 31) == 31) {
                            detprod = frexp(detprod, &e);
                            det_exp += e;
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < PT; ++a) x[a] *= dv[a];
#pragma unroll
                    for (int a = PT - 2; a >= 0; --a) {
#pragma unroll
                        for (int b = a + 1; b < PT; ++b) x[a] = fma(-s_F[b * PS + a], x[b], x[a]);
                    }
                    double af = aug ? W[tid * LD + n] : s_a[tid];
#pragma unroll
                    for (int a = 0; a < PT; ++a) af = fma(x[a], s_v[a], af);
                    if (aug) P[tid * LD + n] = af;
                    else s_af[tid] = af;
#pragma unroll
                    for (int a = 0; a < PT; ++a) {
                        double kg = 0.0;
#pragma unroll
                        for (int b = 0; b < PT; ++b) kg = fma(x[b], s_G[b * PS + a], kg);
                        s_K[tid * PS + a] = x[a];
                        s_M[tid * PS + a] = fma(0.5, kg, -pz[a]);
                    }
                }
            }
            __syncthreads();
            // ---- phase 3: Joseph-form covariance, expanded:  (I-KZ) P (I-KZ)' + K H K' = P - K PZt' - PZt K' + K G K'
            //      = P + K M' + M K'  with M = K G / 2 - PZt;  evaluated once per (i <= j) pair and mirrored, + jitter I
#pragma unroll
            for (int e = 0; e < MAXP; ++e) {
                const int i = pr_i[e], j = pr_j[e];
                if (i >= 0) {
                    double acc = (i == j) ? P[i * LD + i] + jitter : 0.5 * (P[i * LD + j] + P[j * LD + i]);
#pragma unroll
                    for (int a = 0; a < PT; ++a) {
                        acc = fma(s_K[j * PS + a], s_M[i * PS + a], acc);
                        acc = fma(s_K[i * PS + a], s_M[j * PS + a], acc);
                    }
                    P[i * LD + j] = acc;
                    P[j * LD + i] = acc;
                }
            }
            if (!aug && tid < n) {  // no spare column: predicted mean a = T a+ by a plain loop
                double s = 0.0;
                for (int j = lo; j < hi; ++j) s = fma(Tm[tid * LD + j], s_af[j], s);
                s_a[tid] = s;
            }
            __syncthreads();
            // ---- phases 4, 5: W = T [P+ | a+],  P = T P+ T' + R Q R'
            if constexpr (CACHE) {
                const int g = (tid & 31) >> 2, q = tid & 3;
                Acc<NP> w;
                acc_zero(w);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    if (4 * ks >= klo && 4 * ks < khi) {
#pragma unroll
                        for (int ct = 0; ct < CTN; ++ct) {
                            if (ct >= wlo && ct < whi) dmma884(w.v[ct][0], w.v[ct][1], tA[ks], P[(4 * ks + q) * LD + ct * 8 + g]);
                        }
                    }
                }
                acc_store<NP>(w, W, wlo, whi);
                __syncthreads();
                const int r = (tid >> 5) * 8 + g;
                Acc<NP> pn;
#pragma unroll
                for (int ct = 0; ct < CTN; ++ct) {
                    pn.v[ct][0] = rq[ct][0];
                    pn.v[ct][1] = rq[ct][1];
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    if (4 * ks >= klo && 4 * ks < khi) {
                        const double a = W[r * LD + 4 * ks + q];
#pragma unroll
                        for (int ct = 0; ct < CTN; ++ct) dmma884(pn.v[ct][0], pn.v[ct][1], a, tB[ct][ks]);
                    }
                }
                acc_store<NP>(pn, P);
            } else {
                {
                    Acc<NP> w;
                    acc_zero(w);
                    gemm_acc<NP, false, false>(w, Tm, P, 1.0, klo, khi, wlo, whi);
                    acc_store<NP>(w, W, wlo, whi);
                }
                __syncthreads();
                Acc<NP> pn;
                acc_load<NP>(pn, RQ);
                gemm_acc<NP, false, true>(pn, W, Tm, 1.0, klo, khi);
                acc_store<NP>(pn, P);
            }
            __syncthreads();
        }
        if (tid == n) {
            if (!p.ll_t) ll_acc = -0.5 * (n_ll_steps * ll_const + (log(detprod) + (double)det_exp * 0.6931471805599453) + quad_acc);
            if (notpd) status |= GECON_ST_NOT_PD;
            if (!(fabs(ll_acc) <= 1.7e308)) status |= GECON_ST_LL_NONFINITE;
            p.ll[draw] = ll_acc;
            p.status[draw] = status;
        }
        __syncthreads();
    }
}

}  // namespace gecon
