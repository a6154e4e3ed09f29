// Kalman-filter log-likelihood, ONE WARP PER DRAW (padded NP <= 32, i.e. filter dimension <= 31; selector Z).
//
// Same semantics as kalman_ll_kernel (kalman.cuh; pymc_extras StandardFilter as called from
// gEconpy/model/statespace.py:1151-1157, restated in oracle/statespace.py): update -> jitter -> predict, Joseph-form
// covariance update, missing observations masked, a0 = 0, P0 = dlyap(T, R Q R') by Smith doubling.
//
// Why a second kernel: with n <= 31 the filter step is a chain of tiny dependent phases, and in the CTA-per-draw kernel
// five CTA barriers per step leave the SM idle most of the time.  Here a warp owns a draw, so the only synchronisation
// is __syncwarp, and 16 draws per SM advance independently.  Per step:
//   1  every lane redundantly: F = Z P Z' + H + jitter I, its L D L' factorisation, the innovation and its quadratic
//      form (p x p, registers);  lane i: row i of P Z', of K = P Z' F^-1, of N = -(P Z' + jitter K), filtered mean a+_i
//   2  P+ = P + [K | a+] [N | e_n]'  as ONE DMMA product with k = p + 1 on the accumulators that still hold P from the
//      previous step.  This IS the Joseph form: with G = Z P Z' + H = F - jitter I and K = P Z' F^-1,
//      K G = P Z' - jitter K and K (P Z')' = P Z' F^-1 (P Z')' is symmetric, so
//      (I - K Z) P (I - K Z)' + K H K' = P - K (P Z')' - (P Z') K' + K G K' = P - K (P Z' + jitter K)' = P + K N'
//      (rank p instead of the rank 2p of the expanded form P + K M' + M K', M = K G / 2 - P Z', used until round 2).
//      The extra column drops a+ into the spare column n of P+, so the next product also propagates the mean; the
//      jitter on the diagonal of P+ is carried by the constant term of phase 4 (R Q R' + jitter T T')
//   3  W = T [P+ | a+]      (DMMA; T fragments live in registers for the whole draw)
//   4  P = R Q R' + W T'    (DMMA; R Q R' in registers for NP <= 16)
// P is symmetric, so phases 2 and 4 only compute and store the tiles on or above the block diagonal (3 of 4 tiles at NP = 16,
// 6 of 9 at NP = 24) and the readers of phases 1 and 3 take mirror images for the rest.
// P, W and the [K | M] panel go through the warp's private shared-memory tiles only to change fragment layout
// (accumulator -> A/B operand), four __syncwarp per step.  P is not symmetrised explicitly: the update term is
// symmetric, and the rounding-level antisymmetric part of W T' is contracted by the next T . T' (rho(T) < 1).
#pragma once
#include <type_traits>

#include "kalman.cuh"

// Per-configuration build (kalman_spec.cu, compiled on demand for one (filter dimension, observables) pair): the filter dimension is
// the compile-time constant GECON_KW_SPEC_N, so the k-steps of the two products that only multiply padding disappear instead of being
// issued predicated-off (7 of 31 DMMA per step at n = 10, NP = 16: measured 40.9 -> 36.0 ms per 262,144 draws, T_obs = 200).  The
// kernel gets its own name there: it must never be confused with the generic instantiation of the core library.
#ifdef GECON_KW_SPEC_N
#define kalman_ll_warp_kernel kalman_ll_warp_spec_kernel
#endif

namespace gecon {

template <int PT>
struct KmCfg {
    static constexpr int KS2 = (PT + 1 + 3) / 4;                // k-steps of the rank-(p+1) update
    static constexpr int KW = 4 * KS2;                          // rows of the (transposed) panels [K|a+]', [N|e_n]'
};

template <int NP, int PT, int WPC_ = 4>
struct KwSmem {
    static constexpr int WPC = WPC_;  // warps (= draws in flight) per CTA: 4 (several CTAs per SM) or 16 (one CTA per SM, one copy of Y)
    static constexpr int TILE = Cfg<NP>::TILE;
    static constexpr int KM2 = 2 * KmCfg<PT>::KW * Cfg<NP>::LD;    // KM and MK panels, stored transposed: [k][LD]
    static constexpr int ALIAS = KM2 > TILE ? KM2 : TILE;          // the Lyapunov scratch tile shares their storage
    static constexpr bool C0_REGS = (NP <= 8);   // R Q R' + jitter T T' in registers, else in a shared-memory tile
    static constexpr int PER_WARP = 2 * TILE + ALIAS + (C0_REGS ? 0 : TILE) + NP;  // doubles (even)
    static size_t bytes(int Tobs) {
        const size_t ny = ((size_t)Tobs * PT + 1) & ~(size_t)1;
        return sizeof(double) * (ny + (size_t)WPC * PER_WARP) + sizeof(int) * ((size_t)Tobs + 4) + 16;
    }
};

// exact max of non-negative doubles across the warp (NaN-propagating: NaN patterns order above +inf)
__device__ __forceinline__ double warp_max_nonneg(double v) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v));
    const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, (khi == mhi) ? klo : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

template <int NP>
struct WAcc {
    double v[NP / 8][NP / 8][2];  // [row strip][column tile][2]
};

template <int NP>
__device__ __forceinline__ void wacc_zero(WAcc<NP>& a) {
#pragma unroll
    for (int s = 0; s < NP / 8; ++s)
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct) a.v[s][ct][0] = a.v[s][ct][1] = 0.0;
}
template <int NP>
__device__ __forceinline__ void wacc_load(WAcc<NP>& a, const double* __restrict__ M, int lane) {
    constexpr int LD = Cfg<NP>::LD;
    const double* base = M + (lane >> 2) * LD + 2 * (lane & 3);
#pragma unroll
    for (int s = 0; s < NP / 8; ++s)
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct) {
            const double2 t = *reinterpret_cast<const double2*>(base + s * 8 * LD + ct * 8);
            a.v[s][ct][0] = t.x;
            a.v[s][ct][1] = t.y;
        }
}
template <int NP>
__device__ __forceinline__ void wacc_store(const WAcc<NP>& a, double* __restrict__ M, int lane) {
    constexpr int LD = Cfg<NP>::LD;
    double* base = M + (lane >> 2) * LD + 2 * (lane & 3);
#pragma unroll
    for (int s = 0; s < NP / 8; ++s)
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct)
            *reinterpret_cast<double2*>(base + s * 8 * LD + ct * 8) = make_double2(a.v[s][ct][0], a.v[s][ct][1]);
}

// The covariance is symmetric: inside the filter loop only the tiles on or above the block diagonal (ct >= s) are computed
// and stored; a reader that needs an element of a lower tile takes its mirror image from the upper one.
template <int NP>
__device__ __forceinline__ void wacc_load_upper(WAcc<NP>& a, const double* __restrict__ M, int lane) {
    constexpr int LD = Cfg<NP>::LD;
    const double* base = M + (lane >> 2) * LD + 2 * (lane & 3);
#pragma unroll
    for (int s = 0; s < NP / 8; ++s)
#pragma unroll
        for (int ct = s; ct < NP / 8; ++ct) {
            const double2 t = *reinterpret_cast<const double2*>(base + s * 8 * LD + ct * 8);
            a.v[s][ct][0] = t.x;
            a.v[s][ct][1] = t.y;
        }
}
template <int NP>
__device__ __forceinline__ void wacc_store_upper(const WAcc<NP>& a, double* __restrict__ M, int lane) {
    constexpr int LD = Cfg<NP>::LD;
    double* base = M + (lane >> 2) * LD + 2 * (lane & 3);
#pragma unroll
    for (int s = 0; s < NP / 8; ++s)
#pragma unroll
        for (int ct = s; ct < NP / 8; ++ct)
            *reinterpret_cast<double2*>(base + s * 8 * LD + ct * 8) = make_double2(a.v[s][ct][0], a.v[s][ct][1]);
}
// offset of element (r, c) of a symmetric tile whose lower block triangle is not maintained
template <int NP>
__device__ __forceinline__ int sym_off(int r, int c) {
    constexpr int LD = Cfg<NP>::LD;
    return ((r >> 3) > (c >> 3)) ? c * LD + r : r * LD + c;
}

// acc += A * op(B) for one warp, operands in shared-memory tiles; k-steps [0, nks).  TB: acc += A * B_s'.
template <int NP, bool TB>
__device__ __forceinline__ void wgemm(WAcc<NP>& acc, const double* __restrict__ A, const double* __restrict__ B, int nks, int lane) {
    constexpr int LD = Cfg<NP>::LD, NS = NP / 8;
    const int g = lane >> 2, q = lane & 3;
    for (int ks = 0; ks < nks; ++ks) {
        double a[NS], b[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            a[s] = A[(8 * s + g) * LD + 4 * ks + q];
            b[s] = TB ? B[(8 * s + g) * LD + 4 * ks + q] : B[(4 * ks + q) * LD + 8 * s + g];
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int ct = 0; ct < NS; ++ct) dmma884(acc.v[s][ct][0], acc.v[s][ct][1], a[s], b[ct]);
    }
}

template <int NP>
__device__ __forceinline__ void warp_tile_load(double* __restrict__ dst, const double* __restrict__ src, int rows, int cols, int lane) {
    constexpr int LD = Cfg<NP>::LD;
    for (int i = lane; i < Cfg<NP>::TILE; i += 32) {
        const int r = i / LD, c = i - r * LD;
        dst[i] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.0;
    }
}

template <int NP, int PT, int MINB, int WPC_ = 4>
__global__ void __launch_bounds__(WPC_ * 32, MINB) kalman_ll_warp_kernel(const gecon_kalman_args p) {
    using S = KwSmem<NP, PT, WPC_>;
    constexpr int LD = Cfg<NP>::LD, TILE = Cfg<NP>::TILE, NS = NP / 8, KSN = NP / 4;
    constexpr int KS2 = KmCfg<PT>::KS2, KW = KmCfg<PT>::KW, WPC = S::WPC;
    constexpr bool C0_REGS = S::C0_REGS;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
#ifdef GECON_KW_SPEC_N
    constexpr int n = GECON_KW_SPEC_N;  // per-configuration build: the k-step counts of the products are compile-time
    const int k = p.k, Tobs = p.Tobs;
#else
    const int n = p.n, k = p.k, Tobs = p.Tobs;
#endif
    const size_t ny = ((size_t)Tobs * PT + 1) & ~(size_t)1;
    double* s_Y = sm;
    double* wb0 = sm + ny + (size_t)warp * S::PER_WARP;
    double* P = wb0;
    double* W = P + TILE;
    double* Aw = W + TILE;          // Lyapunov scratch; afterwards the same storage holds the two panels
    double* KM = Aw;                // [KW][LD]  rows: K' | a+'    (transposed: lane i writes column i, conflict-free)
    double* MK = KM + KW * LD;      // [KW][LD]  rows: N' | e_n'
    double* C0t = Aw + S::ALIAS;    // R Q R' (only when it does not live in registers)
    double* s_q = C0t + (C0_REGS ? 0 : TILE);
    int* s_wb = reinterpret_cast<int*>(sm + ny + (size_t)WPC * S::PER_WARP);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_wb + Tobs) + 7) & ~(uintptr_t)7);

    // ---- stage the observations once per CTA: 1-D TMA bulk copy (16-byte granules) + plain tail, then the missing masks
    bool any_missing;
    {
        const uint32_t ybytes = (uint32_t)((size_t)Tobs * PT * sizeof(double));
        const uint32_t ybulk = ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) ? (ybytes & ~15u) : 0u;
        if (tid == 0) mbar_init(s_bar, 1);
        __syncthreads();
        if (tid == 0 && ybulk) {
            mbar_expect_tx(s_bar, ybulk);
            tma_bulk_g2s(s_Y, p.Y, ybulk, s_bar);
        }
        for (uint32_t i = ybulk / 8 + tid; i < ybytes / 8; i += WPC * 32) s_Y[i] = p.Y[i];
        if (ybulk) mbar_wait(s_bar, 0);
        __syncthreads();
        bool complete = true;
        for (int t = tid; t < Tobs; t += WPC * 32) {
            int bits = 0;
#pragma unroll
            for (int a = 0; a < PT; ++a) {
                const double yv = s_Y[(size_t)t * PT + a];
                if (!(yv != yv || yv == p.missing_fill)) bits |= 1 << a;
            }
            s_wb[t] = bits;
            complete = complete && (bits == (1 << PT) - 1);
        }
        any_missing = !__syncthreads_and(complete);  // CTA-uniform: a complete sample runs the step without its masks
        // missing entries -> 0 in the staged copy (their mask bit is what the filter looks at): the step loop reads y without a select
        for (uint32_t i = tid; i < ybytes / 8; i += WPC * 32) {
            const uint32_t t = i / PT, a = i - t * PT;
            if (!((s_wb[t] >> a) & 1)) s_Y[i] = 0.0;
        }
        __syncthreads();
    }
    int obs_r[PT];
#pragma unroll
    for (int a = 0; a < PT; ++a) obs_r[a] = p.obs_idx[a];

    const double LOG2PI = 1.8378770664093453;
    const double ll_const = (p.mvn_const_mode == 0) ? PT * LOG2PI : LOG2PI;
    const int lyap_cap = p.lyap_max_iter > 0 ? p.lyap_max_iter : 64;
    const double jitter = p.jitter;
    const bool keep_d = (p.mask_intercept == 0);  // the intercept is NOT masked at missing entries (pymc_extras: d + Z_masked a)
    const int nks = (n + 3) >> 2;
    // k-steps of the two products with T: only its first t_cols columns can be non-zero when the caller says so (gecon_kalman_args.t_cols)
#if defined(GECON_KW_SPEC_N) && defined(GECON_KW_SPEC_TC) && GECON_KW_SPEC_TC > 0
    constexpr int nks_t = (GECON_KW_SPEC_TC + 3) >> 2;
#else
    const int nks_t = (p.t_cols > 0 && p.t_cols < n) ? ((p.t_cols + 3) >> 2) : nks;
#endif
    const int il = lane < NP ? lane : 0;  // row handled by this lane in phase 1 (clamped: lanes >= n compute on row 0 and discard)
    const bool rowlane = lane < n;

    for (long long draw = (long long)blockIdx.x * WPC + warp; draw < p.N; draw += (long long)gridDim.x * WPC) {
        int status = p.status_in ? (p.status_in[draw] & ~GECON_ST_BK_CERTIFIED) : 0;
        if (status & p.gate_mask) {  // warp-uniform
            if (lane == 0) {
                p.ll[draw] = -INFINITY;
                p.status[draw] = status | GECON_ST_SKIPPED;
            }
            if (p.ll_t) {
                for (int t = lane; t < Tobs; t += 32) p.ll_t[(size_t)draw * Tobs + t] = -INFINITY;
            }
            continue;
        }
        const double* gT = p.T + (size_t)draw * n * n;
        // ---- T -> Aw, R -> W, variances; P = R Q R' (mirrored pairs: exactly symmetric)
        warp_tile_load<NP>(Aw, gT, n, n, lane);
        warp_tile_load<NP>(W, p.R + (size_t)draw * n * k, n, k, lane);
        if (lane < k) {
            const double qv = p.qdiag[(size_t)draw * p.q_stride + lane];
            s_q[lane] = p.sigma_inputs ? qv * qv : qv;
        }
        double hv[PT], dv0[PT];
#pragma unroll
        for (int a = 0; a < PT; ++a) {
            const double h = (p.hdiag && (p.h_count <= 0 || a < p.h_count)) ? p.hdiag[(size_t)draw * p.h_stride + a] : 0.0;
            hv[a] = p.sigma_inputs ? h * h : h;
            dv0[a] = p.d ? p.d[(size_t)draw * p.d_stride + a] : 0.0;
        }
        __syncwarp();
        for (int idx = lane; idx < TILE; idx += 32) {
            const int i = idx / LD, j = idx - i * LD;
            double s = 0.0;
            if (i < n && j < n) {
                const int lo = min(i, j), hi = max(i, j);
                for (int c = 0; c < k; ++c) s = fma(W[lo * LD + c] * s_q[c], W[hi * LD + c], s);
            }
            P[idx] = s;
        }
        __syncwarp();
        // c0 = R Q R' + jitter T T': the jitter added to the diagonal of every filtered covariance, carried through the
        // prediction once per draw instead of once per step  (T (P+ + jitter I) T' = T P+ T' + jitter T T')
        WAcc<NP> pacc, c0;
        wacc_load<NP>(pacc, P, lane);
        {
            WAcc<NP> tt;
            wacc_zero(tt);
            wgemm<NP, true>(tt, Aw, Aw, nks, lane);
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ct = 0; ct < NS; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) tt.v[s][ct][e] = fma(jitter, tt.v[s][ct][e], pacc.v[s][ct][e]);
            if constexpr (C0_REGS) c0 = tt;
            else wacc_store<NP>(tt, C0t, lane);
        }

        // ---- P0
        if (p.P0) {
            warp_tile_load<NP>(P, p.P0 + (size_t)draw * n * n, n, n, lane);
            __syncwarp();
            wacc_load<NP>(pacc, P, lane);
        } else {
            // Smith doubling: P <- P + A_j P A_j', A_{j+1} = A_j^2, until max|increment| <= 1e-16 max|P|
            int it = 0;
            bool done = false;
            while (it < lyap_cap) {
                ++it;
                WAcc<NP> w;
                wacc_zero(w);
                wgemm<NP, false>(w, Aw, P, nks, lane);
                wacc_store<NP>(w, W, lane);
                __syncwarp();
                WAcc<NP> d;
                wacc_zero(d);
                wgemm<NP, true>(d, W, Aw, nks, lane);
                wacc_zero(w);
                wgemm<NP, false>(w, Aw, Aw, nks, lane);  // A^2
                double dmax = 0.0, pmax = 0.0;
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int ct = 0; ct < NS; ++ct)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            pacc.v[s][ct][e] += d.v[s][ct][e];
                            const double ad = fabs(d.v[s][ct][e]), ap = fabs(pacc.v[s][ct][e]);
                            if (ad > dmax || ad != ad) dmax = ad;
                            if (ap > pmax || ap != ap) pmax = ap;
                        }
                dmax = warp_max_nonneg(dmax);
                pmax = warp_max_nonneg(pmax);
                __syncwarp();  // every lane is done reading P, Aw, W
                wacc_store<NP>(pacc, P, lane);
                if (dmax != dmax || pmax != pmax) break;
                if (dmax <= 1e-16 * pmax) {
                    done = true;
                    break;
                }
                wacc_store<NP>(w, Aw, lane);
                __syncwarp();
            }
            if (!done) status |= GECON_ST_LYAP;
            __syncwarp();
        }
        // ---- T fragments (registers for the whole draw); panels zeroed (they alias the Lyapunov scratch); a0 = 0
        double tA[NS][KSN];
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int ks = 0; ks < KSN; ++ks) {
                const int r = 8 * s + g, c = 4 * ks + q;
                tA[s][ks] = (r < n && c < n) ? gT[(size_t)r * n + c] : 0.0;
            }
        for (int i = lane; i < 2 * KW * LD; i += 32) KM[i] = 0.0;
        if (lane < NP) W[lane * LD + n] = 0.0;
        __syncwarp();
        if (lane == 0) MK[PT * LD + n] = 1.0;
        __syncwarp();

        double ll_acc = 0.0, quad_acc = 0.0, detprod = 1.0;
        long long det_exp = 0;
        int n_ll_steps = 0;
        bool notpd = false;
        // One filter step.  MISS (compile-time tag): the sample has missing observations and every use of an observable goes through
        // its mask bit; a complete sample (decided once per CTA at staging) runs the same statements with the masks folded to "observed"
        // -- a quarter of the step's instructions (ncu r02: the selects of phase 1) -- and, bit for bit, the same arithmetic.
        auto filter_step = [&](auto miss_tag, const int t) {
            constexpr bool MISS = decltype(miss_tag)::value;
            const double* y = s_Y + (size_t)t * PT;
            const int wb = MISS ? s_wb[t] : (1 << PT) - 1;
            // ---- phase 1: P Z' row, innovation, F = G + jitter I (lower triangle), all from the shared-memory copy of P
            double pz[PT], v[PT], f[PT][PT];
#pragma unroll
            for (int a = 0; a < PT; ++a) {
                const bool ob = (wb >> a) & 1;
                const double s = P[sym_off<NP>(obs_r[a], il)];
                pz[a] = ob ? s : 0.0;
                const double za = W[obs_r[a] * LD + n];
                v[a] = (y[a] - ((ob || keep_d) ? dv0[a] : 0.0)) - (ob ? za : 0.0);  // y is 0 where missing
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    double x = P[sym_off<NP>(obs_r[a], obs_r[b])];
                    x = (ob && ((wb >> b) & 1)) ? x : 0.0;
                    if (a == b) x += (ob ? hv[a] : 0.0) + jitter;
                    f[a][b] = x;
                }
            }
            // L D L' of F (no square roots), log det F = log prod d
            double dinv[PT], det = 1.0;
            bool bad = false;
#pragma unroll
            for (int c = 0; c < PT; ++c) {
                const double dc = f[c][c];
                bad = bad || !(dc > 0.0);
                det *= dc;
                const double inv = rcp_nr(dc);
                dinv[c] = inv;
                double u[PT];
#pragma unroll
                for (int a = c + 1; a < PT; ++a) u[a] = f[a][c];
#pragma unroll
                for (int a = c + 1; a < PT; ++a) {
                    const double l = u[a] * inv;
#pragma unroll
                    for (int b = c + 1; b <= a; ++b) f[a][b] = fma(-l, u[b], f[a][b]);
                    f[a][c] = l;
                }
            }
            notpd = notpd || bad;
            // innovation: v' F^-1 v
            {
                double x[PT];
#pragma unroll
                for (int a = 0; a < PT; ++a) x[a] = v[a];
#pragma unroll
                for (int a = 1; a < PT; ++a)
#pragma unroll
                    for (int b = 0; b < a; ++b) x[a] = fma(-f[a][b], x[b], x[a]);
                double quad = 0.0;
#pragma unroll
                for (int a = 0; a < PT; ++a) quad = fma(x[a] * dinv[a], x[a], quad);
                if (p.ll_t) {
                    const double llt = (wb == 0) ? 0.0 : -0.5 * (ll_const + log(det) + quad);
                    ll_acc += llt;
                    if (lane == 0) p.ll_t[(size_t)draw * Tobs + t] = llt;
                } else if (wb != 0) {
                    // det = m 2^e, m in [0.5, 1) (det > 0 and normal, else the draw is flagged not-PD anyway)
                    const int hi = __double2hiint(det);
                    det_exp += ((hi >> 20) & 0x7ff) - 1022;
                    detprod *= __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(det));
                    quad_acc += quad;
                    ++n_ll_steps;
                    if ((t & 31) == 31) {
                        int e;
                        detprod = frexp(detprod, &e);
                        det_exp += e;
                    }
                }
            }
            // ---- phase 2: row of K = P Z' F^-1, filtered mean, N = -(P Z' + jitter K)
            {
                double x[PT];
#pragma unroll
                for (int a = 0; a < PT; ++a) x[a] = pz[a];
#pragma unroll
                for (int a = 1; a < PT; ++a)
#pragma unroll
                    for (int b = 0; b < a; ++b) x[a] = fma(-f[a][b], x[b], x[a]);
#pragma unroll
                for (int a = 0; a < PT; ++a) x[a] *= dinv[a];
#pragma unroll
                for (int a = PT - 2; a >= 0; --a)
#pragma unroll
                    for (int b = a + 1; b < PT; ++b) x[a] = fma(-f[b][a], x[b], x[a]);
                double af = W[il * LD + n];
#pragma unroll
                for (int a = 0; a < PT; ++a) af = fma(x[a], v[a], af);
                if (rowlane) {
                    double* km = KM + lane;
                    double* mk = MK + lane;
#pragma unroll
                    for (int a = 0; a < PT; ++a) {
                        km[a * LD] = x[a];
                        mk[a * LD] = -fma(jitter, x[a], pz[a]);
                    }
                    km[PT * LD] = af;
                }
            }
            __syncwarp();
            // ---- phase 3: P+ = P + [K | a+] [N | e_n]' (the jitter on the diagonal: see c0), accumulators -> P tile
#pragma unroll
            for (int ks = 0; ks < KS2; ++ks) {
                double a[NS], b[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    a[s] = KM[(4 * ks + q) * LD + 8 * s + g];
                    b[s] = MK[(4 * ks + q) * LD + 8 * s + g];
                }
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int ct = s; ct < NS; ++ct) dmma884(pacc.v[s][ct][0], pacc.v[s][ct][1], a[s], b[ct]);
            }
            wacc_store_upper<NP>(pacc, P, lane);
            __syncwarp();
            // ---- phase 4: W = T [P+ | a+]   (rows of P+ below the block diagonal are read as mirror images)
            WAcc<NP> w;
            wacc_zero(w);
#pragma unroll
            for (int ks = 0; ks < KSN; ++ks) {
                if (ks < nks_t) {
                    double b[NS];
#pragma unroll
                    for (int ct = 0; ct < NS; ++ct) b[ct] = ((ks >> 1) > ct) ? P[(8 * ct + g) * LD + 4 * ks + q] : P[(4 * ks + q) * LD + 8 * ct + g];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int ct = 0; ct < NS; ++ct) dmma884(w.v[s][ct][0], w.v[s][ct][1], tA[s][ks], b[ct]);
                }
            }
            wacc_store<NP>(w, W, lane);
            __syncwarp();
            // ---- phase 5: P = R Q R' + W T'
            if constexpr (C0_REGS) pacc = c0;
            else wacc_load_upper<NP>(pacc, C0t, lane);
#pragma unroll
            for (int ks = 0; ks < KSN; ++ks) {
                if (ks < nks_t) {
                    double a[NS];
#pragma unroll
                    for (int s = 0; s < NS; ++s) a[s] = W[(8 * s + g) * LD + 4 * ks + q];
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int ct = s; ct < NS; ++ct) dmma884(pacc.v[s][ct][0], pacc.v[s][ct][1], a[s], tA[ct][ks]);
                }
            }
            wacc_store_upper<NP>(pacc, P, lane);
            __syncwarp();
        };
        if (any_missing) {
            for (int t = 0; t < Tobs; ++t) filter_step(std::true_type{}, t);
        } else {
            for (int t = 0; t < Tobs; ++t) filter_step(std::false_type{}, t);
        }
        if (!p.ll_t) ll_acc = -0.5 * (n_ll_steps * ll_const + (log(detprod) + (double)det_exp * 0.6931471805599453) + quad_acc);
        if (notpd) status |= GECON_ST_NOT_PD;
        if (!(fabs(ll_acc) <= 1.7e308)) status |= GECON_ST_LL_NONFINITE;
        if (lane == 0) {
            p.ll[draw] = ll_acc;
            p.status[draw] = status;
        }
        __syncwarp();
    }
}

}  // namespace gecon
