// Batched cycle reduction + selection matrix + policy residual + Blanchard-Kahn certificate, ONE WARP PER DRAW
// (padded dimension NP <= 32).  Same semantics as cr_solve_kernel (cr_solve.cu): gEconpy/solvers/cycle_reduction.py:127-183
// (_cycle_reduction_core), gEconpy/solvers/shared.py:74-75 (R = -(C T + B)^-1 D), gEconpy/model/statespace.py:213 (residual).
//
// Why a second kernel (round-1 ncu: barrier stall 3.5 warps per issue, warps active 22 %, DMMA pipe 20 %): in the CTA-per-draw
// kernel one warp runs the 8 dependent pivot steps of every panel while the others wait at a CTA barrier.  Here a warp owns a
// draw from the tile loads to the certificate, the only synchronisation is __syncwarp, and 10-16 draws per SM advance
// independently, so one draw's pivot chain overlaps the tensor-path updates of the others.
//
// Structure used (exact, never approximated): A = d/dx_{t-1} is non-zero only in the lag columns [lag_lo, lag_hi), C = d/dx_{t+1}
// only in the lead columns [lead_lo, lead_hi) (contiguous in the reference's solver order, perturbation.py:130-158), and the
// iteration preserves both ranges: A0, X0 = A1^-1 A0 and the accumulated correction of A1hat live in lag columns, A2 and
// X2 in lead columns.  They are stored PACKED (column c of the range at packed column c - lo, lo rounded down to an even
// number) in tiles of C column tiles, LDC = 8 C + 4:
//   shared memory per warp   A1 (NP x LD), W (NP x LD, Gauss-Jordan workspace), X0 (NP x LDC), X2 (NP x LDC)
//   registers                the A-operand fragments of A0 and A2 (loaded before the in-place solve overwrites them with
//                            X0, X2), the accumulated correction H = sum A2 X0 (A1hat = B - H)
// Per iteration: [X0 | X2] = A1^-1 [A0 | A2] by blocked Gauss-Jordan (8-column panels, lane = row, pivots by redux.sync,
// trailing updates as DMMA products with k = 8; the first block step reads A1 and writes W, so A1 needs no copy); four
// DMMA products from the register fragments through the pivot-row map; A1 -= A0 X2 + A2 X0 read-modify-written in shared
// memory; ||A0||_1 from the packed tile.  The tail (T, R, residual, certificate) reuses the four regions.
#pragma once
#include "linalg.cuh"

namespace gecon {

template <int NP, int C>
struct CwCfg {
    static_assert(NP % 8 == 0 && NP >= 8 && NP <= 32, "one lane per row: NP <= 32");
    static_assert(C >= 1 && 8 * C <= NP, "C column tiles");
    static constexpr int LD = NP + 4;
    static constexpr int LDC = 8 * C + 4;
    static constexpr int NS = NP / 8;
    static constexpr int KS = 2 * C;
    static constexpr int TW = NP * LD, TC = NP * LDC;
    static constexpr int PER_WARP_D = 2 * TW + 2 * TC;  // doubles (even)
    static constexpr int PER_WARP_I = 2 * NP + 8;       // piv[NP], flag[NP], spare
    static constexpr size_t bytes(int wpc) { return (size_t)wpc * (sizeof(double) * PER_WARP_D + sizeof(int) * PER_WARP_I) + sizeof(int) * 2 * NP; }
};

// exact max of non-negative doubles across the warp (NaN-propagating: NaN patterns order above +inf)
__device__ __forceinline__ double cw_max_nonneg(double v) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v));
    const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, (khi == mhi) ? klo : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

// global row-major rows x (columns [col_lo, col_hi)) -> zero-padded NP x LDT tile, one row per step (NP, LDT <= 36)
template <int NP, int LDT>
__device__ __forceinline__ void cw_load(double* __restrict__ dst, const double* __restrict__ src, int rows, int col_lo, int col_hi, int ldg, int lane) {
    static_assert(LDT <= 64, "two lanes' worth of columns at most");
#pragma unroll 4
    for (int r = 0; r < NP; ++r) {
#pragma unroll
        for (int h = 0; h < (LDT + 31) / 32; ++h) {
            const int c = lane + 32 * h;
            const int gc = col_lo + c;
            const double v = (r < rows && gc < col_hi) ? src[(size_t)r * ldg + gc] : 0.0;
            if (c < LDT) dst[r * LDT + c] = v;
        }
    }
}

template <int NP, int LDT>
__device__ __forceinline__ void cw_zero(double* __restrict__ dst, int lane) {
    for (int i = lane; i < NP * LDT; i += 32) dst[i] = 0.0;
}

// max absolute column sum of an NP x LDT tile whose padding is zero; every lane gets the result; NaN-propagating
template <int NP, int LDT>
__device__ __forceinline__ double cw_norm1(const double* __restrict__ M, int lane) {
    double mx = 0.0;
#pragma unroll
    for (int h = 0; h < (LDT + 31) / 32; ++h) {
        const int c = lane + 32 * h;
        double s = 0.0;
        if (c < LDT) {
#pragma unroll 8
            for (int i = 0; i < NP; ++i) s += fabs(M[i * LDT + c]);
        }
        mx = (s > mx || s != s) ? s : mx;
    }
    return cw_max_nonneg(mx);
}

// Blocked Gauss-Jordan by ONE warp: [Xa | Xb] <- M^-1 [Xa | Xb].  M is read from Ms during the first block step and
// worked on in Md afterwards (Ms == Md allowed); Xa (nta column tiles, leading dimension LDA) and Xb (ntb, LDB) are
// transformed in place.  Pivoting, panel arithmetic and the update formula are those of gj_solve_blocked (linalg.cuh):
// rows are never swapped, solution row j ends up in row s_piv[j].  Returns false on a zero / non-finite pivot.
template <int NP, int LDA, int LDB>
__device__ __forceinline__ bool cw_gj(const double* Ms, double* Md, double* Xa, int nta, double* Xb, int ntb, int n, int* __restrict__ s_piv,
                                      int* __restrict__ s_flag, int lane) {
    constexpr int LD = NP + 4, NS = NP / 8;
    constexpr unsigned IDXBITS = 5u, IDXMASK = 31u;
    const int g = lane >> 2, q = lane & 3;
    const int nblk = (n + 7) >> 3;
    bool used = (lane >= n);
    for (int kb = 0; kb < nblk; ++kb) {
        const int c0 = 8 * kb;
        const bool first = (kb == 0);
        const double* Mc = first ? Ms : Md;
        // ------------------------------------------------------------------------------------------ panel (lane = row)
        {
            const int jmax = min(8, n - c0);
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                double2 t = make_double2(0.0, 0.0);
                if (lane < NP) t = *reinterpret_cast<const double2*>(Mc + lane * LD + c0 + c);
                a[c] = t.x;
                a[c + 1] = t.y;
            }
            bool inP = false;
            unsigned fail = 0u;
            int myr = 0;
            double myinv = 1.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                if (jj < jmax) {  // warp-uniform
                    const unsigned kk = ((unsigned)__double2hiint(fabs(a[jj])) & ~IDXMASK) | (IDXMASK - (unsigned)lane);
                    const unsigned key = used ? 0u : kk;
                    const double invo = rcp_nr(a[jj]);
                    const unsigned best = __reduce_max_sync(0xffffffffu, key);
                    const int r = (int)(IDXMASK - (best & IDXMASK));
                    fail |= ((best >> IDXBITS) == 0u || best >= 0x7ff00000u) ? 1u : 0u;
                    const double inv = shfl_f64(invo, r);
                    double pr[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c != jj) pr[c] = shfl_f64(a[c], r);
                    const bool is_r = (lane == r);
                    const double m = is_r ? 0.0 : a[jj] * inv;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c != jj) a[c] = fma(-m, pr[c], a[c]);
                    a[jj] = is_r ? 1.0 : -m;
                    myinv = is_r ? inv : myinv;
                    used = used || is_r;
                    inP = inP || is_r;
                    if (lane == jj) myr = r;
                }
            }
            if (fail) return false;  // warp-uniform
            if (lane < NP) {
#pragma unroll
                for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(Md + lane * LD + c0 + c) = make_double2(a[c] * myinv, a[c + 1] * myinv);
                s_flag[lane] = inP ? 1 : 0;
            }
            if (lane < 8) s_piv[c0 + lane] = myr;
        }
        __syncwarp();
        // ------------------------------------------------------------------------------------------ update (DMMA, k = 8)
        double a0[NS], a1[NS];
        bool keep[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int r = 8 * s + g;
            a0[s] = Md[r * LD + c0 + q];
            a1[s] = Md[r * LD + c0 + 4 + q];
            keep[s] = (s_flag[r] == 0);
        }
        const int p0 = s_piv[c0 + q], p1 = s_piv[c0 + 4 + q];
        // one 8-column tile: rows <- [not a pivot row] old rows + W . old pivot rows
        auto upd = [&](const double* src, double* dst, int ld, bool inplace) {
            const double b0 = src[p0 * ld + g], b1 = src[p1 * ld + g];
            double acc[NS][2];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                double2 o = make_double2(0.0, 0.0);
                if (keep[s]) o = *reinterpret_cast<const double2*>(src + (8 * s + g) * ld + 2 * q);
                acc[s][0] = o.x;
                acc[s][1] = o.y;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) dmma884(acc[s][0], acc[s][1], a0[s], b0);
#pragma unroll
            for (int s = 0; s < NS; ++s) dmma884(acc[s][0], acc[s][1], a1[s], b1);
            if (inplace) __syncwarp();
#pragma unroll
            for (int s = 0; s < NS; ++s) *reinterpret_cast<double2*>(dst + (8 * s + g) * ld + 2 * q) = make_double2(acc[s][0], acc[s][1]);
        };
        const bool m_inplace = !(first && Ms != Md);
        for (int ct = kb + 1; ct < nblk; ++ct) upd(Mc + 8 * ct, Md + 8 * ct, LD, m_inplace);
        for (int ct = 0; ct < nta; ++ct) upd(Xa + 8 * ct, Xa + 8 * ct, LDA, true);
        for (int ct = 0; ct < ntb; ++ct) upd(Xb + 8 * ct, Xb + 8 * ct, LDB, true);
        __syncwarp();
    }
    return true;
}

// acc[s][ct] += af[s][ks] * X[rowmap(kbase + 4 ks + q)][8 ct + g]  for ks < nks, ct < nct: the product of a matrix whose
// A-operand fragments are in registers with the rows of a packed tile, read through the pivot-row map (rows >= n read row 0
// against a zero fragment).
template <int NP, int C, int LDX>
__device__ __forceinline__ void cw_prod(double (&acc)[NP / 8][C][2], const double (&af)[NP / 8][2 * C], const double* __restrict__ X,
                                        const int* __restrict__ rowmap, int kbase, int nks, int nct, int n, int lane) {
    constexpr int NS = NP / 8, KS = 2 * C;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        if (ks < nks) {
            const int k = kbase + 4 * ks + q;
            const int row = (k < n) ? (rowmap ? rowmap[k] : k) : 0;
            const double* xr = X + row * LDX + g;
#pragma unroll
            for (int ct = 0; ct < C; ++ct) {
                if (ct < nct) {
                    const double b = xr[8 * ct];
#pragma unroll
                    for (int s = 0; s < NS; ++s) dmma884(acc[s][ct][0], acc[s][ct][1], af[s][ks], b);
                }
            }
        }
    }
}

template <int NS, int C>
__device__ __forceinline__ void cw_acc_zero(double (&acc)[NS][C][2]) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int ct = 0; ct < C; ++ct) acc[s][ct][0] = acc[s][ct][1] = 0.0;
}

// acc-layout element (row 8 s + g, packed columns 8 ct + 2 q + {0,1}) of a packed tile
template <int NS, int C, int LDX>
__device__ __forceinline__ void cw_acc_store_neg(const double (&acc)[NS][C][2], double* __restrict__ X, int nct, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int ct = 0; ct < C; ++ct)
            if (ct < nct) *reinterpret_cast<double2*>(X + (8 * s + g) * LDX + 8 * ct + 2 * q) = make_double2(-acc[s][ct][0], -acc[s][ct][1]);
}

// M[:, off + packed column] -= acc  (full-width tile, leading dimension LD; off even; columns >= NP skipped: they hold zeros)
template <int NP, int C>
__device__ __forceinline__ void cw_sub_into(double* __restrict__ M, const double (&acc)[NP / 8][C][2], int off, int nct, int lane) {
    constexpr int LD = NP + 4, NS = NP / 8;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ct = 0; ct < C; ++ct) {
        if (ct < nct) {
            const int col = off + 8 * ct + 2 * q;
            if (col < NP) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    double2* ptr = reinterpret_cast<double2*>(M + (8 * s + g) * LD + col);
                    double2 v = *ptr;
                    v.x -= acc[s][ct][0];
                    v.y -= acc[s][ct][1];
                    *ptr = v;
                }
            }
        }
    }
}

// resident CTAs of WPC warps per SM that shared memory allows: the register allocator is held to that (at most 255 registers,
// at least 128: four CTAs of four warps)
template <int NP, int C, int WPC>
constexpr int cw_min_ctas() {
    const int by_smem = (int)((227 * 1024) / (CwCfg<NP, C>::bytes(WPC) + 1024));
    const int cap = 16 / WPC > 0 ? 16 / WPC : 1;  // 16 warps of 128 registers fill the register file
    return by_smem < 1 ? 1 : (by_smem > cap ? cap : by_smem);
}

template <int NP, int C, int WPC>
__global__ void __launch_bounds__(WPC * 32, cw_min_ctas<NP, C, WPC>()) cr_warp_kernel(const gecon_cr_args p, const cw_ranges rg) {
    using K = CwCfg<NP, C>;
    constexpr int LD = K::LD, LDC = K::LDC, NS = K::NS, KS = K::KS, TW = K::TW, TC = K::TC;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    double* A1 = sm + (size_t)warp * K::PER_WARP_D;
    double* W = A1 + TW;
    double* X0 = W + TW;
    double* X2 = X0 + TC;
    int* ibase = reinterpret_cast<int*>(sm + (size_t)WPC * K::PER_WARP_D);
    int* s_piv = ibase + warp * K::PER_WARP_I;
    int* s_flag = s_piv + NP;
    int* s_perm = ibase + WPC * K::PER_WARP_I;
    int* s_lead = s_perm + NP;

    const int n = p.n, k = p.k;
    const int no = (p.unperm && p.n_out > 0) ? p.n_out : n;
    const int nl = p.lead_idx ? p.n_lead : 0;
    for (int i = tid; i < NP; i += WPC * 32) {
        s_perm[i] = (i < no) ? (p.unperm ? p.unperm[i] : i) : 0;
        s_lead[i] = (i < nl) ? p.lead_idx[i] : 0;
    }
    __syncthreads();

    const int o0 = rg.o0, w0 = rg.w0, o2 = rg.o2, w2 = rg.w2;
    const int nt0 = (w0 + 7) >> 3, nt2 = (w2 + 7) >> 3;   // column tiles of the packed ranges
    const int nk0 = (w0 + 3) >> 2, nk2 = (w2 + 3) >> 2;   // k-steps
    const int kd = (p.D && p.R) ? k : 0;
    const int ntd = (kd + 7) >> 3;
    const long long stride = (long long)gridDim.x * WPC;

    for (long long draw = (long long)blockIdx.x * WPC + warp; draw < p.N; draw += stride) {
        const double* gA = p.A + (size_t)draw * n * n;
        const double* gB = p.B + (size_t)draw * n * n;
        const double* gC = p.C + (size_t)draw * n * n;
        const double* gD = p.D ? p.D + (size_t)draw * n * k : nullptr;
        {   // this warp's next draw: pull A, B, C into L2 now (they were just written by the Jacobian kernel, i.e. sit in HBM)
            const long long nxt = draw + stride;
            if (nxt < p.N) {
                const size_t bytes = (size_t)n * n * sizeof(double);
                for (size_t off = (size_t)lane * 128; off < bytes; off += 32 * 128) {
                    prefetch_l2(reinterpret_cast<const char*>(p.A + (size_t)nxt * n * n) + off);
                    prefetch_l2(reinterpret_cast<const char*>(p.B + (size_t)nxt * n * n) + off);
                    prefetch_l2(reinterpret_cast<const char*>(p.C + (size_t)nxt * n * n) + off);
                }
            }
        }
        cw_load<NP, LD>(A1, gB, n, 0, n, n, lane);
        cw_load<NP, LDC>(X0, gA, n, o0, o0 + w0, n, lane);
        cw_load<NP, LDC>(X2, gC, n, o2, o2 + w2, n, lane);
        double H[NS][C][2];  // sum of A2 X0 over the iterations: A1hat = B - H on the lag columns
        cw_acc_zero<NS, C>(H);
        __syncwarp();

        int status = 0;
        bool converged = false;
        int it = 0;
        double a0n = 0.0, a2n = 0.0;
        bool gj_failed = false;
        while (it < p.max_iter) {
            ++it;
            // A-operand fragments of A0, A2 (the solve below overwrites the tiles with X0, X2)
            double a0f[NS][KS], a2f[NS][KS];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    a0f[s][ks] = X0[(8 * s + g) * LDC + 4 * ks + q];
                    a2f[s][ks] = X2[(8 * s + g) * LDC + 4 * ks + q];
                }
            __syncwarp();
            const bool ok = cw_gj<NP, LDC, LDC>(A1, W, X0, nt0, X2, nt2, n, s_piv, s_flag, lane);
            if (!ok) {  // LAPACK: singular U -> inf / NaN in getrs -> NaN norm -> the loop stops (cycle_reduction.py:170-176)
                a0n = __longlong_as_double(0x7ff8000000000000ll);
                status |= GECON_ST_CR_NAN;
                gj_failed = true;
                break;
            }
            double m00[NS][C][2], m22[NS][C][2];
            {
                double m20[NS][C][2];
                cw_acc_zero<NS, C>(m20);
                cw_prod<NP, C, LDC>(m20, a2f, X0, s_piv, o2, nk2, nt0, n, lane);
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int ct = 0; ct < C; ++ct) {
                        H[s][ct][0] += m20[s][ct][0];
                        H[s][ct][1] += m20[s][ct][1];
                    }
                cw_sub_into<NP, C>(A1, m20, o0, nt0, lane);
            }
            __syncwarp();  // the two read-modify-write passes over A1 may touch the same elements from different lanes
            {
                double m02[NS][C][2];
                cw_acc_zero<NS, C>(m02);
                cw_prod<NP, C, LDC>(m02, a0f, X2, s_piv, o0, nk0, nt2, n, lane);
                cw_sub_into<NP, C>(A1, m02, o2, nt2, lane);
            }
            cw_acc_zero<NS, C>(m00);
            cw_acc_zero<NS, C>(m22);
            cw_prod<NP, C, LDC>(m00, a0f, X0, s_piv, o0, nk0, nt0, n, lane);
            cw_prod<NP, C, LDC>(m22, a2f, X2, s_piv, o2, nk2, nt2, n, lane);
            __syncwarp();  // every lane is done reading X0, X2
            cw_acc_store_neg<NS, C, LDC>(m00, X0, nt0, lane);
            cw_acc_store_neg<NS, C, LDC>(m22, X2, nt2, lane);
            __syncwarp();
            a0n = cw_norm1<NP, LDC>(X0, lane);
            if (a0n < p.tol) {
                a2n = cw_norm1<NP, LDC>(X2, lane);
                if (a2n < p.tol) {
                    converged = true;
                    break;
                }
            } else if (a0n != a0n) {
                status |= GECON_ST_CR_NAN;
                break;
            }
        }
        double a1n = 0.0;
        if (!converged) {
            status |= GECON_ST_CR_NOT_CONVERGED;
            if (p.norms) {  // diagnostics of the numpy twin's failure tuple (cycle_reduction.py:101-109)
                a2n = cw_norm1<NP, LDC>(X2, lane);
                a1n = cw_norm1<NP, LD>(A1, lane);
                if (gj_failed) a2n = a1n = a0n;  // a failed solve NaN-fills everything downstream (LAPACK: inf / NaN from getrs)
            }
        }
        if (p.norms && lane == 0) {
            p.norms[3 * draw] = a0n;
            p.norms[3 * draw + 1] = a2n;
            p.norms[3 * draw + 2] = a1n;
        }
        __syncwarp();

        // ---- tail.  The iterated A0, A1, A2 are dead; the four regions are reused:
        //   W   B - H (= A1hat), then B + C T       X0  A (right-hand side), then C's lead columns
        //   X2  T (natural row order)                A1  D -> R, then certificate scratch
        // T = -A1hat^-1 A (cycle_reduction.py:181); 0 if not converged.  T's non-zero columns are the lag columns: packed tile.
        double* Tt = X2;
        bool t_nan = false;
        if (converged) {
            cw_load<NP, LD>(W, gB, n, 0, n, n, lane);
            cw_load<NP, LDC>(X0, gA, n, o0, o0 + w0, n, lane);
            __syncwarp();
            cw_sub_into<NP, C>(W, H, o0, nt0, lane);
            __syncwarp();
            const bool ok = cw_gj<NP, LDC, LDC>(W, W, X0, nt0, nullptr, 0, n, s_piv, s_flag, lane);
            if (!ok) {
                status |= GECON_ST_SINGULAR;
                t_nan = true;
            }
        }
        {
            const double qnan = __longlong_as_double(0x7ff8000000000000ll);
            for (int i = lane; i < TC; i += 32) {
                const int r = i / LDC, c = i - r * LDC;
                double v = 0.0;
                if (t_nan) v = qnan;
                else if (converged && r < n) v = -X0[s_piv[r] * LDC + c];  // natural row order and the sign
                Tt[i] = v;
            }
        }
        __syncwarp();

        // ---- W = B + C T (only the lag columns differ from B); resid = sum((A + W T)^2) = sum((A + B T + C T T)^2)
        cw_load<NP, LD>(W, gB, n, 0, n, n, lane);
        cw_load<NP, LDC>(X0, gC, n, o2, o2 + w2, n, lane);
        __syncwarp();
        {
            double cf[NS][KS];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) cf[s][ks] = -X0[(8 * s + g) * LDC + 4 * ks + q];
            double ct[NS][C][2];
            cw_acc_zero<NS, C>(ct);
            cw_prod<NP, C, LDC>(ct, cf, Tt, nullptr, o2, nk2, nt0, n, lane);  // -(C T)
            cw_sub_into<NP, C>(W, ct, o0, nt0, lane);
        }
        __syncwarp();
        double resid;
        {
            // e = A + W T on the lag columns (every other column of A and of T is zero); k runs over all rows of T
            double e[NS][C][2];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ct = 0; ct < C; ++ct)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = 8 * s + g, pc = 8 * ct + 2 * q + h;
                        e[s][ct][h] = (r < n && pc < w0) ? gA[(size_t)r * n + o0 + pc] : 0.0;  // (L2-resident: read above)
                    }
            const int nkn = (n + 3) >> 2;
            for (int ks = 0; ks < nkn; ++ks) {
                const int kk = 4 * ks + q;
                double a[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) a[s] = W[(8 * s + g) * LD + kk];
                const double* xr = Tt + kk * LDC + g;
#pragma unroll
                for (int ct = 0; ct < C; ++ct) {
                    if (ct < nt0) {
                        const double b = xr[8 * ct];
#pragma unroll
                        for (int s = 0; s < NS; ++s) dmma884(e[s][ct][0], e[s][ct][1], a[s], b);
                    }
                }
            }
            double ss = 0.0;
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ct = 0; ct < C; ++ct) ss += e[s][ct][0] * e[s][ct][0] + e[s][ct][1] * e[s][ct][1];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            resid = ss;
        }
        if (p.resid_tol > 0.0 && !(resid < p.resid_tol)) status |= GECON_ST_RESID;

        // ---- T out (sub-block gather through unperm); columns outside the lag range are zero (NaN when the solve failed)
        {
            double* gT = p.T + (size_t)draw * (p.t_stride ? (size_t)p.t_stride : (size_t)no * no);
            const int ldt = p.t_ld ? p.t_ld : no;
            const double fill = t_nan ? __longlong_as_double(0x7ff8000000000000ll) : 0.0;
            for (int i = lane; i < no * no; i += 32) {
                const int r = i / no, c = i - r * no;
                const int pc = s_perm[c] - o0;
                gT[(size_t)r * ldt + c] = (pc >= 0 && pc < w0) ? Tt[s_perm[r] * LDC + pc] : fill;
            }
        }

        // ---- R = -W^-1 D (shared.py:74-75) and, for the Blanchard-Kahn certificate, W^-1 C on the lead columns (= -F):
        // one solve, right-hand sides D (packed tile in the A1 region) and C's lead columns (already in X0)
        const bool want_cert = p.lead_idx && converged && !t_nan;
        bool have_F = false;
        double* Dt = A1;
        if (kd || want_cert) {
            if (kd) cw_load<NP, LDC>(Dt, gD, n, 0, kd, k, lane);
            __syncwarp();
            const bool ok = cw_gj<NP, LDC, LDC>(W, W, Dt, ntd, X0, want_cert ? nt2 : 0, n, s_piv, s_flag, lane);
            if (!ok) status |= GECON_ST_SINGULAR;
            have_F = ok && want_cert;
            if (kd) {
                double* gR = p.R + (size_t)draw * (p.r_stride ? (size_t)p.r_stride : (size_t)no * k);
                const double qnan = __longlong_as_double(0x7ff8000000000000ll);
                for (int i = lane; i < no * k; i += 32) {
                    const int r = i / k, c = i - r * k;
                    gR[(size_t)r * k + c] = ok ? -Dt[s_piv[s_perm[r]] * LDC + c] : qnan;
                }
            }
        }
        if (lane == 0) {
            if (p.resid) p.resid[draw] = resid;
            if (p.n_iter) p.n_iter[draw] = it;
        }
        __syncwarp();

        // ---- Blanchard-Kahn certificate: rho(T_LL) < 1 and rho(F_FF) < 1 by repeated squaring (see gecon_cr_args.lead_idx).
        // Z1 = T[lag][:, lag] -> W region, Z2 = (W^-1 C)[lead][:, lead] -> A1 region; their sources (X2, X0) become the
        // ping-pong buffers once copied.
        bool certified = false;
        if (p.lead_idx) {
            if (have_F) {
                const int s1 = min(w0, n - o0);  // the lag block sits at rows o0.. / packed columns 0.. of T
                double* Z1 = W;
                double* Z2 = A1;
                double* Z1b = X2;
                double* Z2b = X0;
                for (int i = lane; i < TC; i += 32) {
                    const int r = i / LDC, c = i - r * LDC;
                    Z1[i] = (r < s1 && c < s1) ? Tt[(o0 + r) * LDC + c] : 0.0;
                    double z = 0.0;
                    if (r < nl && c < nl) {
                        const int pc = s_lead[c] - o2;
                        z = (pc >= 0 && pc < w2) ? X0[s_piv[s_lead[r]] * LDC + pc] : 0.0;
                    }
                    Z2[i] = z;
                }
                __syncwarp();
                cw_zero<NP, LDC>(Z1b, lane);
                cw_zero<NP, LDC>(Z2b, lane);
                __syncwarp();
                const int k1 = (s1 + 3) >> 2, c1 = (s1 + 7) >> 3, k2 = (nl + 3) >> 2, c2 = (nl + 7) >> 3;
                bool ok1 = (s1 == 0), ok2 = (nl == 0);
                // Zb = Z Z on the leading nc x nc tiles
                auto square = [&](const double* Z, double* Zb, int nk, int nc) {
                    double acc[NS][C][2];
                    cw_acc_zero<NS, C>(acc);
                    for (int ks = 0; ks < nk; ++ks) {
                        const int kk = 4 * ks + q;
                        double a[NS];
#pragma unroll
                        for (int s = 0; s < NS; ++s) a[s] = (s < nc) ? Z[(8 * s + g) * LDC + kk] : 0.0;
#pragma unroll
                        for (int ct = 0; ct < C; ++ct) {
                            if (ct < nc) {
                                const double b = Z[kk * LDC + 8 * ct + g];
#pragma unroll
                                for (int s = 0; s < NS; ++s)
                                    if (s < nc) dmma884(acc[s][ct][0], acc[s][ct][1], a[s], b);
                            }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < NS; ++s)
#pragma unroll
                        for (int ct = 0; ct < C; ++ct)
                            if (s < nc && ct < nc) *reinterpret_cast<double2*>(Zb + (8 * s + g) * LDC + 8 * ct + 2 * q) = make_double2(acc[s][ct][0], acc[s][ct][1]);
                };
                for (int sq = 0; sq <= 14; ++sq) {
                    // the norms are only inspected after every second squaring (and at the start)
                    const bool look = (sq & 1) == 0 || sq == 14;
                    const double n1 = (ok1 || !look) ? (ok1 ? 0.0 : 2.0) : cw_norm1<NP, LDC>(Z1, lane);
                    const double n2 = (ok2 || !look) ? (ok2 ? 0.0 : 2.0) : cw_norm1<NP, LDC>(Z2, lane);
                    ok1 = ok1 || (n1 < 1.0);
                    ok2 = ok2 || (n2 < 1.0);
                    if (ok1 && ok2) {
                        certified = true;
                        break;
                    }
                    if (!(n1 < 1e100) || !(n2 < 1e100) || sq == 14) break;  // growing powers / NaN: leave it to bk_count
                    if (!ok1) {
                        square(Z1, Z1b, k1, c1);
                        double* t = Z1;
                        Z1 = Z1b;
                        Z1b = t;
                    }
                    if (!ok2) {
                        square(Z2, Z2b, k2, c2);
                        double* t = Z2;
                        Z2 = Z2b;
                        Z2b = t;
                    }
                    __syncwarp();
                }
            }
            if (certified) status |= GECON_ST_BK_CERTIFIED;
        }
        if (lane == 0) {
            p.status[draw] = p.accumulate ? (p.status[draw] | status) : status;
            if (p.lead_idx && p.n_unstable) p.n_unstable[draw] = certified ? nl : -1;
        }
        __syncwarp();
    }
}

}  // namespace gecon
