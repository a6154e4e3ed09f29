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
// number), C column tiles each, next to the Gauss-Jordan workspace in ONE augmented tile:
//   shared memory per warp   A1 (NP x LD)  and  WA = [W | X0 | X2]  (NP x LDW, LDW = NP + 16 C + 4)
//   registers                the A-operand fragments of A0 and A2 (loaded before the in-place solve overwrites them with
//                            X0, X2), the accumulated correction H = sum A2 X0 (A1hat = B - H)
// Per iteration: W <- A1; [X0 | X2] = A1^-1 [A0 | A2] by blocked Gauss-Jordan on the augmented tile (8-column panels, lane =
// row, pivots by redux.sync, every trailing column tile updated by a DMMA product with k = 8); four DMMA products from the
// register fragments through the pivot-row map; A1 -= A0 X2 + A2 X0 read-modify-written in shared memory; ||A0||_1 from the
// packed block.  The tail (T, R, residual, certificate) reuses the same two regions.
//
// Code size is a first-class constraint here (first version, fully unrolled and inlined three times: 9.6 k SASS instructions,
// ncu stall "no instruction" 5.3 warps per issue -- twelve warps at twelve different places of 150 KB of code against a
// 32 KB L1.5 instruction cache): the Gauss-Jordan solve and the power bound are single __noinline__ functions, the pivot
// loop of a panel is rolled (the panel registers rotate so that the current column is always a[0]), and the trailing update
// is one loop over the column tiles of the augmented matrix.
#pragma once
#include "linalg.cuh"

// Per-model build (cr_warp_spec.cu, compiled into every generated model library): the number of variables and the packed lag / lead
// column ranges are compile-time constants
//     GECON_CW_SPEC_N, GECON_CW_SPEC_O0, GECON_CW_SPEC_W0, GECON_CW_SPEC_O2, GECON_CW_SPEC_W2
// so every range guard of the products, the panel's column counts and the norm loops fold away (measured on the medium NK model: 33.4 ->
// 28.3 ms per 262,144 draws, 0.43 -> 0.51 of the fp64 peak).  The kernel gets its own name there: it must never be confused with the
// generic instantiation of the core library.
#ifdef GECON_CW_SPEC_N
#define cr_warp_kernel cr_warp_spec_kernel
#endif

namespace gecon {

template <int NP, int C>
struct CwCfg {
    static_assert(NP % 8 == 0 && NP >= 8 && NP <= 32, "one lane per row: NP <= 32");
    static_assert(C >= 1 && 8 * C <= NP, "C column tiles");
    static constexpr int LD = NP + 4;             // A1
    static constexpr int LDW = NP + 16 * C + 4;   // augmented [W | Xa | Xb]
    static constexpr int LDC = 8 * C + 4;         // packed tiles of the tail (T, certificate)
    static constexpr int NS = NP / 8;             // row strips = column tiles of W
    static constexpr int KS = 2 * C;              // k-steps spanned by a packed range
    static constexpr int XA = NP, XB = NP + 8 * C;  // column offsets of the two right-hand-side blocks
    static constexpr int TA1 = NP * LD, TWA = NP * LDW;
    static constexpr int PER_WARP_D = TA1 + TWA;  // doubles (even)
    static constexpr int PER_WARP_I = 2 * NP + 8; // piv[NP], flag[NP], spare
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

// global row-major `rows` x (columns [col_lo, col_hi)) -> `width` columns of a tile with leading dimension ldt, zero padded to
// NP rows; one row per step (width <= 64)
template <int NP>
__device__ __noinline__ void cw_load(double* __restrict__ dst, int ldt, int width, const double* __restrict__ src, int rows, int col_lo, int col_hi, int ldg,
                                     int lane) {
    for (int c = lane; c < width; c += 32) {
        const int gc = col_lo + c;
        const bool colok = gc < col_hi;
#pragma unroll 1
        for (int r0 = 0; r0 < NP; r0 += 8) {  // eight loads in flight per lane
            double v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (colok && r0 + e < rows) ? src[(size_t)(r0 + e) * ldg + gc] : 0.0;
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[(r0 + e) * ldt + c] = v[e];
        }
    }
}

// The same from a compact Jacobian (gecon_compact_jac): matrix q (0 = A, 1 = B, 2 = C, 3 = D) of draw `draw`, columns
// [col_lo, col_hi) -> `width` columns of a zeroed NP-row tile.  The structural non-zeros are ~5 % of the dense entries, so
// this is a handful of scattered stores instead of NP row loads.
template <int NP>
__device__ __noinline__ void cw_load_compact(double* __restrict__ dst, int ldt, int width, const gecon_compact_jac& cj, long long draw, int q, int col_lo,
                                             int col_hi, int lane) {
    for (int c = lane; c < width; c += 32) {
#pragma unroll 4
        for (int r = 0; r < NP; ++r) dst[r * ldt + c] = 0.0;
    }
    __syncwarp();
    const double* v = cj.vals + (size_t)draw * cj.stride;
    for (int e = cj.off[q] + lane; e < cj.off[q + 1]; e += 32) {
        const int rc = cj.table[e];
        const int c = (rc & 0xffff) - col_lo;
        if (c >= 0 && (rc & 0xffff) < col_hi && c < width) dst[(rc >> 16) * ldt + c] = v[e];
    }
}

// max absolute column sum over `rows` rows x `width` columns (leading dimension ldt); every lane gets it; NaN-propagating
static __device__ __noinline__ double cw_norm1(const double* __restrict__ M, int ldt, int rows, int width, int lane) {
    double mx = 0.0;
    for (int c = lane; c < width; c += 32) {
        double s0 = 0.0, s1 = 0.0;
        int i = 0;
        for (; i + 1 < rows; i += 2) {
            s0 += fabs(M[i * ldt + c]);
            s1 += fabs(M[(i + 1) * ldt + c]);
        }
        if (i < rows) s0 += fabs(M[i * ldt + c]);
        const double s = s0 + s1;
        mx = (s > mx || s != s) ? s : mx;
    }
    return cw_max_nonneg(mx);
}

// Blocked Gauss-Jordan by ONE warp on the augmented tile WA = [M | Xa | Xb] (NP rows, leading dimension LDW): the
// right-hand-side blocks (nta / ntb column tiles at column offsets XA / XB) are replaced by M^-1 [Xa | Xb]; M is destroyed.
// Pivoting, panel arithmetic and the update formula are those of gj_solve_blocked (linalg.cuh): partial pivoting (largest
// |.| among the rows not used yet, lowest row on ties), rows are never swapped, solution row j ends up in row s_piv[j].
// Returns false on a zero / non-finite pivot (the blocks then hold garbage).
template <int NP, int C>
__device__ __noinline__ bool cw_gj(double* __restrict__ WA, int nta, int ntb, int n, int* __restrict__ s_piv, int* __restrict__ s_flag, int lane) {
    using K = CwCfg<NP, C>;
    constexpr int LDW = K::LDW, NS = K::NS;
    constexpr unsigned IDXBITS = 5u, IDXMASK = 31u;
    const int g = lane >> 2, q = lane & 3;
#ifdef GECON_CW_SPEC_N
    n = GECON_CW_SPEC_N;
#endif
    const int nblk = (n + 7) >> 3;
    bool used = (lane >= n);
    unsigned fail = 0u;
#pragma unroll 1
    for (int kb = 0; kb < nblk; ++kb) {
        const int c0 = 8 * kb;
        // ------------------------------------------------------------------------------------------ panel (lane = row)
        // The row's 8 panel entries live in registers.  Per pivot step every row publishes [1 / a_u, other entries] in its own
        // panel columns of WA (they are dead until the write-back below), the pivot row is found with one redux.sync and read
        // back by every lane with four broadcast 16-byte loads: 10 shared-memory instructions instead of 18 shuffles and
        // the 16 moves that reassemble doubles from them.  Steps are unrolled by four; the two register halves swap in between.
        {
            const int jmax = min(8, n - c0);
            double* prow = WA + (lane < NP ? lane : 0) * LDW + c0;
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                double2 t = make_double2(0.0, 0.0);
                if (lane < NP) t = *reinterpret_cast<const double2*>(prow + c);
                a[c] = t.x;
                a[c + 1] = t.y;
            }
            bool inP = false;
            double myinv = 1.0;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int jj = 4 * h + u;
                    if (jj < jmax) {  // warp-uniform
                        const unsigned kk = ((unsigned)__double2hiint(fabs(a[u])) & ~IDXMASK) | (IDXMASK - (unsigned)lane);
                        const unsigned key = used ? 0u : kk;
                        const double invo = rcp_nr2(a[u]);  // every row's own reciprocal: no division after the pivot search
                        if (lane < NP) {
#pragma unroll
                            for (int c = 0; c < 8; c += 2)
                                *reinterpret_cast<double2*>(prow + c) = make_double2(c == u ? invo : a[c], c + 1 == u ? invo : a[c + 1]);
                        }
                        const unsigned best = __reduce_max_sync(0xffffffffu, key);
                        const int r = (int)(IDXMASK - (best & IDXMASK));
                        fail |= ((best >> IDXBITS) == 0u || best >= 0x7ff00000u) ? 1u : 0u;  // zero / subnormal / inf / NaN pivot
                        __syncwarp();
                        double pv[8];
#pragma unroll
                        for (int c = 0; c < 8; c += 2) {
                            const double2 t = *reinterpret_cast<const double2*>(WA + r * LDW + c0 + c);
                            pv[c] = t.x;
                            pv[c + 1] = t.y;
                        }
                        __syncwarp();  // the next step's publication overwrites these rows
                        const bool is_r = (lane == r);
                        const double m = is_r ? 0.0 : a[u] * pv[u];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c != u) a[c] = fma(-m, pv[c], a[c]);
                        a[u] = is_r ? 1.0 : -m;
                        myinv = is_r ? pv[u] : myinv;
                        used = used || is_r;
                        inP = inP || is_r;
                        if (lane == 0) s_piv[c0 + jj] = r;
                    } else if (lane == 0) {
                        s_piv[c0 + jj] = 0;  // padding column: any valid row (its coefficient is zero)
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double t = a[c];
                    a[c] = a[c + 4];
                    a[c + 4] = t;
                }
            }
            // pivot rows are scaled by 1 / pivot once, here (commutes with the later eliminations acting on them)
            if (lane < NP) {
#pragma unroll
                for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(prow + c) = make_double2(a[c] * myinv, a[c + 1] * myinv);
                s_flag[lane] = inP ? 1 : 0;
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------------------------------ update (DMMA, k = 8)
        // every live column tile: rows <- [not a pivot row of this panel] old rows + W_panel . old pivot rows
        double a0[NS], a1[NS];
        bool keep[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int r = 8 * s + g;
            a0[s] = WA[r * LDW + c0 + q];
            a1[s] = WA[r * LDW + c0 + 4 + q];
            keep[s] = (s_flag[r] == 0);
        }
        const int p0 = s_piv[c0 + q] * LDW + g, p1 = s_piv[c0 + 4 + q] * LDW + g;
        const int ro = g * LDW + 2 * q;
#pragma unroll 1
        for (int ct = kb + 1; ct < NS + 2 * C; ++ct) {
            if (ct < NS ? (ct >= nblk) : (ct < NS + C ? (ct - NS >= nta) : (ct - NS - C >= ntb))) continue;  // warp-uniform
            double* tile = WA + 8 * ct;
            const double b0 = tile[p0], b1 = tile[p1];
            double acc[NS][2];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                double2 o = make_double2(0.0, 0.0);
                if (keep[s]) o = *reinterpret_cast<const double2*>(tile + 8 * s * LDW + ro);
                acc[s][0] = o.x;
                acc[s][1] = o.y;
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) dmma884(acc[s][0], acc[s][1], a0[s], b0);
#pragma unroll
            for (int s = 0; s < NS; ++s) dmma884(acc[s][0], acc[s][1], a1[s], b1);
            __syncwarp();  // in place: every lane has read the pivot rows before anybody overwrites them
#pragma unroll
            for (int s = 0; s < NS; ++s) *reinterpret_cast<double2*>(tile + 8 * s * LDW + ro) = make_double2(acc[s][0], acc[s][1]);
        }
        __syncwarp();
    }
    return fail == 0u;
}

// acc[s][ct] += af[s][ks] * X[rowmap(kbase + 4 ks + q)][8 ct + g]  for ks < nks, ct < nct: the product of a matrix whose
// A-operand fragments are in registers with the rows of a packed block (leading dimension ldx), read through the pivot-row
// map (rows >= n read row 0 against a zero fragment).
template <int NP, int C>
__device__ __forceinline__ void cw_prod(double (&acc)[NP / 8][C][2], const double (&af)[NP / 8][2 * C], const double* __restrict__ X, int ldx,
                                        const int* __restrict__ rowmap, int kbase, int nks, int nct, int n, int lane) {
    constexpr int NS = NP / 8, KS = 2 * C;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        if (ks < nks) {
            const int k = kbase + 4 * ks + q;
            const int row = (k < n) ? (rowmap ? rowmap[k] : k) : 0;
            const double* xr = X + row * ldx + g;
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

// packed block <- -acc   (accumulator layout: row 8 s + g, packed columns 8 ct + 2 q + {0,1})
template <int NS, int C>
__device__ __forceinline__ void cw_acc_store_neg(const double (&acc)[NS][C][2], double* __restrict__ X, int ldx, int nct, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int ct = 0; ct < C; ++ct)
            if (ct < nct) *reinterpret_cast<double2*>(X + (8 * s + g) * ldx + 8 * ct + 2 * q) = make_double2(-acc[s][ct][0], -acc[s][ct][1]);
}

// M[:, off + packed column] -= acc  (full-width matrix with leading dimension ldm; off even; columns >= NP skipped: zeros)
template <int NP, int C>
__device__ __forceinline__ void cw_sub_into(double* __restrict__ M, int ldm, const double (&acc)[NP / 8][C][2], int off, int nct, int lane) {
    constexpr int NS = NP / 8;
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ct = 0; ct < C; ++ct) {
        if (ct < nct) {
            const int col = off + 8 * ct + 2 * q;
            if (col < NP) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    double2* ptr = reinterpret_cast<double2*>(M + (8 * s + g) * ldm + col);
                    double2 v = *ptr;
                    v.x -= acc[s][ct][0];
                    v.y -= acc[s][ct][1];
                    *ptr = v;
                }
            }
        }
    }
}

// Is rho(Z) < 1 provable by a power of Z having 1-norm < 1?  Z: s x s (s <= 8 C) in a tile of leading dimension LDC whose
// padding (up to 8 ceil(s/8) rows and columns) is zero; Zb: scratch tile of the same shape (zeroed here).  Repeated squaring
// on the tensor path, norms inspected at the start and after every second squaring; growing powers / NaN give up.
template <int C>
__device__ __noinline__ bool cw_power_bound(double* Z, double* Zb, int s, int lane) {
    constexpr int LDC = 8 * C + 4;
    if (s == 0) return true;
    const int g = lane >> 2, q = lane & 3;
    const int nk = (s + 3) >> 2, nc = (s + 7) >> 3;
    __syncwarp();  // Zb may be the tile the previous call was still reading norms from (racecheck, round 2: the reduction that ends
                   // cw_norm1 synchronises the lanes' execution, not their shared-memory accesses)
    for (int i = lane; i < 8 * nc * LDC; i += 32) Zb[i] = 0.0;
    __syncwarp();
#pragma unroll 1
    for (int sq = 0; sq <= 14; ++sq) {
        if ((sq & 1) == 0 || sq == 14) {
            const double nrm = cw_norm1(Z, LDC, s, s, lane);
            if (nrm < 1.0) return true;
            if (!(nrm < 1e100) || sq == 14) return false;
        }
#pragma unroll 1
        for (int st = 0; st < nc; ++st) {  // row strip st of Zb = Z Z
            double acc[C][2];
#pragma unroll
            for (int ct = 0; ct < C; ++ct) acc[ct][0] = acc[ct][1] = 0.0;
            for (int ks = 0; ks < nk; ++ks) {
                const int kk = 4 * ks + q;
                const double a = Z[(8 * st + g) * LDC + kk];
#pragma unroll
                for (int ct = 0; ct < C; ++ct)
                    if (ct < nc) dmma884(acc[ct][0], acc[ct][1], a, Z[kk * LDC + 8 * ct + g]);
            }
#pragma unroll
            for (int ct = 0; ct < C; ++ct)
                if (ct < nc) *reinterpret_cast<double2*>(Zb + (8 * st + g) * LDC + 8 * ct + 2 * q) = make_double2(acc[ct][0], acc[ct][1]);
        }
        __syncwarp();
        double* t = Z;
        Z = Zb;
        Zb = t;
    }
    return false;
}

// resident CTAs of WPC warps per SM that shared memory allows: the register allocator is held to that (at most 255 registers,
// at least 128: four CTAs of four warps)
template <int NP, int C, int WPC>
constexpr int cw_min_ctas() {
    const int by_smem = (int)((227 * 1024) / (CwCfg<NP, C>::bytes(WPC) + 1024));
#ifndef GECON_CW_WARPS16
#define GECON_CW_WARPS16 16  // resident warps per SM the register allocator is held to at NP <= 16.  Measured on the RBC workload (n = 9,
                             // 65,536 draws): 16 warps (128 registers) 2.47 ms, 20 (96, spills) 2.60 ms, 24 (80, spills) 2.49 ms: not warp-count bound
#endif
    const int warps = NP <= 16 ? GECON_CW_WARPS16 : 16;  // 16 warps of 128 registers fill the register file
    const int cap = warps / WPC > 0 ? warps / WPC : 1;
    return by_smem < 1 ? 1 : (by_smem > cap ? cap : by_smem);
}

template <int NP, int C, int WPC>
__global__ void __launch_bounds__(WPC * 32, cw_min_ctas<NP, C, WPC>()) cr_warp_kernel(const gecon_cr_args p, const cw_ranges rg,
                                                                                      const gecon_compact_jac cj) {
    using K = CwCfg<NP, C>;
    constexpr int LD = K::LD, LDW = K::LDW, LDC = K::LDC, NS = K::NS, KS = K::KS, XA = K::XA, XB = K::XB;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;
    double* A1 = sm + (size_t)warp * K::PER_WARP_D;
    double* WA = A1 + K::TA1;
    double* X0 = WA + XA;  // packed lag-column block  (leading dimension LDW)
    double* X2 = WA + XB;  // packed lead-column block
    int* ibase = reinterpret_cast<int*>(sm + (size_t)WPC * K::PER_WARP_D);
    int* s_piv = ibase + warp * K::PER_WARP_I;
    int* s_flag = s_piv + NP;
    int* s_perm = ibase + WPC * K::PER_WARP_I;
    int* s_lead = s_perm + NP;

#ifdef GECON_CW_SPEC_N
    constexpr int n = GECON_CW_SPEC_N;
    const int k = p.k;
#else
    const int n = p.n, k = p.k;
#endif
    const int no = (p.unperm && p.n_out > 0) ? p.n_out : n;
    const int nl = p.lead_idx ? p.n_lead : 0;
    for (int i = tid; i < NP; i += WPC * 32) {
        s_perm[i] = (i < no) ? (p.unperm ? p.unperm[i] : i) : 0;
        s_lead[i] = (i < nl) ? p.lead_idx[i] : 0;
    }
    __syncthreads();

#ifdef GECON_CW_SPEC_N
    constexpr int o0 = GECON_CW_SPEC_O0, w0 = GECON_CW_SPEC_W0, o2 = GECON_CW_SPEC_O2, w2 = GECON_CW_SPEC_W2;
    (void)rg;
#else
    const int o0 = rg.o0, w0 = rg.w0, o2 = rg.o2, w2 = rg.w2;
#endif
    const int nt0 = (w0 + 7) >> 3, nt2 = (w2 + 7) >> 3;   // column tiles of the packed ranges
    const int nk0 = (w0 + 3) >> 2, nk2 = (w2 + 3) >> 2;   // k-steps
    const int kd = ((p.D || cj.vals) && p.R) ? k : 0;
    const int ntd = (kd + 7) >> 3;
    const long long stride = (long long)gridDim.x * WPC;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    for (long long draw = (long long)blockIdx.x * WPC + warp; draw < p.N; draw += stride) {
        const bool cmp = cj.vals != nullptr;
        const double* gA = cmp ? nullptr : p.A + (size_t)draw * n * n;
        const double* gB = cmp ? nullptr : p.B + (size_t)draw * n * n;
        const double* gC = cmp ? nullptr : p.C + (size_t)draw * n * n;
        const double* gD = (cmp || !p.D) ? nullptr : p.D + (size_t)draw * n * k;
        // matrix q (A, B, C, D) of this draw, columns [lo, hi) -> tile: dense rows or compact scatter
        auto load = [&](double* dst, int ldt, int width, int q, const double* src, int lo, int hi, int ldg) {
            if (cmp) cw_load_compact<NP>(dst, ldt, width, cj, draw, q, lo, hi, lane);
            else cw_load<NP>(dst, ldt, width, src, n, lo, hi, ldg, lane);
        };
        if (cmp) {  // this warp's next draw: its compact vector (a few cache lines) into L2
            const long long nxt = draw + stride;
            if (nxt < p.N && lane * 16 < cj.off[4]) prefetch_l2(cj.vals + (size_t)nxt * cj.stride + lane * 16);
        } else {   // this warp's next draw: pull A, B, C into L2 now (they were just written by the Jacobian kernel, i.e. sit in HBM)
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
        load(A1, LD, LD, 1, gB, 0, n, n);
        load(X0, LDW, 8 * C, 0, gA, o0, o0 + w0, n);
        load(X2, LDW, 8 * C + 4, 2, gC, o2, o2 + w2, n);  // (+ 4: the tile's padding columns)
        double H[NS][C][2];  // sum of A2 X0 over the iterations: A1hat = B - H on the lag columns
        cw_acc_zero<NS, C>(H);
        __syncwarp();

        int status = 0;
        bool converged = false, gj_failed = false;
        int it = 0;
        double a0n = 0.0, a2n = 0.0;
#pragma unroll 1
        while (it < p.max_iter) {
            ++it;
            // A-operand fragments of A0, A2 (the solve below overwrites the blocks with X0, X2), and W <- A1
            double a0f[NS][KS], a2f[NS][KS];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    a0f[s][ks] = X0[(8 * s + g) * LDW + 4 * ks + q];
                    a2f[s][ks] = X2[(8 * s + g) * LDW + 4 * ks + q];
                }
#pragma unroll 2
            for (int i = lane; i < NP * (NP / 2); i += 32) {
                const int r = i / (NP / 2), c2 = i - r * (NP / 2);
                *reinterpret_cast<double2*>(WA + r * LDW + 2 * c2) = *reinterpret_cast<const double2*>(A1 + r * LD + 2 * c2);
            }
            __syncwarp();
            if (!cw_gj<NP, C>(WA, nt0, nt2, n, s_piv, s_flag, lane)) {
                // LAPACK: singular U -> inf / NaN in getrs -> NaN norm -> the loop stops (cycle_reduction.py:170-176)
                a0n = qnan;
                status |= GECON_ST_CR_NAN;
                gj_failed = true;
                break;
            }
            double m00[NS][C][2], m22[NS][C][2];
            {
                double m20[NS][C][2];
                cw_acc_zero<NS, C>(m20);
                cw_prod<NP, C>(m20, a2f, X0, LDW, s_piv, o2, nk2, nt0, n, lane);
#pragma unroll
                for (int s = 0; s < NS; ++s)
#pragma unroll
                    for (int ct = 0; ct < C; ++ct) {
                        H[s][ct][0] += m20[s][ct][0];
                        H[s][ct][1] += m20[s][ct][1];
                    }
                cw_sub_into<NP, C>(A1, LD, m20, o0, nt0, lane);
            }
            __syncwarp();  // the two read-modify-write passes over A1 may touch the same elements from different lanes
            {
                double m02[NS][C][2];
                cw_acc_zero<NS, C>(m02);
                cw_prod<NP, C>(m02, a0f, X2, LDW, s_piv, o0, nk0, nt2, n, lane);
                cw_sub_into<NP, C>(A1, LD, m02, o2, nt2, lane);
            }
            cw_acc_zero<NS, C>(m00);
            cw_acc_zero<NS, C>(m22);
            cw_prod<NP, C>(m00, a0f, X0, LDW, s_piv, o0, nk0, nt0, n, lane);
            cw_prod<NP, C>(m22, a2f, X2, LDW, s_piv, o2, nk2, nt2, n, lane);
            __syncwarp();  // every lane is done reading X0, X2
            cw_acc_store_neg<NS, C>(m00, X0, LDW, nt0, lane);
            cw_acc_store_neg<NS, C>(m22, X2, LDW, nt2, lane);
            __syncwarp();
            a0n = cw_norm1(X0, LDW, NP, 8 * C, lane);
            if (a0n < p.tol) {
                if (p.scan_semantics) {  // the scan twin tests ||A0||_1 only (cycle_reduction.py:268-273)
                    converged = true;
                    break;
                }
                a2n = cw_norm1(X2, LDW, NP, 8 * C, lane);
                if (a2n < p.tol) {
                    converged = true;
                    break;
                }
            } else if (a0n != a0n) {
                status |= GECON_ST_CR_NAN;
                break;
            }
        }
        if (p.scan_semantics && (status & GECON_ST_CR_NAN)) it = p.max_iter;  // the scan keeps stepping on NaNs to the end
        double a1n = 0.0;
        if (!converged) {
            status |= GECON_ST_CR_NOT_CONVERGED;
            if (p.norms) {  // diagnostics of the numpy twin's failure tuple (cycle_reduction.py:101-109)
                a2n = cw_norm1(X2, LDW, NP, 8 * C, lane);
                a1n = cw_norm1(A1, LD, NP, NP, lane);
                if (gj_failed) a2n = a1n = a0n;  // a failed solve NaN-fills everything downstream (LAPACK: inf / NaN from getrs)
            }
        }
        if (p.norms && lane == 0) {
            p.norms[3 * draw] = a0n;
            p.norms[3 * draw + 1] = a2n;
            p.norms[3 * draw + 2] = a1n;
        }
        __syncwarp();

        // ---- tail.  The iterated A0, A1, A2 are dead; the two regions are reused:
        //   WA   [B - H | A] -> solve -> T;  then [B + C T | D | C] -> solve -> R, W^-1 C        A1 region   T (packed, LDC)
        // T = -A1hat^-1 A (cycle_reduction.py:181); 0 if not converged.  T's non-zero columns are the lag columns.
        double* Tt = A1;
        bool t_nan = (p.scan_semantics && gj_failed);                         // (... from NaN-filled matrices after a failed solve)
        const bool solve_t = converged || (p.scan_semantics && !gj_failed);  // the scan twin always solves for T
        if (solve_t) {
            load(WA, LDW, NP, 1, gB, 0, n, n);
            load(X0, LDW, 8 * C, 0, gA, o0, o0 + w0, n);
            __syncwarp();
            cw_sub_into<NP, C>(WA, LDW, H, o0, nt0, lane);
            __syncwarp();
            if (!cw_gj<NP, C>(WA, nt0, 0, n, s_piv, s_flag, lane)) {
                status |= GECON_ST_SINGULAR;
                t_nan = true;
            }
        }
#pragma unroll 1
        for (int i = lane; i < NP * LDC; i += 32) {
            const int r = i / LDC, c = i - r * LDC;
            double v = 0.0;
            if (t_nan) v = qnan;
            else if (solve_t && r < n && c < 8 * C) v = -X0[s_piv[r] * LDW + c];  // natural row order and the sign
            Tt[i] = v;
        }
        __syncwarp();

        // ---- W = B + C T (only the lag columns differ from B); resid = sum((A + W T)^2) = sum((A + B T + C T T)^2)
        load(WA, LDW, NP, 1, gB, 0, n, n);
        load(X2, LDW, 8 * C, 2, gC, o2, o2 + w2, n);
        load(X0, LDW, 8 * C, 0, gA, o0, o0 + w0, n);  // A again, for the residual (the Xa block is free: T lives in the A1 region)
        __syncwarp();
        {
            double cf[NS][KS];
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) cf[s][ks] = -X2[(8 * s + g) * LDW + 4 * ks + q];
            double ct[NS][C][2];
            cw_acc_zero<NS, C>(ct);
            cw_prod<NP, C>(ct, cf, Tt, LDC, nullptr, o2, nk2, nt0, n, lane);  // -(C T)
            cw_sub_into<NP, C>(WA, LDW, ct, o0, nt0, lane);
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
                    for (int h = 0; h < 2; ++h) e[s][ct][h] = X0[(8 * s + g) * LDW + 8 * ct + 2 * q + h];
            const int nkn = (n + 3) >> 2;
#pragma unroll 1
            for (int ks = 0; ks < nkn; ++ks) {
                const int kk = 4 * ks + q;
                double a[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) a[s] = WA[(8 * s + g) * LDW + kk];
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
            const double fill = t_nan ? qnan : 0.0;
#pragma unroll 1
            for (int r = 0; r < no; ++r) {
                const double* trow = Tt + s_perm[r] * LDC;
                for (int c = lane; c < no; c += 32) {
                    const int pc = s_perm[c] - o0;
                    gT[(size_t)r * ldt + c] = (pc >= 0 && pc < w0) ? trow[pc] : fill;
                }
            }
        }

        // ---- R = -W^-1 D (shared.py:74-75) and, for the Blanchard-Kahn certificate, W^-1 C on the lead columns (= -F):
        // one solve, right-hand sides D (Xa block) and C's lead columns (already in the Xb block)
        const bool want_cert = p.lead_idx && converged && !t_nan;
        bool have_F = false;
        if (kd || want_cert) {
            if (kd) load(X0, LDW, 8 * C, 3, gD, 0, kd, k);
            __syncwarp();
            const bool ok = cw_gj<NP, C>(WA, ntd, want_cert ? nt2 : 0, n, s_piv, s_flag, lane);
            if (!ok) status |= GECON_ST_SINGULAR;
            have_F = ok && want_cert;
            if (kd) {
                double* gR = p.R + (size_t)draw * (p.r_stride ? (size_t)p.r_stride : (size_t)no * k);
#pragma unroll 1
                for (int r = lane; r < no; r += 32) {  // one row per lane (k is small)
                    const double* xrow = X0 + s_piv[s_perm[r]] * LDW;
#pragma unroll 1
                    for (int c = 0; c < k; ++c) gR[(size_t)r * k + c] = ok ? -xrow[c] : qnan;
                }
            }
        }
        if (lane == 0) {
            if (p.resid) p.resid[draw] = resid;
            if (p.n_iter) p.n_iter[draw] = it;
        }

        // ---- Blanchard-Kahn certificate: rho(T_LL) < 1 and rho(F_FF) < 1, each by a power with 1-norm < 1 (see
        // gecon_cr_args.lead_idx).  T_LL = the lag block of T (rows o0.., packed columns 0..), F_FF = (W^-1 C)[lead][:, lead].
        bool certified = false;
        if (have_F) {
            const int s1 = min(w0, n - o0);
            // F_FF -> registers first (its source, the Xb block, is where the scratch tiles go)
            constexpr int PERZ = (8 * C * LDC + 31) / 32;
            double zf[PERZ];
#pragma unroll
            for (int e = 0; e < PERZ; ++e) {
                const int i = lane + 32 * e;
                const int r = i / LDC, c = i - r * LDC;
                double z = 0.0;
                if (r < nl && c < nl) {
                    const int pc = s_lead[c] - o2;
                    z = (pc >= 0 && pc < w2) ? X2[s_piv[s_lead[r]] * LDW + pc] : 0.0;
                }
                zf[e] = z;
            }
            __syncwarp();
            double* Z = WA;  // two 8 C x LDC tiles inside WA; F_FF goes where T was once its lag block has been copied out
            double* Zb = WA + 8 * C * LDC;
            double* Zf = A1;
#pragma unroll 1
            for (int i = lane; i < 8 * C * LDC; i += 32) {
                const int r = i / LDC, c = i - r * LDC;
                Z[i] = (r < s1 && c < s1) ? Tt[(o0 + r) * LDC + c] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < PERZ; ++e) {
                const int i = lane + 32 * e;
                if (i < 8 * C * LDC) Zf[i] = zf[e];
            }
            __syncwarp();
            certified = cw_power_bound<C>(Z, Zb, s1, lane) && cw_power_bound<C>(Zf, Z, nl, lane);
        }
        if (certified) status |= GECON_ST_BK_CERTIFIED;
        if (lane == 0) {
            p.status[draw] = p.accumulate ? (p.status[draw] | status) : status;
            if (p.lead_idx && p.n_unstable) p.n_unstable[draw] = certified ? nl : -1;
        }
        __syncwarp();
    }
}

}  // namespace gecon
