// Device-side fp64 dense linear algebra for one parameter draw per CTA (sm_100a).
//
// Every n x n state matrix of a draw lives in shared memory as an NP x LD tile (NP = n rounded up to a
// multiple of 8, LD = NP + 4, zero padded).  LD = 4 (mod 8) makes the four fragment access patterns of
// mma.sync.m8n8k4.f64 (A, A^T, B, B^T operands read from row-major tiles) bank-conflict free: within a
// half-warp the 16 eight-byte words land in 16 distinct eight-byte banks.
//
// A CTA has NW = NP/8 warps; warp w owns the 8-row strip [8w, 8w+8) of every GEMM output, so an output
// tile is read-modify-written by exactly one warp and `C op= A*B` can run in place.
//
// GEMM-shaped steps use the fp64 tensor path (DMMA, mma.sync.m8n8k4.f64): measured on B200 it feeds from
// shared memory at 34.5 TFLOP/s against 23 TFLOP/s for register-tiled DFMA (profiles/r01_fp64_peak_microbench.json).
// Elimination, norms and the small p x p observation algebra use plain DFMA.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/gecon_b200.h"

namespace gecon {

template <int NP>
struct Cfg {
    static_assert(NP % 8 == 0 && NP >= 8 && NP <= 96, "NP must be a multiple of 8 in [8, 96]");
    static constexpr int LD = NP + 4;
    static constexpr int TILE = NP * LD;  // doubles per tile (even, so consecutive tiles stay 16-byte aligned)
    static constexpr int NW = NP / 8;     // warps per CTA
    static constexpr int NT = NW * 32;    // threads per CTA
    static constexpr int CT = NP / 8;     // 8-column tiles per strip
    static constexpr int CH = (NP + 31) / 32;  // 32-lane chunks per row
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------ accumulators
template <int NP>
struct Acc {
    double v[NP / 8][2];
};

template <int NP>
__device__ __forceinline__ void acc_zero(Acc<NP>& acc) {
#pragma unroll
    for (int ct = 0; ct < NP / 8; ++ct) acc.v[ct][0] = acc.v[ct][1] = 0.0;
}

// element owned by this lane in column tile ct:  row 8*warp + lane/4, columns 8*ct + 2*(lane%4) + {0,1}
template <int NP>
__device__ __forceinline__ void acc_load(Acc<NP>& acc, const double* __restrict__ Cm) {
    constexpr int LD = Cfg<NP>::LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* row = Cm + (warp * 8 + (lane >> 2)) * LD + 2 * (lane & 3);
#pragma unroll
    for (int ct = 0; ct < NP / 8; ++ct) {
        const double2 t = *reinterpret_cast<const double2*>(row + ct * 8);
        acc.v[ct][0] = t.x;
        acc.v[ct][1] = t.y;
    }
}

template <int NP>
__device__ __forceinline__ void acc_store(const Acc<NP>& acc, double* __restrict__ Cm, int ct_lo = 0, int ct_hi = NP / 8) {
    constexpr int LD = Cfg<NP>::LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* row = Cm + (warp * 8 + (lane >> 2)) * LD + 2 * (lane & 3);
#pragma unroll
    for (int ct = 0; ct < NP / 8; ++ct)
        if (ct >= ct_lo && ct < ct_hi) *reinterpret_cast<double2*>(row + ct * 8) = make_double2(acc.v[ct][0], acc.v[ct][1]);
}

// acc += sign * op(A) * op(B) over k in [klo, khi) (multiples of 4), output column tiles [ct_lo, ct_hi).
// TA: A is stored transposed (A_s[k][m]);  TB: B is stored transposed (B_s[n][k]), i.e. acc += A * B_s^T.
template <int NP, bool TA, bool TB>
__device__ __forceinline__ void gemm_acc(Acc<NP>& acc, const double* __restrict__ A, const double* __restrict__ B, double sign,
                                         int klo = 0, int khi = NP, int ct_lo = 0, int ct_hi = NP / 8) {
    constexpr int LD = Cfg<NP>::LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int r = warp * 8 + g;
    for (int k0 = klo; k0 < khi; k0 += 4) {
        const double a = sign * (TA ? A[(k0 + q) * LD + r] : A[r * LD + k0 + q]);
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct) {
            if (ct >= ct_lo && ct < ct_hi) {
                const double b = TB ? B[(ct * 8 + g) * LD + k0 + q] : B[(k0 + q) * LD + ct * 8 + g];
                dmma884(acc.v[ct][0], acc.v[ct][1], a, b);
            }
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// Compact Jacobian (gecon_compact_jac) -> dense row-major scratch in global memory, by the whole CTA: matrices `first..last`
// (0 = A, 1 = B, 2 = C: n x n; 3 = D: n x k) at dst + q * n * n.  The CTA-per-draw kernels expand the draw they are about to
// work on into a per-CTA scratch (it stays in L2) and then read it exactly like caller-provided dense matrices.
// Contains barriers: every thread of the CTA must call it.
__device__ __forceinline__ void expand_compact(double* __restrict__ dst, const gecon_compact_jac& cj, long long draw, int n, int k, int first,
                                               int last) {
    const int nt = blockDim.x;
    const size_t total = (size_t)(last < 3 ? (last - first + 1) * n * n : (3 - first) * n * n + n * k);
    double* base = dst + (size_t)first * n * n;
    for (size_t i = threadIdx.x; i < total; i += nt) base[i] = 0.0;
    __syncthreads();
    const double* v = cj.vals + (size_t)draw * cj.stride;
    for (int q = first; q <= last; ++q) {
        const int ldq = (q == 3) ? k : n;
        double* m = dst + (size_t)q * n * n;
        for (int e = cj.off[q] + threadIdx.x; e < cj.off[q + 1]; e += nt) {
            const int rc = cj.table[e];
            m[(rc >> 16) * ldq + (rc & 0xffff)] = v[e];
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ tile helpers
template <int NP>
__device__ __forceinline__ void tile_zero(double* __restrict__ dst) {
    for (int i = threadIdx.x; i < Cfg<NP>::TILE; i += Cfg<NP>::NT) dst[i] = 0.0;
}

template <int NP>
__device__ __forceinline__ void tile_copy(double* __restrict__ dst, const double* __restrict__ src) {
    for (int i = threadIdx.x; i < Cfg<NP>::TILE; i += Cfg<NP>::NT) dst[i] = src[i];
}

// global row-major rows x cols (leading dimension ldg) -> zero-padded tile
template <int NP>
__device__ __forceinline__ void tile_load(double* __restrict__ dst, const double* __restrict__ src, int rows, int cols, int ldg) {
    constexpr int LD = Cfg<NP>::LD;
    for (int i = threadIdx.x; i < Cfg<NP>::TILE; i += Cfg<NP>::NT) {
        const int r = i / LD, c = i - r * LD;
        dst[i] = (r < rows && c < cols) ? src[(size_t)r * ldg + c] : 0.0;
    }
}

// tile -> global, optionally through a row/column gather: out[i][j] = tile[rp[i]][cp[j]] (rp/cp may be null)
template <int NP>
__device__ __forceinline__ void tile_store(double* __restrict__ dst, const double* __restrict__ src, int rows, int cols, int ldg,
                                           double scale, const int* __restrict__ rp, const int* __restrict__ cp) {
    constexpr int LD = Cfg<NP>::LD;
    for (int i = threadIdx.x; i < rows * cols; i += Cfg<NP>::NT) {
        const int r = i / cols, c = i - r * cols;
        const int sr = rp ? rp[r] : r, sc = cp ? cp[c] : c;
        dst[(size_t)r * ldg + c] = scale * src[sr * LD + sc];
    }
}

// max absolute column sum (numpy.linalg.norm(M, ord=1)); NaN-propagating.  s_red: >= NP doubles of scratch.
// Contains barriers: every thread of the CTA must call it.
template <int NP>
__device__ double norm1(const double* __restrict__ M, int n, double* __restrict__ s_red) {
    constexpr int LD = Cfg<NP>::LD;
    if ((int)threadIdx.x < n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += fabs(M[i * LD + threadIdx.x]);
        s_red[threadIdx.x] = s;
    }
    __syncthreads();
    double mx = 0.0;
    for (int c = 0; c < n; ++c) {
        const double v = s_red[c];
        if (v > mx || v != v) mx = v;  // once NaN, stays NaN
    }
    __syncthreads();
    return mx;
}

// norm1 for full tiles whose padding rows and columns are zero, all 4 NP threads busy: four partial column sums per
// column, then an exact max of the (non-negative) sums through redux.sync on their bit patterns; NaN-propagating.
// s_red: 4 NP doubles.  One barrier; NO trailing barrier: a barrier must separate two uses of the same s_red.
template <int NP>
__device__ __forceinline__ double norm1_fast(const double* __restrict__ M, double* __restrict__ s_red) {
    constexpr int LD = Cfg<NP>::LD, RP = NP / 4, CH = Cfg<NP>::CH;
    const int part = threadIdx.x / NP, c = threadIdx.x - part * NP;
    const double* col = M + part * RP * LD + c;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < RP; ++i) s += fabs(col[i * LD]);
    s_red[threadIdx.x] = s;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned long long key = 0ull;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
        const int cc = lane + 32 * ch;
        if (cc < NP) {
            const double v = (s_red[cc] + s_red[NP + cc]) + (s_red[2 * NP + cc] + s_red[3 * NP + cc]);
            const unsigned long long kk = (unsigned long long)__double_as_longlong(fabs(v));
            key = kk > key ? kk : key;
        }
    }
    const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, (khi == mhi) ? klo : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

// ------------------------------------------------------------------------------------------------ linear solve
// acc += sign * A * B[bmap[.], :]: like gemm_acc<NP, false, false>, but row k of the B operand is read from row bmap[k]
// of the tile.  Lets the products of the cycle-reduction step consume the row-permuted output of gj_solve_blocked
// (solution row k lives in row piv[k]) without an un-permutation pass.
template <int NP>
__device__ __forceinline__ void gemm_acc_bmap(Acc<NP>& acc, const double* __restrict__ A, const double* __restrict__ B,
                                              const int* __restrict__ bmap, double sign, int klo, int khi, int ct_lo, int ct_hi) {
    constexpr int LD = Cfg<NP>::LD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int r = warp * 8 + g;
    for (int k0 = klo; k0 < khi; k0 += 4) {
        const double a = sign * A[r * LD + k0 + q];
        const double* brow = B + bmap[k0 + q] * LD + g;
#pragma unroll
        for (int ct = 0; ct < NP / 8; ++ct) {
            if (ct >= ct_lo && ct < ct_hi) dmma884(acc.v[ct][0], acc.v[ct][1], a, brow[ct * 8]);
        }
    }
}

// Blocked Gauss-Jordan elimination with partial pivoting, panels of 8 columns, trailing updates on the fp64 tensor
// path:  X = M^-1 [R1 | R2] for the 8-column tiles [c1lo, c1hi) of R1 and [c2lo, c2hi) of R2.
//
// Block step kb (columns J = [8 kb, 8 kb + 8)):
//   panel   warp 0, lane = row (two rows per lane for NP > 32), the row's 8 panel entries in registers.  In-place
//           Gauss-Jordan inversion of the panel: for each column the pivot is the entry of largest magnitude among the
//           rows not used yet (lowest row on ties: the row getrf picks), found with redux.sync; the pivot row travels
//           by shuffles.  Rows are never swapped.  The panel ends up holding W (n x 8) with
//               new_row_i = [i not a pivot row of this panel] old_row_i + sum_a W[i][a] old_row_{piv[8 kb + a]}
//           (W = -multipliers for ordinary rows, the inverse of the 8 x 8 pivot block for the pivot rows).
//   update  every warp, its 8-row strip: that formula as one DMMA product with k = 8 for every live column tile
//           (tiles of M to the right of the panel, the right-hand-side tiles); pivot rows are gathered through piv[].
// Three barriers per block step instead of one per column, and the O(n^3) work moves from load/DFMA/store triples to
// mma.sync.  The pivots are the ones of unblocked partial pivoting (1 / pivot_j -> s_pivinv[j] when given, so det M = +- prod
// pivots).  The first step may read from a different set of tiles than it writes (Ms/R1s/R2s -> Md/R1d/R2d: no
// copy of the inputs is needed, and that step needs no barrier between reading and writing); the remaining steps work
// in place in the destination tiles.  On exit M is destroyed and solution row j is in row piv[j] of the right-hand-
// side tiles (piv -> s_piv); with `unpermute` the rows are moved to their natural positions.  Destination columns
// outside the tile ranges are not written.  Returns false on a zero / non-finite pivot (caller NaN-fills, as the
// reference's _solve_gen does).  Contains barriers: every thread of the CTA must call it with identical arguments.
// s_piv: NP ints, s_flag: NP + 1 ints.
// reciprocal to <= 1 ulp: MUFU.RCP64H seed + three Newton steps (the pivots only need a consistent, accurate 1/x;
// IEEE division's special-case handling is not wanted inside the elimination's critical path)
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// two Newton steps: MUFU.RCP64H is good to ~2^-20, so 2^-40 after one step and below half an ulp after two (what the CUDA
// math library's own reciprocal does on its fast path); used where the reciprocal sits on a dependent chain
__device__ __forceinline__ double rcp_nr2(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// 64-bit shuffle as two explicit 32-bit shuffles (the double overload of __shfl_sync costs three extra LOP3 per value)
__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src);
    const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src);
    return __hiloint2double(hi, lo);
}

template <int NP>
__device__ __noinline__ bool gj_solve_blocked(const double* Ms, double* Md, const double* R1s, double* R1d, int c1lo, int c1hi,
                                              const double* R2s, double* R2d, int c2lo, int c2hi, int n, bool unpermute,
                                              int* __restrict__ s_piv, int* __restrict__ s_flag, double* __restrict__ s_pivinv = nullptr) {
    constexpr int LD = Cfg<NP>::LD, CT = Cfg<NP>::CT, RPL = (NP + 31) / 32, NW = Cfg<NP>::NW;
    constexpr unsigned IDXBITS = (RPL == 1) ? 5u : (RPL == 2) ? 6u : 7u, IDXMASK = (1u << IDXBITS) - 1u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    const int nblk = (n + 7) >> 3;
    bool used[RPL];
#pragma unroll
    for (int h = 0; h < RPL; ++h) used[h] = (lane + 32 * h >= n);
    for (int kb = 0; kb < nblk; ++kb) {
        const int c0 = 8 * kb;
        const bool first = (kb == 0);
        const double* Mc = first ? Ms : Md;
        // ------------------------------------------------------------------------------------------ panel (warp 0)
        if (warp == 0) {
            const int jmax = min(8, n - c0);
            double a[RPL][8];
            bool inP[RPL];
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                const int i = lane + 32 * h;
                inP[h] = false;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 t = make_double2(0.0, 0.0);
                    if (i < NP) t = *reinterpret_cast<const double2*>(Mc + i * LD + c0 + c);
                    a[h][c] = t.x;
                    a[h][c + 1] = t.y;
                }
            }
            unsigned fail = 0u;
            int myr = 0;           // lane jj keeps the pivot row of panel column jj (0 for padding columns: any valid row)
            double mypinv = 1.0;   // ... and 1 / pivot of that column (callers that want det M)
            double myinv[RPL];     // 1 / pivot of the row(s) this lane holds, once they have been pivot rows
#pragma unroll
            for (int h = 0; h < RPL; ++h) myinv[h] = 1.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                if (jj < jmax) {  // warp-uniform
                    // pivot = largest |.| among the rows not used yet, lowest row on ties.  One redux on a 32-bit key:
                    // the high word of |x| (sign, exponent, 20 - IDXBITS leading mantissa bits: candidates closer than
                    // ~2^-14 relative count as tied) above the complemented row index.
                    unsigned key = 0u;
                    double invo[RPL];
#pragma unroll
                    for (int h = 0; h < RPL; ++h) {
                        const unsigned kk = ((unsigned)__double2hiint(fabs(a[h][jj])) & ~IDXMASK) | (IDXMASK - (unsigned)(lane + 32 * h));
                        key = max(key, used[h] ? 0u : kk);
                        invo[h] = rcp_nr(a[h][jj]);  // every row's own reciprocal, off the pivot search's critical path
                    }
                    const unsigned best = __reduce_max_sync(0xffffffffu, key);
                    const int r = (int)(IDXMASK - (best & IDXMASK));
                    fail |= ((best >> IDXBITS) == 0u || best >= 0x7ff00000u) ? 1u : 0u;  // zero / subnormal / inf / NaN pivot
                    const int rl = r & 31;
                    const int rh = r >> 5;
                    double inv = invo[0];
                    if constexpr (RPL >= 2) inv = (rh == 1) ? invo[1] : inv;
                    if constexpr (RPL >= 3) inv = (rh == 2) ? invo[2] : inv;
                    inv = shfl_f64(inv, rl);
                    double pr[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        if (c != jj) {
                            double v = a[0][c];
                            if constexpr (RPL >= 2) v = (rh == 1) ? a[1][c] : v;
                            if constexpr (RPL >= 3) v = (rh == 2) ? a[2][c] : v;
                            pr[c] = shfl_f64(v, rl);
                        }
                    }
                    // The pivot row is NOT scaled here (its slot takes the coefficient 1 of its own old row); it is
                    // scaled by 1/pivot once, after the panel, which commutes with the later eliminations acting on it.
#pragma unroll
                    for (int h = 0; h < RPL; ++h) {
                        const bool is_r = (lane + 32 * h == r);
                        const double m = is_r ? 0.0 : a[h][jj] * inv;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (c != jj) a[h][c] = fma(-m, pr[c], a[h][c]);
                        }
                        a[h][jj] = is_r ? 1.0 : -m;
                        myinv[h] = is_r ? inv : myinv[h];
                        used[h] = used[h] || is_r;
                        inP[h] = inP[h] || is_r;
                    }
                    if (lane == jj) {
                        myr = r;
                        mypinv = inv;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < RPL; ++h) {
                const int i = lane + 32 * h;
#pragma unroll
                for (int c = 0; c < 8; ++c) a[h][c] *= myinv[h];  // 1 for the rows that were not pivot rows of this panel
                if (i < NP) {
#pragma unroll
                    for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(Md + i * LD + c0 + c) = make_double2(a[h][c], a[h][c + 1]);
                    s_flag[i] = inP[h] ? 1 : 0;
                }
            }
            if (lane < 8) s_piv[c0 + lane] = myr;
            if (s_pivinv && lane < 8) s_pivinv[c0 + lane] = mypinv;
            if (lane == 0) s_flag[NP] = (int)fail;
        }
        __syncthreads();
        if (s_flag[NP]) {
            __syncthreads();
            return false;
        }
        // ------------------------------------------------------------------------------------------ update (all warps)
        const int r = warp * 8 + g;
        const double a0 = Md[r * LD + c0 + q], a1 = Md[r * LD + c0 + 4 + q];
        const int p0 = s_piv[c0 + q] * LD + g, p1 = s_piv[c0 + 4 + q] * LD + g;
        const bool keep = (s_flag[r] == 0);
        const int ro = r * LD + 2 * q;
        const double* R1c = first ? R1s : R1d;
        const double* R2c = first ? R2s : R2d;
        double accM[CT][2], acc1[CT][2], acc2[CT][2];
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            if (ct > kb && ct < nblk) {
                double2 o = make_double2(0.0, 0.0);
                if (keep) o = *reinterpret_cast<const double2*>(Mc + ro + ct * 8);
                accM[ct][0] = o.x;
                accM[ct][1] = o.y;
                dmma884(accM[ct][0], accM[ct][1], a0, Mc[p0 + ct * 8]);
                dmma884(accM[ct][0], accM[ct][1], a1, Mc[p1 + ct * 8]);
            }
            if (ct >= c1lo && ct < c1hi) {
                double2 o = make_double2(0.0, 0.0);
                if (keep) o = *reinterpret_cast<const double2*>(R1c + ro + ct * 8);
                acc1[ct][0] = o.x;
                acc1[ct][1] = o.y;
                dmma884(acc1[ct][0], acc1[ct][1], a0, R1c[p0 + ct * 8]);
                dmma884(acc1[ct][0], acc1[ct][1], a1, R1c[p1 + ct * 8]);
            }
            if (ct >= c2lo && ct < c2hi) {
                double2 o = make_double2(0.0, 0.0);
                if (keep) o = *reinterpret_cast<const double2*>(R2c + ro + ct * 8);
                acc2[ct][0] = o.x;
                acc2[ct][1] = o.y;
                dmma884(acc2[ct][0], acc2[ct][1], a0, R2c[p0 + ct * 8]);
                dmma884(acc2[ct][0], acc2[ct][1], a1, R2c[p1 + ct * 8]);
            }
        }
        // in place: every warp must have read the pivot rows before anybody overwrites them
        const bool oop = first && Ms != Md && (c1lo >= c1hi || R1s != R1d) && (c2lo >= c2hi || R2s != R2d);
        if (!oop) __syncthreads();
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            if (ct > kb && ct < nblk) *reinterpret_cast<double2*>(Md + ro + ct * 8) = make_double2(accM[ct][0], accM[ct][1]);
            if (ct >= c1lo && ct < c1hi) *reinterpret_cast<double2*>(R1d + ro + ct * 8) = make_double2(acc1[ct][0], acc1[ct][1]);
            if (ct >= c2lo && ct < c2hi) *reinterpret_cast<double2*>(R2d + ro + ct * 8) = make_double2(acc2[ct][0], acc2[ct][1]);
        }
        __syncthreads();
    }
    if (unpermute) {
        // X[j] = rows[piv[j]]: rows j = warp + e * NW (e < 8 covers NP rows), through registers
        constexpr int CH = Cfg<NP>::CH;
        const int w1 = 8 * (c1hi - c1lo), w2 = 8 * (c2hi - c2lo);
        double x1[8][CH], x2[8][CH];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = warp + e * NW;
            const int pr = (j < n) ? s_piv[j] : 0;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const int c = lane + 32 * ch;
                x1[e][ch] = (j < n && c < w1) ? R1d[pr * LD + 8 * c1lo + c] : 0.0;
                x2[e][ch] = (j < n && c < w2) ? R2d[pr * LD + 8 * c2lo + c] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = warp + e * NW;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                const int c = lane + 32 * ch;
                if (j < NP && c < w1) R1d[j * LD + 8 * c1lo + c] = x1[e][ch];
                if (j < NP && c < w2) R2d[j * LD + 8 * c2lo + c] = x2[e][ch];
            }
        }
        __syncthreads();
    }
    return true;
}

// [lo, hi) = range of columns of the n x n corner of M holding a non-zero (or NaN) entry; lo = hi = 0 if none.
// s_i: 2 ints of scratch.  Contains barriers.
template <int NP>
__device__ void nonzero_col_range(const double* __restrict__ M, int n, int* __restrict__ s_i, int& lo, int& hi) {
    constexpr int LD = Cfg<NP>::LD;
    if (threadIdx.x == 0) {
        s_i[0] = NP;
        s_i[1] = 0;
    }
    __syncthreads();
    if ((int)threadIdx.x < n) {
        bool nz = false;
        for (int i = 0; i < n; ++i) nz |= (M[i * LD + threadIdx.x] != 0.0);
        if (nz) {
            atomicMin(&s_i[0], (int)threadIdx.x);
            atomicMax(&s_i[1], (int)threadIdx.x + 1);
        }
    }
    __syncthreads();
    lo = s_i[0];
    hi = s_i[1];
    if (lo >= hi) lo = hi = 0;
    __syncthreads();
}

// fill the n x m corner of a tile with NaN (reference: LAPACK failure -> NaN fill, cycle_reduction.py:179-181)
template <int NP>
__device__ __forceinline__ void tile_nanfill(double* __restrict__ dst, int n, int m) {
    constexpr int LD = Cfg<NP>::LD;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = threadIdx.x; i < n * m; i += Cfg<NP>::NT) {
        const int r = i / m, c = i - r * m;
        dst[r * LD + c] = qnan;
    }
}

// CTA-wide reductions through shared memory.  s_red: >= 8 doubles.  Contain barriers.
template <int NP>
__device__ double block_sum(double v, double* __restrict__ s_red) {
    constexpr int NW = Cfg<NP>::NW;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += s_red[w];
    __syncthreads();
    return s;
}

template <int NP>
__device__ double block_max(double v, double* __restrict__ s_red) {  // NaN-propagating max of non-negative values
    constexpr int NW = Cfg<NP>::NW;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, off);
        if (o > v || o != o) v = o;
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < NW; ++w) {
        const double o = s_red[w];
        if (o > s || o != o) s = o;
    }
    __syncthreads();
    return s;
}

// in-place symmetrisation of the n x n corner: M = (M + M^T)/2
template <int NP>
__device__ __forceinline__ void tile_symmetrize(double* __restrict__ M, int n) {
    constexpr int LD = Cfg<NP>::LD;
    for (int idx = threadIdx.x; idx < n * n; idx += Cfg<NP>::NT) {
        const int i = idx / n, j = idx - i * n;
        if (i < j) {
            const double s = 0.5 * (M[i * LD + j] + M[j * LD + i]);
            M[i * LD + j] = s;
            M[j * LD + i] = s;
        }
    }
}

}  // namespace gecon
