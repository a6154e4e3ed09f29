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

namespace gecon {

template <int NP>
struct Cfg {
    static_assert(NP % 8 == 0 && NP >= 8 && NP <= 64, "NP must be a multiple of 8 in [8, 64]");
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

// ------------------------------------------------------------------------------------------------ linear solve
// Gauss-Jordan elimination with implicit partial pivoting: X = M^{-1} [R1 | R2], in place in the column ranges
// [lo1, hi1) of R1 and [lo2, hi2) of R2 (columns outside the ranges are not touched: callers pass the range that
// holds the non-zero columns of the right-hand side, whose other solution columns are zero); M (n x n) is destroyed.  The pivot of column j is the entry of largest magnitude among the rows
// not used yet (first index on ties) -- the row LAPACK's getrf picks -- so the result agrees with an LU solve to
// rounding.  Rows are never swapped: the pivot row r_j and 1/pivot are recorded and the solution rows are
// un-permuted and scaled at the end (X[j] = R[r_j] / pivot_j), through registers.
// One barrier per column.  Returns false on a zero or non-finite pivot (caller NaN-fills, as the reference's
// _solve_gen does).  Contains barriers: every thread of the CTA must call it with identical arguments.
template <int NP>
__device__ bool gj_solve(double* M, double* R1, int lo1, int hi1, double* R2, int lo2, int hi2, int n, int* __restrict__ s_piv,
                         double* __restrict__ s_inv) {
    constexpr int LD = Cfg<NP>::LD, NW = Cfg<NP>::NW, CH = Cfg<NP>::CH;
    constexpr int SL = (3 * NP + 31) / 32;  // slots per lane over the concatenated live columns [M | R1 | R2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w1 = hi1 - lo1, w2 = hi2 - lo2;
    // R1 / R2 live in the same shared-memory array as M: address them as offsets from M so that one slot loop serves all
    const int off1 = (w1 > 0) ? (int)(R1 - M) + lo1 : 0;
    const int off2 = (w2 > 0) ? (int)(R2 - M) + lo2 : 0;
    // ---- slots: lane q + 32 s of the concatenated columns [all of M | R1 range | R2 range], fixed for the whole solve.
    // A lane without a column in a slot points at the tile's first padding column (index NP: never read as matrix data)
    // with a zero pivot-row value, so that the update loop runs without per-element predicates.
    const int ns = (n + w1 + w2 + 31) >> 5;
    int coff[SL], cmin[SL];  // cmin: M column index (live once > j); RHS columns are always live (NP + 1)
#pragma unroll
    for (int s = 0; s < SL; ++s) {
        const int q = lane + 32 * s;
        int c = NP, cm = -1;
        if (q < n) {
            c = q;
            cm = q;
        } else if (q < n + w1) {
            c = off1 + (q - n);
            cm = NP + 1;
        } else if (q < n + w1 + w2) {
            c = off2 + (q - n - w1);
            cm = NP + 1;
        }
        coff[s] = c;
        cmin[s] = cm;
    }
    unsigned long long used = 0ull;
    bool ok = true;
    for (int j = 0; j < n; ++j) {
        // ---- pivot search, done redundantly by every warp on identical data (no barrier needed to share it).
        // |x| as an orderable 64-bit key; max over the warp with two 32-bit redux.sync, lowest row index on ties.
        unsigned long long key = 0ull;
        int row = NP;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            const int i = lane + 32 * ch;
            if (i < n && !((used >> i) & 1ull)) {
                const unsigned long long kk = (unsigned long long)__double_as_longlong(fabs(M[i * LD + j]));
                if (kk > key || row == NP) {
                    key = kk;
                    row = i;
                }
            }
        }
        const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, khi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, (khi == mhi) ? klo : 0u);
        const int r = (int)__reduce_min_sync(0xffffffffu, (unsigned)((khi == mhi && klo == mlo) ? row : NP));
        const double best = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
        if (r >= n || !(best > 0.0) || best > 1.7e308) {
            ok = false;
            break;  // uniform across the CTA
        }
        const double inv = 1.0 / M[r * LD + j];
        used |= 1ull << r;
        if (threadIdx.x == 0) {
            s_piv[j] = r;
            s_inv[j] = inv;
        }
        // ---- pivot-row values of this lane's slots; columns of M up to j are dead: a zero value makes their update an
        // exact no-op (x - m * 0 = x, so the concurrent multiplier reads of column j by other warps see the same bits)
        double pv[SL];
#pragma unroll
        for (int s = 0; s < SL; ++s) pv[s] = (s < ns && cmin[s] > j) ? M[r * LD + coff[s]] : 0.0;
        // ---- eliminate column j from every other row; warps split rows (two at a time for ILP), lanes split columns.
        // Row r and rows with a zero multiplier get multiplier 0 (an exact no-op) instead of a branch.
        for (int i0 = warp; i0 < n; i0 += 2 * NW) {
            const int i1 = i0 + NW;
            double* row0 = M + i0 * LD;
            const double m0 = (i0 == r) ? 0.0 : row0[j] * inv;
            if (i1 < n) {
                double* row1 = M + i1 * LD;
                const double m1 = (i1 == r) ? 0.0 : row1[j] * inv;
                if (m0 == 0.0 && m1 == 0.0) continue;  // structural zeros: nothing to do (reference BLAS skips them too)
                double x0[SL], x1[SL];
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    if (s < ns) {
                        x0[s] = row0[coff[s]];
                        x1[s] = row1[coff[s]];
                    }
                }
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    if (s < ns) {
                        row0[coff[s]] = fma(-m0, pv[s], x0[s]);
                        row1[coff[s]] = fma(-m1, pv[s], x1[s]);
                    }
                }
            } else {
                if (m0 == 0.0) continue;
#pragma unroll
                for (int s = 0; s < SL; ++s) {
                    if (s < ns) row0[coff[s]] = fma(-m0, pv[s], row0[coff[s]]);
                }
            }
        }
        __syncthreads();
    }
    if (!ok) {
        __syncthreads();
        return false;
    }
    // ---- un-permute and scale: X[j][c] = R[piv[j]][c] * inv[j]; rows j = warp + e*NW (e < 8 covers NP rows)
    double x1[8][CH], x2[8][CH];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = warp + e * NW;
        const int pr = (j < n) ? s_piv[j] : 0;
        const double iv = (j < n) ? s_inv[j] : 0.0;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            const int c = lane + 32 * ch;
            x1[e][ch] = (j < n && lo1 + c < hi1) ? R1[pr * LD + lo1 + c] * iv : 0.0;
            x2[e][ch] = (j < n && lo2 + c < hi2) ? R2[pr * LD + lo2 + c] * iv : 0.0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int j = warp + e * NW;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            const int c = lane + 32 * ch;
            if (j < n && lo1 + c < hi1) R1[j * LD + lo1 + c] = x1[e][ch];
            if (j < n && lo2 + c < hi2) R2[j * LD + lo2 + c] = x2[e][ch];
        }
    }
    __syncthreads();
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
