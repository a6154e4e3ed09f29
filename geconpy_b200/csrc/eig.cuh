// Eigenvalues of ONE real general matrix by ONE warp, the matrix resident in shared memory (device routine shared by
// real_eig_kernel, eig.cu, and by the Blanchard-Kahn count kernel's fallback, bk_count.cu).
//
// Algorithm (the LAPACK dgeev / EISPACK route for eigenvalues only):
//   1  balancing as dgebal: permutation that isolates eigenvalues, then scaling by powers of two (Parlett-Reinsch)
//   2  Householder reduction to upper Hessenberg form; reflectors applied with one lane per column / per row
//   3  Francis double-shift QR sweeps with deflation and the two exceptional shifts (EISPACK hqr): the scalar logic runs
//      redundantly in every lane on broadcast shared-memory reads, the row / column modifications of a sweep step are spread
//      over the lanes
#pragma once
#include "common.cuh"

namespace gecon {

__device__ __forceinline__ double eig_warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Parlett-Reinsch scaling by powers of two (the scaling half of dgebal; a diagonal similarity, exact in floating point): rows and
// columns of H (m x m, leading dimension ld, shared memory) are rescaled until every row / column norm pair is within a factor of
// two, at most 20 sweeps.  One warp; only __syncwarp inside.
static __device__ __noinline__ void warp_balance_scale(double* H, int ld, int m, int lane) {
    for (int pass = 0; pass < 20; ++pass) {
        bool last = true;
        for (int i = 0; i < m; ++i) {
            double c = 0.0, r = 0.0;
            for (int j = lane; j < m; j += 32) {
                if (j != i) {
                    c += fabs(H[j * ld + i]);
                    r += fabs(H[i * ld + j]);
                }
            }
            c = eig_warp_sum(c);
            r = eig_warp_sum(r);
            // (non-finite sums -- an inf / NaN entry, or overflow -- leave the row alone: inf * 0.25 never gets below anything)
            if (c != 0.0 && r != 0.0 && c <= 1.7e308 && r <= 1.7e308) {
                double gg = r * 0.5, f = 1.0;
                const double s = c + r;
                for (int guard = 0; c < gg && guard < 1100; ++guard) {
                    f *= 2.0;
                    c *= 4.0;
                }
                gg = r * 2.0;
                for (int guard = 0; c > gg && guard < 1100; ++guard) {
                    f *= 0.5;
                    c *= 0.25;
                }
                if ((c + r) / f < 0.95 * s) {
                    last = false;
                    const double gi = 1.0 / f;
                    for (int j = lane; j < m; j += 32) H[i * ld + j] *= gi;
                    __syncwarp();
                    for (int j = lane; j < m; j += 32) H[j * ld + i] *= f;
                    __syncwarp();
                }
            }
        }
        if (last) break;
    }
}

// H: m x m, row-major with leading dimension ld (destroyed); ort, wr, wi: m doubles each (shared memory).  All 32 lanes of
// the calling warp take part (only __syncwarp inside).  Returns false when a QR sweep does not converge (60 iterations for one
// eigenvalue); on success wr / wi hold the eigenvalues in deflation order.
static __device__ __noinline__ bool warp_real_eig(double* H, int ld, int m, int balance, double* ort, double* wr, double* wi, int lane) {
    bool failed = false;
    {
    // ---- 1. balancing, as dgebal: first PERMUTE so that rows / columns which isolate an eigenvalue move to the bottom / top
    // (their eigenvalues are then diagonal entries, exact), then scale the remaining block [lo, hi] by powers of two
    int lo = 0, hi = m - 1;
    if (balance) {
        auto exchange = [&](int a, int b) {  // similarity permutation: swap rows a, b and columns a, b
            if (a == b) return;
            for (int j = lane; j < m; j += 32) {
                const double t = H[a * ld + j];
                H[a * ld + j] = H[b * ld + j];
                H[b * ld + j] = t;
            }
            __syncwarp();
            for (int i = lane; i < m; i += 32) {
                const double t = H[i * ld + a];
                H[i * ld + a] = H[i * ld + b];
                H[i * ld + b] = t;
            }
            __syncwarp();
        };
        bool again = true;
        while (again && hi > lo) {  // rows with zero off-diagonal part inside the active block -> bottom
            again = false;
            for (int j = hi; j >= lo; --j) {
                double r = 0.0;
                for (int i = lo + lane; i <= hi; i += 32)
                    if (i != j) r += fabs(H[j * ld + i]);
                if (eig_warp_sum(r) == 0.0) {
                    exchange(j, hi);
                    --hi;
                    again = true;
                    break;
                }
            }
        }
        again = true;
        while (again && hi > lo) {  // columns with zero off-diagonal part inside the active block -> top
            again = false;
            for (int j = lo; j <= hi; ++j) {
                double c = 0.0;
                for (int i = lo + lane; i <= hi; i += 32)
                    if (i != j) c += fabs(H[i * ld + j]);
                if (eig_warp_sum(c) == 0.0) {
                    exchange(j, lo);
                    ++lo;
                    again = true;
                    break;
                }
            }
        }
    }
    for (int i = lane; i < m; i += 32) {  // isolated eigenvalues (overwritten below for the active block)
        wr[i] = H[i * ld + i];
        wi[i] = 0.0;
    }
    __syncwarp();
    double* const Hfull = H;
    double* const wr_full = wr;
    double* const wi_full = wi;
    (void)Hfull;
    // everything below works on the active block only
    H = Hfull + lo * ld + lo;
    wr = wr_full + lo;
    wi = wi_full + lo;
    m = hi - lo + 1;
    if (balance) warp_balance_scale(H, ld, m, lane);
    // ---- 2. Householder reduction to upper Hessenberg form
    for (int k = 0; k + 2 < m; ++k) {
        double sc = 0.0;
        for (int i = k + 1 + lane; i < m; i += 32) sc += fabs(H[i * ld + k]);
        sc = eig_warp_sum(sc);
        if (sc == 0.0) continue;  // warp-uniform
        double h = 0.0;
        for (int i = k + 1 + lane; i < m; i += 32) {
            const double v = H[i * ld + k] / sc;
            ort[i] = v;
            h += v * v;
        }
        h = eig_warp_sum(h);
        __syncwarp();
        const double o1 = ort[k + 1];
        const double gg = (o1 > 0.0) ? -sqrt(h) : sqrt(h);
        h -= o1 * gg;
        __syncwarp();
        if (lane == 0) ort[k + 1] = o1 - gg;
        __syncwarp();
        // (I - u u' / h) H : one lane per column j >= k
        for (int j = k + lane; j < m; j += 32) {
            double f = 0.0;
            for (int i = k + 1; i < m; ++i) f += ort[i] * H[i * ld + j];
            f /= h;
            for (int i = k + 1; i < m; ++i) H[i * ld + j] -= f * ort[i];
        }
        __syncwarp();
        // H (I - u u' / h) : one lane per row i
        for (int i = lane; i < m; i += 32) {
            double f = 0.0;
            for (int j = k + 1; j < m; ++j) f += ort[j] * H[i * ld + j];
            f /= h;
            for (int j = k + 1; j < m; ++j) H[i * ld + j] -= f * ort[j];
        }
        __syncwarp();
        if (lane == 0) H[(k + 1) * ld + k] = sc * gg;
        for (int i = k + 2 + lane; i < m; i += 32) H[i * ld + k] = 0.0;
        __syncwarp();
    }
    // ---- 3. Francis double-shift QR on the Hessenberg matrix (eigenvalues only)
    double anorm = 0.0;
    for (int idx = lane; idx < m * m; idx += 32) {
        const int i = idx / m, j = idx - i * m;
        if (j + 1 >= i) anorm += fabs(H[i * ld + j]);
    }
    anorm = eig_warp_sum(anorm);
    int nn = m - 1;
    double t = 0.0;
    failed = false;
    while (nn >= 0 && !failed) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                double s = fabs(H[(l - 1) * ld + l - 1]) + fabs(H[l * ld + l]);
                if (s == 0.0) s = anorm;
                if (fabs(H[l * ld + l - 1]) + s == s) {
                    __syncwarp();
                    if (lane == 0) H[l * ld + l - 1] = 0.0;
                    __syncwarp();
                    break;
                }
            }
            double x = H[nn * ld + nn];
            if (l == nn) {  // one root
                if (lane == 0) {
                    wr[nn] = x + t;
                    wi[nn] = 0.0;
                }
                --nn;
            } else {
                double y = H[(nn - 1) * ld + nn - 1];
                double w = H[nn * ld + nn - 1] * H[(nn - 1) * ld + nn];
                if (l == nn - 1) {  // two roots
                    const double p = 0.5 * (y - x);
                    const double q = p * p + w;
                    double z = sqrt(fabs(q));
                    x += t;
                    if (lane == 0) {
                        if (q >= 0.0) {
                            z = p + (p >= 0.0 ? fabs(z) : -fabs(z));
                            wr[nn - 1] = wr[nn] = x + z;
                            if (z != 0.0) wr[nn] = x - w / z;
                            wi[nn - 1] = wi[nn] = 0.0;
                        } else {
                            wr[nn - 1] = wr[nn] = x + p;
                            wi[nn - 1] = z;
                            wi[nn] = -z;
                        }
                    }
                    nn -= 2;
                } else {  // no root yet: one more sweep
                    if (its == 60) {
                        failed = true;
                        break;
                    }
                    if (its == 10 || its == 20 || its == 40) {  // exceptional shift
                        t += x;
                        __syncwarp();
                        for (int i = lane; i <= nn; i += 32) H[i * ld + i] -= x;
                        __syncwarp();
                        const double s = fabs(H[nn * ld + nn - 1]) + fabs(H[(nn - 1) * ld + nn - 2]);
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int mm;
                    double p = 0.0, q = 0.0, r = 0.0, z;
                    for (mm = nn - 2; mm >= l; --mm) {  // two consecutive small sub-diagonal elements
                        z = H[mm * ld + mm];
                        r = x - z;
                        double s = y - z;
                        p = (r * s - w) / H[(mm + 1) * ld + mm] + H[mm * ld + mm + 1];
                        q = H[(mm + 1) * ld + mm + 1] - z - r - s;
                        r = H[(mm + 2) * ld + mm + 1];
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s;
                        q /= s;
                        r /= s;
                        if (mm == l) break;
                        const double u = fabs(H[mm * ld + mm - 1]) * (fabs(q) + fabs(r));
                        const double v = fabs(p) * (fabs(H[(mm - 1) * ld + mm - 1]) + fabs(z) + fabs(H[(mm + 1) * ld + mm + 1]));
                        if (u + v == v) break;
                    }
                    __syncwarp();
                    for (int i = mm + 2 + lane; i <= nn; i += 32) {
                        H[i * ld + i - 2] = 0.0;
                        if (i != mm + 2) H[i * ld + i - 3] = 0.0;
                    }
                    __syncwarp();
                    for (int k = mm; k <= nn - 1; ++k) {  // double QR step on rows l..nn, columns mm..nn
                        if (k != mm) {
                            p = H[k * ld + k - 1];
                            q = H[(k + 1) * ld + k - 1];
                            r = (k != nn - 1) ? H[(k + 2) * ld + k - 1] : 0.0;
                            x = fabs(p) + fabs(q) + fabs(r);
                            if (x != 0.0) {
                                p /= x;
                                q /= x;
                                r /= x;
                            }
                        }
                        const double nrm = sqrt(p * p + q * q + r * r);
                        const double s = (p >= 0.0) ? nrm : -nrm;
                        if (s != 0.0) {
                            __syncwarp();
                            if (lane == 0) {
                                if (k == mm) {
                                    if (l != mm) H[k * ld + k - 1] = -H[k * ld + k - 1];
                                } else {
                                    H[k * ld + k - 1] = -s * x;
                                }
                            }
                            p += s;
                            x = p / s;
                            y = q / s;
                            z = r / s;
                            q /= p;
                            r /= p;
                            __syncwarp();
                            for (int j = k + lane; j <= nn; j += 32) {  // row modification
                                double pp = H[k * ld + j] + q * H[(k + 1) * ld + j];
                                if (k != nn - 1) {
                                    pp += r * H[(k + 2) * ld + j];
                                    H[(k + 2) * ld + j] -= pp * z;
                                }
                                H[(k + 1) * ld + j] -= pp * y;
                                H[k * ld + j] -= pp * x;
                            }
                            __syncwarp();
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l + lane; i <= mmin; i += 32) {  // column modification
                                double pp = x * H[i * ld + k] + y * H[i * ld + k + 1];
                                if (k != nn - 1) {
                                    pp += z * H[i * ld + k + 2];
                                    H[i * ld + k + 2] -= pp * r;
                                }
                                H[i * ld + k + 1] -= pp * q;
                                H[i * ld + k] -= pp;
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        } while (l < nn - 1 && !failed);
    }
    }
    return !failed;
}

}  // namespace gecon
