// Eigenvalues of batched real general matrices, one WARP per matrix, the matrix resident in shared memory.
//
// Replaces the eigenvalue call behind gEconpy/pytensorf/real_eig.py:10-36 (RealEig.perform: numpy.linalg.eig, eigenvalues
// only on the forward path) and gEconpy/model/perturbation.py:448-505 (compute_bk_eigenvalues_pt: eigenvalues of
// M = (-Gamma0_sel + 1e-8 I)^-1 Gamma1_sel), SURVEY.md Appendix A.4 option (1).  Diagnostics, not the hot path: the
// likelihood pipeline only needs the COUNT of |lambda| > 1 (bk_count.cu); this kernel gives the per-eigenvalue table of
// check_bk_condition(return_value="dataframe") and lets the count be cross-checked.
//
// Algorithm (the LAPACK dgeev / EISPACK route for eigenvalues only):
//   1  balancing by powers of two (Parlett-Reinsch scaling; the regularised Sims pencil has entries of order 1e8)
//   2  Householder reduction to upper Hessenberg form; reflectors applied with one lane per column / per row
//   3  Francis double-shift QR sweeps with deflation and the two exceptional shifts (EISPACK hqr): the scalar logic runs
//      redundantly in every lane on broadcast shared-memory reads, the row / column modifications of a sweep step are spread
//      over the lanes
// Output order is the deflation order; the host wrappers sort by modulus as RealEig does.
#include "common.cuh"

namespace gecon {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__global__ void __launch_bounds__(32) real_eig_kernel(const double* __restrict__ M, long long N, int m, int balance, double* __restrict__ re,
                                                      double* __restrict__ im, int* __restrict__ status) {
    // (m, H, wr, wi are re-pointed at the active block of the balanced matrix while a matrix is being worked on)
    extern __shared__ __align__(16) double sm_eig[];
    const int ld = m | 1;  // odd leading dimension: rows and columns are both conflict-free
    double* H = sm_eig;
    double* ort = H + (size_t)m * ld;  // [m] Householder vector
    double* wr = ort + m;              // [m]
    double* wi = wr + m;               // [m]
    const int m_in = m;
    const int lane = threadIdx.x;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    for (long long mat = blockIdx.x; mat < N; mat += gridDim.x) {
        m = m_in;
        const double* g = M + (size_t)mat * m * m;
        bool finite = true;
        for (int idx = lane; idx < m * m; idx += 32) {
            const int i = idx / m, j = idx - i * m;
            const double v = g[idx];
            finite = finite && (fabs(v) <= 1.7e308);
            H[i * ld + j] = v;
        }
        finite = __all_sync(0xffffffffu, finite);
        __syncwarp();
        int st = 0;
        if (!finite) {
            st = GECON_ST_LL_NONFINITE;
        } else {
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
                        if (warp_sum(r) == 0.0) {
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
                        if (warp_sum(c) == 0.0) {
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
            const int m_full = m;
            m = hi - lo + 1;
            if (balance) {
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
                        c = warp_sum(c);
                        r = warp_sum(r);
                        if (c != 0.0 && r != 0.0) {
                            double gg = r * 0.5, f = 1.0;
                            const double s = c + r;
                            while (c < gg) {
                                f *= 2.0;
                                c *= 4.0;
                            }
                            gg = r * 2.0;
                            while (c > gg) {
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
            // ---- 2. Householder reduction to upper Hessenberg form
            for (int k = 0; k + 2 < m; ++k) {
                double sc = 0.0;
                for (int i = k + 1 + lane; i < m; i += 32) sc += fabs(H[i * ld + k]);
                sc = warp_sum(sc);
                if (sc == 0.0) continue;  // warp-uniform
                double h = 0.0;
                for (int i = k + 1 + lane; i < m; i += 32) {
                    const double v = H[i * ld + k] / sc;
                    ort[i] = v;
                    h += v * v;
                }
                h = warp_sum(h);
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
            anorm = warp_sum(anorm);
            int nn = m - 1;
            double t = 0.0;
            bool failed = false;
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
            if (failed) st = GECON_ST_BK_INCONCLUSIVE;
            H = Hfull;
            wr = wr_full;
            wi = wi_full;
            m = m_full;
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            re[(size_t)mat * m + i] = st ? qnan : wr[i];
            im[(size_t)mat * m + i] = st ? qnan : wi[i];
        }
        if (lane == 0 && status) status[mat] = st;
        __syncwarp();
    }
}

static int check_eig(const double* M, long long N, int m, double* re, double* im) {
    if (!M || !re || !im || N < 0 || m < 1) {
        set_last_error("gecon_real_eig: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (m > 160) {
        set_last_error("gecon_real_eig: m = %d > 160 (the matrix lives in shared memory)", m);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_real_eig_batched(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status, void* stream) {
    int rc = check_eig(M, N, m, re, im);
    if (rc) return rc;
    if (N == 0) return 0;
    const size_t smem = sizeof(double) * ((size_t)m * (m | 1) + 3 * (size_t)m);
    int grid = 0;
    rc = persistent_grid(real_eig_kernel, 32, smem, N, &grid, nullptr);
    if (rc) return rc;
    real_eig_kernel<<<grid, 32, smem, (cudaStream_t)stream>>>(M, N, m, balance, re, im, status);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gecon_real_eig_host(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status) {
    int rc = check_eig(M, N, m, re, im);
    if (rc) return rc;
    if (N == 0) return 0;
    DevBuf dM, dRe, dIm, dSt;
    const size_t bm = (size_t)N * m * m * sizeof(double), bv = (size_t)N * m * sizeof(double);
    GECON_CUDA(dM.alloc(bm));
    GECON_CUDA(dRe.alloc(bv));
    GECON_CUDA(dIm.alloc(bv));
    GECON_CUDA(dSt.alloc((size_t)N * sizeof(int32_t)));
    GECON_CUDA(cudaMemcpy(dM.p, M, bm, cudaMemcpyHostToDevice));
    rc = gecon_real_eig_batched(dM.as<double>(), N, m, balance, dRe.as<double>(), dIm.as<double>(), dSt.as<int32_t>(), nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(re, dRe.p, bv, cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(im, dIm.p, bv, cudaMemcpyDeviceToHost));
    if (status) GECON_CUDA(cudaMemcpy(status, dSt.p, (size_t)N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return 0;
}
