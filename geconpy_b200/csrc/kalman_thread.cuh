// Kalman-filter log-likelihood, ONE THREAD PER DRAW: filter dimension U <= 4, p <= 2 observables, selector Z.
//
// Same semantics as kalman_ll_warp_kernel (kalman_warp.cuh; pymc_extras StandardFilter as called from
// gEconpy/model/statespace.py:1151-1157, restated in oracle/statespace.py): update -> jitter -> predict, the Joseph update in its
// rank-p form P+ = P - K (P Z' + jitter K)' (F built from one triangle of P, P+ mirrored: the stable variant),
// missing observations masked, a0 = 0, P0 = dlyap(T, R Q R') by Smith doubling.
//
// Why a third kernel: textbook-sized models (the RBC of BASELINE configs 1-2 filters at u = 3, p = 1 after the exact state
// reduction) give a warp nothing to do -- in the warp-per-draw kernel a 3 x 3 problem costs the same ~250 warp-instructions per
// step as a 10 x 10 one (the 8 x 8 x 4 tensor tiles are 95 % padding, every lane repeats the scalar p x p algebra), 0.05 of the fp64
// roofline.  Here every matrix of a draw lives in the registers of ONE thread (U, p compile-time, everything unrolled, no
// shared-memory round trips, no synchronisation): ~80 fp64 operations per step at u = 3, all of them useful.
#pragma once
#include "kalman.cuh"

namespace gecon {

// v[idx] for a run-time idx without indexing a register array (which would send it to local memory)
template <int U>
__device__ __forceinline__ double reg_pick(const double (&v)[U], int idx) {
    double r = v[0];
#pragma unroll
    for (int j = 1; j < U; ++j) r = (idx == j) ? v[j] : r;
    return r;
}

constexpr int KT_THREADS = 128;

template <int U, int PT>
__global__ void __launch_bounds__(KT_THREADS) kalman_ll_thread_kernel(const gecon_kalman_args p) {
    extern __shared__ __align__(16) double sm_kt[];
    const int tid = threadIdx.x, Tobs = p.Tobs, k = p.k;
    double* s_Y = sm_kt;
    int* s_wb = reinterpret_cast<int*>(sm_kt + (((size_t)Tobs * PT + 1) & ~(size_t)1));
    // ---- stage the observations once per CTA: masks first, then missing entries -> 0
    for (int i = tid; i < Tobs * PT; i += KT_THREADS) s_Y[i] = p.Y[i];
    __syncthreads();
    for (int t = tid; t < Tobs; t += KT_THREADS) {
        int bits = 0;
#pragma unroll
        for (int a = 0; a < PT; ++a) {
            const double yv = s_Y[t * PT + a];
            if (!(yv != yv || yv == p.missing_fill)) bits |= 1 << a;
        }
        s_wb[t] = bits;
    }
    __syncthreads();
    for (int i = tid; i < Tobs * PT; i += KT_THREADS) {
        const int t = i / PT, a = i - t * PT;
        if (!((s_wb[t] >> a) & 1)) s_Y[i] = 0.0;
    }
    __syncthreads();
    const long long draw = (long long)blockIdx.x * KT_THREADS + tid;
    if (draw >= p.N) return;

    int status = p.status_in ? (p.status_in[draw] & ~GECON_ST_BK_CERTIFIED) : 0;
    if (status & p.gate_mask) {
        p.ll[draw] = -INFINITY;
        p.status[draw] = status | GECON_ST_SKIPPED;
        if (p.ll_t) {
            for (int t = 0; t < Tobs; ++t) p.ll_t[(size_t)draw * Tobs + t] = -INFINITY;
        }
        return;
    }
    const double LOG2PI = 1.8378770664093453;
    const double ll_const = (p.mvn_const_mode == 0) ? PT * LOG2PI : LOG2PI;
    const double jitter = p.jitter;
    const bool keep_d = (p.mask_intercept == 0);
    const int lyap_cap = p.lyap_max_iter > 0 ? p.lyap_max_iter : 64;
    int obs_r[PT];
    double hv[PT], dv0[PT];
#pragma unroll
    for (int a = 0; a < PT; ++a) {
        obs_r[a] = p.obs_idx[a];
        const double h = (p.hdiag && (p.h_count <= 0 || a < p.h_count)) ? p.hdiag[(size_t)draw * p.h_stride + a] : 0.0;
        hv[a] = p.sigma_inputs ? h * h : h;
        dv0[a] = p.d ? p.d[(size_t)draw * p.d_stride + a] : 0.0;
    }
    // ---- T, C0 = R Q R' (mirrored pairs: exactly symmetric), c0 = C0 + jitter T T'
    double T[U][U], P[U][U], c0[U][U];
    {
        const double* gT = p.T + (size_t)draw * U * U;
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) {
                T[i][j] = gT[i * U + j];
                P[i][j] = 0.0;
            }
        const double* gR = p.R + (size_t)draw * U * k;
        for (int c = 0; c < k; ++c) {
            const double qv = p.qdiag[(size_t)draw * p.q_stride + c];
            const double q = p.sigma_inputs ? qv * qv : qv;
            double r[U];
#pragma unroll
            for (int i = 0; i < U; ++i) r[i] = gR[i * k + c];
#pragma unroll
            for (int i = 0; i < U; ++i)
#pragma unroll
                for (int j = i; j < U; ++j) P[i][j] = fma(r[i] * q, r[j], P[i][j]);
        }
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = i; j < U; ++j) {
                double tt = 0.0;
#pragma unroll
                for (int l = 0; l < U; ++l) tt = fma(T[i][l], T[j][l], tt);
                c0[i][j] = c0[j][i] = fma(jitter, tt, P[i][j]);
                P[j][i] = P[i][j];
            }
    }
    // ---- P0
    if (p.P0) {
        const double* g0 = p.P0 + (size_t)draw * U * U;
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) P[i][j] = g0[i * U + j];
    } else {
        // Smith doubling: P <- P + A_j P A_j', A_{j+1} = A_j^2, until max|increment| <= 1e-16 max|P|
        double A[U][U];
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) A[i][j] = T[i][j];
        bool done = false;
        for (int it = 0; it < lyap_cap; ++it) {
            double W[U][U], A2[U][U];
            double dmax = 0.0, pmax = 0.0;
#pragma unroll
            for (int i = 0; i < U; ++i)
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    double w = 0.0, a2 = 0.0;
#pragma unroll
                    for (int l = 0; l < U; ++l) {
                        w = fma(A[i][l], P[l][j], w);
                        a2 = fma(A[i][l], A[l][j], a2);
                    }
                    W[i][j] = w;
                    A2[i][j] = a2;
                }
#pragma unroll
            for (int i = 0; i < U; ++i)
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    double d = 0.0;
#pragma unroll
                    for (int l = 0; l < U; ++l) d = fma(W[i][l], A[j][l], d);
                    P[i][j] += d;
                    const double ad = fabs(d), ap = fabs(P[i][j]);
                    if (ad > dmax || ad != ad) dmax = ad;
                    if (ap > pmax || ap != ap) pmax = ap;
                }
            if (dmax != dmax || pmax != pmax) break;
            if (dmax <= 1e-16 * pmax) {
                done = true;
                break;
            }
#pragma unroll
            for (int i = 0; i < U; ++i)
#pragma unroll
                for (int j = 0; j < U; ++j) A[i][j] = A2[i][j];
        }
        if (!done) status |= GECON_ST_LYAP;
    }

    double am[U];
#pragma unroll
    for (int i = 0; i < U; ++i) am[i] = 0.0;
    double ll_acc = 0.0, quad_acc = 0.0, detprod = 1.0;
    long long det_exp = 0;
    int n_ll_steps = 0;
    bool notpd = false;
    for (int t = 0; t < Tobs; ++t) {
        const int wb = s_wb[t];
        // ---- P Z' (columns of P), innovation, F (lower triangle, from the rows obs_r[a] of P)
        double pz[U][PT], v[PT], f[PT][PT], prow[PT][U];
#pragma unroll
        for (int a = 0; a < PT; ++a) {
            const bool ob = (wb >> a) & 1;
#pragma unroll
            for (int j = 0; j < U; ++j) {
                double s = P[0][j];
#pragma unroll
                for (int i = 1; i < U; ++i) s = (obs_r[a] == i) ? P[i][j] : s;  // row obs_r[a] of P (P is exactly symmetric here)
                prow[a][j] = s;
                pz[j][a] = ob ? s : 0.0;
            }
            const double za = reg_pick<U>(am, obs_r[a]);
            v[a] = (s_Y[t * PT + a] - ((ob || keep_d) ? dv0[a] : 0.0)) - (ob ? za : 0.0);
        }
#pragma unroll
        for (int a = 0; a < PT; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                double x = reg_pick<U>(prow[a], obs_r[b]);
                x = (((wb >> a) & 1) && ((wb >> b) & 1)) ? x : 0.0;
                if (a == b) x += (((wb >> a) & 1) ? hv[a] : 0.0) + jitter;
                f[a][b] = x;
            }
        // L D L' of F, log det F = log prod d
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
                p.ll_t[(size_t)draw * Tobs + t] = llt;
            } else if (wb != 0) {
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
        // ---- K = P Z' F^-1 (row by row), a+ = a + K v, N = -(P Z' + jitter K), P+ = P + K N' (upper triangle, mirrored)
        double Kk[U][PT], Nn[U][PT], af[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
            double x[PT];
#pragma unroll
            for (int a = 0; a < PT; ++a) x[a] = pz[i][a];
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
            double s = am[i];
#pragma unroll
            for (int a = 0; a < PT; ++a) {
                s = fma(x[a], v[a], s);
                Kk[i][a] = x[a];
                Nn[i][a] = -fma(jitter, x[a], pz[i][a]);
            }
            af[i] = s;
        }
        double Pp[U][U];
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = i; j < U; ++j) {
                double s = P[i][j];
#pragma unroll
                for (int a = 0; a < PT; ++a) s = fma(Kk[i][a], Nn[j][a], s);
                Pp[i][j] = Pp[j][i] = s;
            }
        // ---- predict: W = T P+, P = c0 + W T' (upper triangle, mirrored), a = T a+
        double W[U][U];
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < U; ++j) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < U; ++l) s = fma(T[i][l], Pp[l][j], s);
                W[i][j] = s;
            }
#pragma unroll
        for (int i = 0; i < U; ++i) {
#pragma unroll
            for (int j = i; j < U; ++j) {
                double s = c0[i][j];
#pragma unroll
                for (int l = 0; l < U; ++l) s = fma(W[i][l], T[j][l], s);
                P[i][j] = P[j][i] = s;
            }
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < U; ++l) s = fma(T[i][l], af[l], s);
            am[i] = s;
        }
    }
    if (!p.ll_t) ll_acc = -0.5 * (n_ll_steps * ll_const + (log(detprod) + (double)det_exp * 0.6931471805599453) + quad_acc);
    if (notpd) status |= GECON_ST_NOT_PD;
    if (!(fabs(ll_acc) <= 1.7e308)) status |= GECON_ST_LL_NONFINITE;
    p.ll[draw] = ll_acc;
    p.status[draw] = status;
}

template <int U, int PT>
static int launch_kalman_thread(const gecon_kalman_args& a, cudaStream_t st, int* info) {
    const size_t smem = sizeof(double) * ((((size_t)a.Tobs * PT + 1) & ~(size_t)1)) + sizeof(int) * ((size_t)a.Tobs + 4);
    if (smem > 200 * 1024) {
        set_last_error("observation matrix does not fit in shared memory (%zu bytes needed)", smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    GECON_CUDA(cudaFuncSetAttribute(kalman_ll_thread_kernel<U, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (info) {
        int per_sm = 0;
        GECON_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kalman_ll_thread_kernel<U, PT>, KT_THREADS, smem));
        info[0] = per_sm;
        info[1] = (int)smem;
        info[2] = KT_THREADS;
        return 0;
    }
    const long long grid = (a.N + KT_THREADS - 1) / KT_THREADS;
    kalman_ll_thread_kernel<U, PT><<<(unsigned)grid, KT_THREADS, smem, st>>>(a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gecon
