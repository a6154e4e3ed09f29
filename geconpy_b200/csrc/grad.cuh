// Gradient path: reverse-mode kernels for the Kalman log-likelihood and for the policy function.
//
// What the reference differentiates with pytensor for NUTS (gEconpy/model/statespace.py:812-820,1151-1157 through the
// solver Ops' pullbacks, gEconpy/solvers/cycle_reduction.py:212-213 -> gEconpy/solvers/shared.py:12-71) is done here as
// two hand-written sweeps per draw, one CTA per draw:
//
//   kalman_grad_draw   forward filter storing the predicted moments (a_t, P_t) of every step in a per-CTA trajectory
//                      buffer, then the reverse sweep:  ll,  dll/dT, dll/dR, dll/dq, dll/dh, dll/dd
//                      (adjoint of P0 = dlyap(T, R Q R') by a second Smith doubling, S = T' S T + P0_bar)
//   policy_adjoint_draw  R = -(C T + B)^-1 D and A + B T + C T T = 0 in reverse:  the multipliers of
//                      o1_policy_function_adjoints solve W' S + C' S T' = -T_bar (W = C T + B); instead of the
//                      reference's n^2 x n^2 Kronecker system this is the Stein equation S = Q + G S T',
//                      G = -W^-T C', Q = -W^-T T_bar, solved by doubling (rho(G) rho(T) < 1 under Blanchard-Kahn)
//
// Every phase is a data-parallel loop over output elements (GFOR) separated by barriers (GSYNC), and nothing else:
// the same source compiles as plain C++ with -DGECON_HOST_CHECK (one "thread", barriers vanish), which is how the
// CPU test-suite checks the arithmetic against oracle/adjoints.py without a GPU (tests/test_adjoints_cpu.py).
// Plain DFMA inner products on odd-leading-dimension shared tiles; these kernels are a first correct gradient path,
// not yet tuned like the forward filter.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef GECON_HOST_CHECK
#define GHD inline
#define GHH inline
#define G_TID 0
#define G_NT 1
#define GSYNC() \
    do {        \
    } while (0)
#else
#define GHD __device__ __forceinline__
#define GHH __host__ __device__ __forceinline__
#define G_TID ((int)threadIdx.x)
#if defined(GECON_GRAD_CN) && GECON_GRAD_CN <= 12
// per-configuration build with one warp per draw (grad_spec.cu): the CTA IS a warp -- compile-time thread count (every GFOR over a
// compile-time extent resolves to straight-line predicated code: 8.1 k -> 4.1 k SASS instructions at n = 10), warp-level barrier
#define G_NT 32
#define GSYNC() __syncwarp()
#elif defined(GECON_GRAD_CN)
#define G_NT (GECON_GRAD_CN <= 32 ? 128 : 256)  // the thread count grad_spec.cu launches with
#define GSYNC() __syncthreads()
#else
#define G_NT ((int)blockDim.x)
#define GSYNC() __syncthreads()
#endif
#endif
#define GFOR(i, count) for (int i = G_TID; i < (count); i += G_NT)

namespace gecon_grad {

constexpr int PMAXG = 8;
constexpr int ST_SINGULAR = 0x004, ST_LYAP = 0x040, ST_NOT_PD = 0x080, ST_LL_NONFINITE = 0x100;

GHH int ldim(int n) { return n | 1; }  // odd leading dimension: column walks hit distinct 8-byte banks

// Register-blocked product: every work item is one row i and four consecutive columns j0..j0+3 of the output,
//   epi(i, j, sum_k a(i, k) * b(k, j)),   i < m, j < n, k < kk,
// so one a-load and four b-loads feed four FMA (1.25 shared loads per FMA instead of 2) and the index decode is paid once
// per four outputs.  a, b, epi are inlined lambdas; column indices past n are clamped (their results are dropped).
// tab (optional): work item -> (i << 16) | j0, precomputed once so that the sweeps do no integer division.
template <class FA, class FB, class FE>
GHD void gemm4(int m, int n, int kk, FA a, FB b, FE epi, const int* tab = nullptr) {
    const int nb = (n + 3) >> 2;
    GFOR(w, m * nb) {
        const int i = tab ? (tab[w] >> 16) : w / nb, j0 = tab ? (tab[w] & 0xffff) : (w - (w / nb) * nb) << 2;
        const int j1 = (j0 + 1 < n) ? j0 + 1 : n - 1, j2 = (j0 + 2 < n) ? j0 + 2 : n - 1, j3 = (j0 + 3 < n) ? j0 + 3 : n - 1;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
        for (int k = 0; k < kk; ++k) {
            const double av = a(i, k);
            s0 = fma(av, b(k, j0), s0);
            s1 = fma(av, b(k, j1), s1);
            s2 = fma(av, b(k, j2), s2);
            s3 = fma(av, b(k, j3), s3);
        }
        epi(i, j0, s0);
        if (j0 + 1 < n) epi(i, j0 + 1, s1);
        if (j0 + 2 < n) epi(i, j0 + 2, s2);
        if (j0 + 3 < n) epi(i, j0 + 3, s3);
    }
}

// Square product (m = n = kk = the filter dimension) of the sweeps.  Generic builds and the CPU check: the scalar gemm4 above.
// Per-configuration builds (grad_spec.cu, compile-time n): the fp64 tensor path -- every warp accumulates whole 8-row strips of the
// output with mma.sync.m8n8k4, all tiles of a strip (all tiles of the matrix when the CTA is one warp) at once, so the only dependent
// chain is the k loop; operands and epilogue are the same inlined lambdas, guarded at the edges (the shared-memory tiles carry no
// padding).  At n = 10: 12 DMMA + 12 predicated loads + 8 epilogue elements per lane instead of ~210 load / FMA / index instructions.
// (With run-time dimensions the same idea was slower than gemm4 -- guards and loop control ate the gain)
// KP = true: the contraction runs over the p observables instead of the n states (the rank-p terms K N', PZ_bar Zm).
template <bool KP = false, class FA, class FB, class FE>
GHD void gemm_sq(int n_rt, int kk_rt, FA a, FB b, FE epi, const int* tab) {
#if defined(GECON_GRAD_CN) && !defined(GECON_HOST_CHECK)
    (void)n_rt;
    (void)kk_rt;
    (void)tab;
    constexpr int n = GECON_GRAD_CN, kdim = KP ? GECON_GRAD_CP : GECON_GRAD_CN, NSQ = (n + 7) / 8, NK = (kdim + 3) / 4, NWARP = G_NT / 32;
    constexpr int SPW = NWARP == 1 ? NSQ : 1;  // strips per pass of a warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    for (int s0 = (NWARP == 1 ? 0 : warp); s0 < NSQ; s0 += (NWARP == 1 ? NSQ : NWARP)) {
        double acc[SPW][NSQ][2];
#pragma unroll
        for (int s = 0; s < SPW; ++s)
#pragma unroll
            for (int ct = 0; ct < NSQ; ++ct) acc[s][ct][0] = acc[s][ct][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            const int kq = 4 * ks + q;
            double av[SPW], bv[NSQ];
#pragma unroll
            for (int s = 0; s < SPW; ++s) {
                const int r = 8 * (s0 + s) + g;
                av[s] = (r < n && kq < kdim) ? a(r, kq) : 0.0;
            }
#pragma unroll
            for (int ct = 0; ct < NSQ; ++ct) {
                const int c = 8 * ct + g;
                bv[ct] = (c < n && kq < kdim) ? b(kq, c) : 0.0;
            }
#pragma unroll
            for (int s = 0; s < SPW; ++s)
#pragma unroll
                for (int ct = 0; ct < NSQ; ++ct)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc[s][ct][0]), "+d"(acc[s][ct][1])
                                 : "d"(av[s]), "d"(bv[ct]));
        }
#pragma unroll
        for (int s = 0; s < SPW; ++s) {
            const int r = 8 * (s0 + s) + g;
#pragma unroll
            for (int ct = 0; ct < NSQ; ++ct) {
                const int oc = 8 * ct + 2 * q;
                if (r < n && oc < n) epi(r, oc, acc[s][ct][0]);
                if (r < n && oc + 1 < n) epi(r, oc + 1, acc[s][ct][1]);
            }
        }
    }
#else
    gemm4(n_rt, n_rt, kk_rt, a, b, epi, tab);
#endif
}

// C (m x n, ldc) = alpha * op(A) * op(B) + beta * C;  op(A) is m x kk, op(B) is kk x n.  C must not alias A or B.
template <bool TA, bool TB>
GHD void mm(double* C, int ldc, const double* A, int lda, const double* B, int ldb, int m, int n, int kk, double alpha, double beta) {
    gemm4(
        m, n, kk, [&](int i, int k) { return TA ? A[k * lda + i] : A[i * lda + k]; },
        [&](int k, int j) { return TB ? B[j * ldb + k] : B[k * ldb + j]; },
        [&](int i, int j, double v) { C[i * ldc + j] = (beta != 0.0) ? fma(alpha, v, beta * C[i * ldc + j]) : alpha * v; });
}

// max |x| over m x n (NaN-propagating: a NaN makes the result NaN); s_red: G_NT doubles.  Contains barriers.
GHD double absmax(const double* X, int ldx, int m, int n, double* s_red) {
    double mx = 0.0;
    GFOR(idx, m * n) {
        const int i = idx / n, j = idx - i * n;
        const double a = fabs(X[i * ldx + j]);
        if (a > mx || a != a) mx = a;
    }
    s_red[G_TID] = mx;
    GSYNC();
    double r = 0.0;
    for (int t = 0; t < G_NT; ++t) {
        const double a = s_red[t];
        if (a > r || a != a) r = a;
    }
    GSYNC();
    return r;
}

// In-place inverse of the SPD p x p matrix F (row-major, ld = PMAXG) by Gauss-Jordan without pivoting, executed by one
// thread; returns det F through *logdet and false if a pivot is not positive.
GHD bool spd_inverse_small(double* F, int p, int ps, double* logdet) {
    double det = 1.0;  // p <= 8 pivots of an innovation covariance: their product does not leave the double range
    bool ok = true;
    for (int c = 0; c < p; ++c) {
        const double piv = F[c * ps + c];
        ok = ok && (piv > 0.0);
        det *= piv;
        const double inv = 1.0 / piv;
        for (int j = 0; j < p; ++j) F[c * ps + j] *= inv;
        F[c * ps + c] = inv;
        for (int i = 0; i < p; ++i) {
            if (i == c) continue;
            const double f = F[i * ps + c];
            F[i * ps + c] = 0.0;
            for (int j = 0; j < p; ++j) F[i * ps + j] = fma(-f, F[c * ps + j], F[i * ps + j]);
        }
    }
    if (logdet) *logdet = det;  // the DETERMINANT: the caller multiplies them up and takes one logarithm per draw
    return ok;
}

struct KalmanGradArgs {
    const double* T;      // [N][n][n]
    const double* R;      // [N][n][k]
    const double* qdiag;  // [N][k] or [k]
    long long q_stride;
    const double* hdiag;  // [N][p], [p] or NULL
    long long h_stride;
    const double* Z;        // [p][n] shared (z_stride = 0) or [N][p][n] (z_stride = p n), or NULL
    long long z_stride;
    const int32_t* obs_idx;  // [p] or NULL
    const double* d;        // [N][p], [p] or NULL
    long long d_stride;
    const double* Y;  // [Tobs][p]
    long long N;
    int n, k, p, Tobs;
    double jitter, missing_fill;
    int mvn_const_mode, lyap_max_iter;
    const int32_t* status_in;
    int gate_mask, sigma_inputs;
    int mask_intercept;  // 1: the intercept is masked at missing entries like Z and H
    double* ll;       // [N]
    int32_t* status;  // [N]
    double* T_bar;    // [N][n][n]
    double* R_bar;    // [N][n][k]
    double* q_bar;    // [N][k]   (w.r.t. sigma when sigma_inputs)
    double* h_bar;    // [N][p]
    double* d_bar;    // [N][p]
    double* Z_bar;    // [N][p][n] or NULL: dll/dZ (dense design matrices only)
    const double* qfull;  // [N][k][k] (qfull_stride = k k) or [k][k] (0), or NULL: full shock covariance (then qdiag / q_bar are unused)
    long long qfull_stride;
    double* qfull_bar;    // [N][k][k]: dll/dQ
    double* traj;     // workspace [n_cta][Tobs][n n + n]
    double* c0bar_ws;  // workspace [n_cta][2][n n]: adjoint of R Q R', and R Q R' itself
};

// doubles per step of the trajectory buffer: predicted P (n n) and a (n), then F^-1 (p p), v (p) and K (n p) of the update
GHH size_t kalman_grad_traj_stride(int n, int p) { return (size_t)n * n + n + (size_t)p * p + p + (size_t)n * p; }

// doubles of shared memory needed by kalman_grad_draw
GHH size_t kalman_grad_smem_doubles(int n, int k, int p, int nt) {
    const int ld = ldim(n);
    return (size_t)8 * n * ld + (size_t)5 * n * p + (size_t)2 * p * n + 5 * (size_t)n + 2 * (size_t)p * p + 9 * (size_t)p + 6 + (size_t)n * (k > 0 ? k : 1) +
           k + (size_t)k * k + nt + 8 + ((size_t)n * n + (size_t)n * p + (size_t)n * ((n + 3) / 4) + 1) / 2;
}

// One draw.  sm: shared memory (kalman_grad_smem_doubles), cta: index of this CTA's workspace slot.
//
// Forward step (what is differentiated; same function of the parameters as the Joseph form of oracle/statespace.py):
//   PZ = P Zm',  v = ym - dm - Zm a,  F = Zm PZ + Hm + j I,  K = PZ F^-1,  e = F^-1 v,  af = a + K v,
//   N = PZ + j K,  Pf = P - K N' + j I     [= (I - K Zm) P (I - K Zm)' + K Hm K' + j I: K PZ' is symmetric and K (F - j I) = PZ - j K]
//   a' = T af,  P' = T Pf T' + C0
// so the update and its adjoint are O(n^2 p); the only n^3 products are T Pf, (T Pf) T' forward and, in reverse,
//   T_bar += (P'_bar + P'_bar') (T Pf) + a'_bar af',   Pf_bar = T' (P'_bar T),   C0_bar += P'_bar,   af_bar = T' a'_bar
//   N_bar = -Pf_bar' K,  K_bar = -Pf_bar N + j N_bar + af_bar v',  v_bar = -e + K' af_bar,  F_bar = -(F^-1 - e e') / 2
//   G = K_bar F^-1,  PZ_bar = N_bar + G + Zm' F_bar',  F_bar' = sym(F_bar - K' G),  h_bar += w diag(F_bar'),
//   P_bar = Pf_bar + PZ_bar Zm,  a_bar = af_bar - Zm' v_bar,  d_bar -= dmask v_bar,
//   Zm_bar = F_bar' PZ' + PZ_bar' P - v_bar a'
// (13 n^3 products per step with the expanded Joseph form and its adjoint until round 2; 6 now).
// F is symmetrised before it is inverted and F_bar' is symmetrised before it is used: both are part of the graph.  Without them
// the rank-p update is the same function of the parameters but NOT a stable recursion -- an antisymmetric perturbation of P is
// neither damped by the update (the Joseph form damps it by L . L') nor kept out of F, and the reverse sweep amplifies
// rounding-level antisymmetry of P_bar by the same mechanism (measured on the medium NK model with error variances 1e-6:
// gradient wrong in the first digit after 50 steps; with the two symmetrisations it agrees with the Joseph-form adjoint to 1e-12).
GHD void kalman_grad_draw(const KalmanGradArgs& g, long long draw, int cta, double* sm) {
#ifdef GECON_GRAD_CN  // per-configuration build (grad_spec.cu): the dimensions are compile-time, every loop below unrolls
    constexpr int n = GECON_GRAD_CN, k = GECON_GRAD_CK, p = GECON_GRAD_CP, ld = GECON_GRAD_CN | 1;
    const int Tobs = g.Tobs;
#else
    const int n = g.n, k = g.k, p = g.p, Tobs = g.Tobs, ld = ldim(n);
#endif
    const int tile = n * ld;
    double* Tm = sm;
    double* P = Tm + tile;
    double* Pf = P + tile;
    double* Pb = Pf + tile;   // adjoint of the predicted covariance
    double* Pfb = Pb + tile;  // adjoint of the filtered covariance (and the A_k of the two doublings)
    double* W1 = Pfb + tile;
    double* W2 = W1 + tile;
    double* Tb = W2 + tile;   // accumulated adjoint of T
    const int ps = p;             // row stride of the n x p panels and of the p x p matrices
    double* PZ = Tb + tile;       // [n][p]
    double* K = PZ + n * ps;
    double* Kb = K + n * ps;
    double* PZb = Kb + n * ps;
    double* G1 = PZb + n * ps;    // K_bar F^-1
    double* Zs = G1 + n * ps;     // [p][n]
    double* Zb = Zs + p * n;      // [p][n] adjoint of the design matrix
    double* a = Zb + p * n;
    double* af = a + n;
    double* ab = af + n;
    double* afb = ab + n;
    double* an = afb + n;
    double* F = an + n;           // [p][p]  F, then F^-1
    double* Fb = F + p * p;       // adjoint of F
    double* v = Fb + p * p;
    double* e = v + p;
    double* vb = e + p;
    double* w = vb + p;
    double* ym = w + p;
    double* hv = ym + p;
    double* dv = hv + p;
    double* hb = dv + p;
    double* db = hb + p;
    double* sc = db + p;  // [0] det F, [1] ok flag, [2] ll without the log-determinants, [3] all-missing flag, [4], [5] their running product
    double* Rs = sc + 6;  // [n][k]   (sc[4], sc[5]: mantissa and exponent of the running product of the determinants)
    double* qs = Rs + n * (k > 0 ? k : 1);
    double* Qs = qs + k;  // [k][k] full shock covariance (when g.qfull)
    double* s_red = Qs + k * k;
    int* dec = reinterpret_cast<int*>(s_red + G_NT + 8);  // [n n] element index -> (i << 16) | j: no integer division inside the sweeps
    int* decp = dec + n * n;                              // [n p] panel index -> (i << 16) | c
    int* decg = decp + n * p;                             // [n ceil(n / 4)] work item of an n x n gemm4 -> (i << 16) | j0

    const double LOG2PI = 1.8378770664093453;
    const double ll_const = (g.mvn_const_mode == 0) ? p * LOG2PI : LOG2PI;
    const double jit = g.jitter;
    int status = g.status_in ? (g.status_in[draw] & ~0x800) : 0;
    double* gTb = g.T_bar + (size_t)draw * n * n;
    double* gRb = g.R_bar + (size_t)draw * n * k;
    if (status & g.gate_mask) {
        GFOR(i, n * n) gTb[i] = 0.0;
        GFOR(i, n * k) gRb[i] = 0.0;
        if (g.qfull) GFOR(i, k * k) g.qfull_bar[(size_t)draw * k * k + i] = 0.0;
        else GFOR(i, k) g.q_bar[(size_t)draw * k + i] = 0.0;
        GFOR(i, p) {
            if (g.h_bar) g.h_bar[(size_t)draw * p + i] = 0.0;
            if (g.d_bar) g.d_bar[(size_t)draw * p + i] = 0.0;
        }
        if (g.Z_bar) GFOR(i, p * n) g.Z_bar[(size_t)draw * p * n + i] = 0.0;
        if (G_TID == 0) {
            g.ll[draw] = -INFINITY;
            g.status[draw] = status | 0x400;
        }
        return;
    }
    const size_t TS = kalman_grad_traj_stride(n, p);
    double* traj = g.traj + (size_t)cta * Tobs * TS;
    double* gC0b = g.c0bar_ws + (size_t)cta * 2 * n * n;
    double* C0 = gC0b + n * n;  // R Q R' (read once per forward step): global, L2-resident, to keep the shared-memory footprint down
    const double* gT = g.T + (size_t)draw * n * n;
    const double* gR = g.R + (size_t)draw * n * k;

    // ---- load
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        Tm[i * ld + j] = gT[idx];
        Tb[i * ld + j] = 0.0;
        gC0b[idx] = 0.0;
    }
    GFOR(idx, n * k) Rs[idx] = gR[idx];
    if (g.qfull) {
        GFOR(idx, k * k) Qs[idx] = g.qfull[(size_t)draw * (size_t)g.qfull_stride + idx];
    } else {
        GFOR(c, k) {
            const double qv = g.qdiag[(size_t)draw * g.q_stride + c];
            qs[c] = g.sigma_inputs ? qv * qv : qv;
        }
    }
    GFOR(i, p) {
        const double h = g.hdiag ? g.hdiag[(size_t)draw * g.h_stride + i] : 0.0;
        hv[i] = g.sigma_inputs ? h * h : h;
        dv[i] = g.d ? g.d[(size_t)draw * g.d_stride + i] : 0.0;
        hb[i] = 0.0;
        db[i] = 0.0;
    }
    GFOR(idx, p * n) {
        const int i = idx / n, j = idx - i * n;
        Zs[idx] = g.Z ? g.Z[(size_t)draw * (size_t)g.z_stride + idx] : ((g.obs_idx[i] == j) ? 1.0 : 0.0);
        Zb[idx] = 0.0;
    }
    GFOR(i, n) {
        a[i] = 0.0;
        ab[i] = 0.0;
    }
    GFOR(idx, n * n) dec[idx] = ((idx / n) << 16) | (idx % n);
    GFOR(idx, n * p) decp[idx] = ((idx / p) << 16) | (idx % p);
    GFOR(idx, n * ((n + 3) >> 2)) decg[idx] = ((idx / ((n + 3) >> 2)) << 16) | ((idx % ((n + 3) >> 2)) << 2);
    GSYNC();
    // C0 = R Q R'
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        double s = 0.0;
        if (g.qfull) {  // R Q R' exactly as written (every entry of Q is an independent input of the derivative)
            for (int c = 0; c < k; ++c) {
                double rq = 0.0;
                for (int c2 = 0; c2 < k; ++c2) rq = fma(Qs[c * k + c2], Rs[j * k + c2], rq);
                s = fma(Rs[i * k + c], rq, s);
            }
        } else {
            for (int c = 0; c < k; ++c) s = fma(Rs[lo * k + c] * qs[c], Rs[hi * k + c], s);
        }
        C0[i * n + j] = s;
        P[i * ld + j] = s;
        Pfb[i * ld + j] = Tm[i * ld + j];  // A_0 = T
    }
    GSYNC();
    // ---- P0 by Smith doubling: P <- P + A P A', A <- A^2
    {
        const int cap = g.lyap_max_iter > 0 ? g.lyap_max_iter : 64;
        bool done = false;
        for (int it = 0; it < cap && !done; ++it) {
            mm<false, false>(W1, ld, Pfb, ld, P, ld, n, n, n, 1.0, 0.0);
            GSYNC();
            mm<false, true>(W2, ld, W1, ld, Pfb, ld, n, n, n, 1.0, 0.0);
            mm<false, false>(Pf, ld, Pfb, ld, Pfb, ld, n, n, n, 1.0, 0.0);
            GSYNC();
            GFOR(idx, n * n) {
                const int i = idx / n, j = idx - i * n;
                P[i * ld + j] += W2[i * ld + j];
                Pfb[i * ld + j] = Pf[i * ld + j];
            }
            GSYNC();
            const double dmax = absmax(W2, ld, n, n, s_red), pmax = absmax(P, ld, n, n, s_red);
            if (dmax != dmax || pmax != pmax) break;
            if (dmax <= 1e-16 * pmax) done = true;
        }
        if (!done) status |= ST_LYAP;
    }
    if (G_TID == 0) {
        sc[2] = 0.0;
        sc[1] = 1.0;
        sc[4] = 1.0;
        sc[5] = 0.0;
    }
    GSYNC();

    // P Zm' (n x p) and the missing-value masks of step t
    auto masks = [&](int t) {
        GFOR(i, p) {
            const double yv = g.Y[(size_t)t * p + i];
            const bool miss = (yv != yv) || (yv == g.missing_fill);
            w[i] = miss ? 0.0 : 1.0;
            ym[i] = miss ? 0.0 : yv;
        }
    };
    auto pz_panel = [&]() {
        GFOR(idx, n * p) {
            const int i = decp[idx] >> 16, c = decp[idx] & 0xffff;
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(P[i * ld + j], Zs[c * n + j], s);
            PZ[i * ps + c] = w[c] * s;
        }
    };
    // Pf = P - K (PZ + j K)' + j I   (N = PZ + j K is never stored)
    auto filtered_cov = [&]() {
        gemm_sq<true>(n, p, [&](int i, int c) { return -K[i * ps + c]; }, [&](int c, int j) { return fma(jit, K[j * ps + c], PZ[j * ps + c]); },
                      [&](int i, int j, double v_) { Pf[i * ld + j] = v_ + P[i * ld + j] + ((i == j) ? jit : 0.0); }, decg);
    };

    // ---- forward sweep: store the predicted moments and the update's F^-1, v, K; filter; predict
    for (int t = 0; t < Tobs; ++t) {
        double* tr = traj + (size_t)t * TS;
        GFOR(idx, n * n) tr[idx] = P[(dec[idx] >> 16) * ld + (dec[idx] & 0xffff)];
        GFOR(i, n) tr[n * n + i] = a[i];
        masks(t);
        GSYNC();
        pz_panel();
        GFOR(c, p) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(Zs[c * n + j], a[j], s);
            v[c] = ym[c] - ((g.mask_intercept ? w[c] : 1.0) * dv[c] + w[c] * s);
        }
        GSYNC();
        GFOR(idx, p * p) {
            const int c = idx / p, b = idx - c * p;
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(Zs[c * n + j], PZ[j * ps + b], s);
            s *= w[c];
            if (c == b) s += w[c] * hv[c] + jit;
            F[c * ps + b] = s;
        }
        GSYNC();
        if (G_TID == 0) {
            // symmetrise (the two triangles differ by rounding), invert
            for (int c = 0; c < p; ++c)
                for (int b = 0; b < c; ++b) {
                    const double s = 0.5 * (F[c * ps + b] + F[b * ps + c]);
                    F[c * ps + b] = F[b * ps + c] = s;
                }
            double logdet = 0.0;
            const bool ok = spd_inverse_small(F, p, ps, &logdet);
            if (!ok) sc[1] = 0.0;
            sc[0] = logdet;
            double allmiss = 1.0;
            for (int c = 0; c < p; ++c)
                if (w[c] != 0.0) allmiss = 0.0;
            sc[3] = allmiss;
        }
        GSYNC();
        double* trF = tr + n * n + n;
        double* trv = trF + p * p;
        double* trK = trv + p;
        GFOR(idx, n * p) {
            const int i = decp[idx] >> 16, c = decp[idx] & 0xffff;
            double s = 0.0;
            for (int b = 0; b < p; ++b) s = fma(PZ[i * ps + b], F[b * ps + c], s);
            K[i * ps + c] = s;
            trK[idx] = s;
        }
        GFOR(c, p) {
            double s = 0.0;
            for (int b = 0; b < p; ++b) s = fma(F[c * ps + b], v[b], s);
            e[c] = s;
            trv[c] = v[c];
        }
        GFOR(idx, p * p) trF[idx] = F[idx];
        GSYNC();
        GFOR(i, n) {
            double s = a[i];
            for (int c = 0; c < p; ++c) s = fma(K[i * ps + c], v[c], s);
            af[i] = s;
        }
        filtered_cov();
        if (G_TID == 0 && sc[3] == 0.0) {
            double quad = 0.0;
            for (int c = 0; c < p; ++c) quad = fma(v[c], e[c], quad);
            int ex, ex2;
            sc[4] = frexp(sc[4] * frexp(sc[0], &ex), &ex2);  // running product of the determinants: mantissa in sc[4] ...
            sc[5] += (double)(ex + ex2);                     // ... exponent in sc[5]
            sc[2] += -0.5 * (ll_const + quad);
        }
        GSYNC();
        gemm_sq(n, n, [&](int i, int k_) { return Tm[i * ld + k_]; }, [&](int k_, int j) { return Pf[k_ * ld + j]; },
              [&](int i, int j, double v_) { W2[i * ld + j] = v_; }, decg);
        GFOR(i, n) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(Tm[i * ld + j], af[j], s);
            an[i] = s;
        }
        GSYNC();
        gemm_sq(n, n, [&](int i, int k_) { return W2[i * ld + k_]; }, [&](int k_, int j) { return Tm[j * ld + k_]; },
              [&](int i, int j, double v_) { P[i * ld + j] = v_ + C0[i * n + j]; }, decg);
        GFOR(i, n) a[i] = an[i];
        GSYNC();
    }
    const double ll = sc[2] - 0.5 * (log(sc[4]) + sc[5] * 0.6931471805599453);
    if (sc[1] == 0.0) status |= ST_NOT_PD;
    if (!(fabs(ll) <= 1.7e308)) status |= ST_LL_NONFINITE;

    // ---- reverse sweep
    GFOR(idx, n * ld) Pb[idx] = 0.0;
#if defined(GECON_GRAD_CN) && !defined(GECON_HOST_CHECK)
    constexpr int NPRE = (n * n + n + p * p + p + n * p + G_NT - 1) / G_NT;  // doubles of one trajectory record per thread
    double pre[NPRE];
    if (Tobs > 0) {
        const double* trn = traj + (size_t)(Tobs - 1) * TS;
#pragma unroll
        for (int m_ = 0; m_ < NPRE; ++m_) {
            const int idx = G_TID + m_ * G_NT;
            pre[m_] = (idx < (int)TS) ? trn[idx] : 0.0;
        }
    }
#endif
    GSYNC();
    for (int t = Tobs - 1; t >= 0; --t) {
        const double* tr = traj + (size_t)t * TS;
        const double* trF = tr + n * n + n;
        const double* trv = trF + p * p;
        const double* trK = trv + p;
        // replay of the update: P, a, F^-1, v, K come back from the trajectory; PZ, e, af, Pf are recomputed (O(n^2 p))
#if defined(GECON_GRAD_CN) && !defined(GECON_HOST_CHECK)
        // per-configuration build: the record of step t was fetched into registers one step ahead (its global-memory latency was 11 % of
        // the stall samples); commit it to the tiles, then start the loads of step t - 1
        (void)trF, (void)trv, (void)trK;
#pragma unroll
        for (int m_ = 0; m_ < NPRE; ++m_) {
            const int idx = G_TID + m_ * G_NT;
            const double val = pre[m_];
            if (idx < n * n) P[(dec[idx] >> 16) * ld + (dec[idx] & 0xffff)] = val;
            else if (idx < n * n + n) a[idx - n * n] = val;
            else if (idx < n * n + n + p * p) F[idx - n * n - n] = val;
            else if (idx < n * n + n + p * p + p) v[idx - n * n - n - p * p] = val;
            else if (idx < (int)TS) K[idx - n * n - n - p * p - p] = val;
        }
        if (t > 0) {
            const double* trn = tr - TS;
#pragma unroll
            for (int m_ = 0; m_ < NPRE; ++m_) {
                const int idx = G_TID + m_ * G_NT;
                pre[m_] = (idx < (int)TS) ? trn[idx] : 0.0;
            }
        }
#else
        GFOR(idx, n * n) P[(dec[idx] >> 16) * ld + (dec[idx] & 0xffff)] = tr[idx];
        GFOR(i, n) a[i] = tr[n * n + i];
        GFOR(idx, p * p) F[idx] = trF[idx];
        GFOR(c, p) v[c] = trv[c];
        GFOR(idx, n * p) K[idx] = trK[idx];
#endif
        masks(t);
        GSYNC();
        pz_panel();
        GFOR(c, p) {
            double s = 0.0;
            for (int b = 0; b < p; ++b) s = fma(F[c * ps + b], v[b], s);
            e[c] = s;
        }
        GFOR(i, n) {
            double s = a[i];
            for (int c = 0; c < p; ++c) s = fma(K[i * ps + c], v[c], s);
            af[i] = s;
        }
        if (G_TID == 0) {
            double allmiss = 1.0;
            for (int c = 0; c < p; ++c)
                if (w[c] != 0.0) allmiss = 0.0;
            sc[3] = allmiss;
        }
        GSYNC();
        filtered_cov();
        GSYNC();
        // predict in reverse:  P' = T Pf T' + C0,  a' = T af
        gemm_sq(n, n, [&](int i, int k_) { return Tm[i * ld + k_]; }, [&](int k_, int j) { return Pf[k_ * ld + j]; },
              [&](int i, int j, double v_) { W2[i * ld + j] = v_; }, decg);
        GSYNC();
        gemm_sq(n, n, [&](int i, int k_) { return Pb[i * ld + k_] + Pb[k_ * ld + i]; }, [&](int k_, int j) { return W2[k_ * ld + j]; },
              [&](int i, int j, double v_) {
                  Tb[i * ld + j] += v_ + ab[i] * af[j];
                  gC0b[i * n + j] += Pb[i * ld + j];
              }, decg);
        gemm_sq(n, n, [&](int i, int k_) { return Pb[i * ld + k_]; }, [&](int k_, int j) { return Tm[k_ * ld + j]; },
              [&](int i, int j, double v_) { W1[i * ld + j] = v_; }, decg);
        GFOR(i, n) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s = fma(Tm[j * ld + i], ab[j], s);
            afb[i] = s;
        }
        GSYNC();
        gemm_sq(n, n, [&](int i, int k_) { return Tm[k_ * ld + i]; }, [&](int k_, int j) { return W1[k_ * ld + j]; },
              [&](int i, int j, double v_) { Pfb[i * ld + j] = v_; }, decg);
        // log-likelihood term
        GFOR(idx, p * p) {
            const int c = idx / p, b = idx - c * p;
            Fb[c * ps + b] = (sc[3] == 0.0) ? -0.5 * (F[c * ps + b] - e[c] * e[b]) : 0.0;
        }
        GFOR(c, p) {
            double u = (sc[3] == 0.0) ? -e[c] : 0.0;
            for (int i = 0; i < n; ++i) u = fma(K[i * ps + c], afb[i], u);
            vb[c] = u;
        }
        GSYNC();
        // update in reverse (all O(n^2 p))
        GFOR(idx, n * p) {
            const int i = decp[idx] >> 16, c = decp[idx] & 0xffff;
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < n; ++j) {
                const double kj = K[j * ps + c];
                s1 = fma(Pfb[i * ld + j], fma(jit, kj, PZ[j * ps + c]), s1);
                s2 = fma(Pfb[j * ld + i], kj, s2);
            }
            PZb[i * ps + c] = -s2;                                  // N_bar
            Kb[i * ps + c] = fma(afb[i], v[c], -fma(jit, s2, s1));  // -Pf_bar N + j N_bar + af_bar v'
        }
        GSYNC();
        GFOR(idx, n * p) {  // G = K_bar F^-1
            const int i = decp[idx] >> 16, c = decp[idx] & 0xffff;
            double s = 0.0;
            for (int b = 0; b < p; ++b) s = fma(Kb[i * ps + b], F[b * ps + c], s);
            G1[i * ps + c] = s;
        }
        GSYNC();
        GFOR(idx, p * p) {  // F_bar -= K' G
            const int c = idx / p, b = idx - c * p;
            double s = 0.0;
            for (int i = 0; i < n; ++i) s = fma(K[i * ps + c], G1[i * ps + b], s);
            Fb[c * ps + b] -= s;
        }
        GSYNC();
        GFOR(idx, n * p) {  // PZ_bar = N_bar + G + Zm' F_bar
            const int i = decp[idx] >> 16, b = decp[idx] & 0xffff;
            double s = PZb[i * ps + b] + G1[i * ps + b];
            for (int c = 0; c < p; ++c) s = fma(w[c] * Zs[c * n + i], 0.5 * (Fb[c * ps + b] + Fb[b * ps + c]), s);
            PZb[i * ps + b] = s;
        }
        GFOR(c, p) {
            hb[c] += w[c] * Fb[c * ps + c];
            db[c] -= (g.mask_intercept ? w[c] : 1.0) * vb[c];
        }
        GSYNC();
        // P_bar = Pf_bar + PZ_bar Zm
        gemm_sq<true>(n, p, [&](int i, int c) { return PZb[i * ps + c] * w[c]; }, [&](int c, int j) { return Zs[c * n + j]; },
                      [&](int i, int j, double v_) { Pb[i * ld + j] = v_ + Pfb[i * ld + j]; }, decg);
        if (g.Z_bar) {  // Zm = diag(w) Z enters v = ym - d - Zm a, PZ = P Zm', F = Zm PZ + ...
            GFOR(idx, p * n) {
                const int c = idx / n, j = idx - c * n;
                double s = -vb[c] * a[j];
                for (int i = 0; i < n; ++i) s = fma(PZb[i * ps + c], P[i * ld + j], s);
                for (int b = 0; b < p; ++b) s = fma(0.5 * (Fb[c * ps + b] + Fb[b * ps + c]), PZ[j * ps + b], s);
                Zb[idx] += w[c] * s;
            }
        }
        GFOR(j, n) {
            double s = afb[j];
            for (int c = 0; c < p; ++c) s = fma(-w[c] * Zs[c * n + j], vb[c], s);
            ab[j] = s;
        }
        GSYNC();
    }
    // ---- P0 = dlyap(T, C0): S = T' S T + P0_bar by doubling (S in Pb, A in Pfb);  P still holds P0
    GFOR(idx, n * ld) Pfb[idx] = Tm[idx];
    GSYNC();
    {
        const int cap = g.lyap_max_iter > 0 ? g.lyap_max_iter : 64;
        for (int it = 0; it < cap; ++it) {
            mm<false, false>(W1, ld, Pb, ld, Pfb, ld, n, n, n, 1.0, 0.0);
            GSYNC();
            mm<true, false>(W2, ld, Pfb, ld, W1, ld, n, n, n, 1.0, 0.0);
            mm<false, false>(Pf, ld, Pfb, ld, Pfb, ld, n, n, n, 1.0, 0.0);
            GSYNC();
            GFOR(idx, n * n) {
                const int i = idx / n, j = idx - i * n;
                Pb[i * ld + j] += W2[i * ld + j];
                Pfb[i * ld + j] = Pf[i * ld + j];
            }
            GSYNC();
            const double dmax = absmax(W2, ld, n, n, s_red), smax = absmax(Pb, ld, n, n, s_red);
            if (!(dmax > 1e-17 * smax)) break;
        }
    }
    // C0_bar += S;  T_bar = accumulated + (S + S') T P0
    mm<false, false>(W1, ld, Tm, ld, P, ld, n, n, n, 1.0, 0.0);
    GFOR(idx, n * n) gC0b[idx] += Pb[(idx / n) * ld + (idx % n)];
    GSYNC();
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        double s = Tb[i * ld + j];
        for (int kk = 0; kk < n; ++kk) s = fma(Pb[i * ld + kk] + Pb[kk * ld + i], W1[kk * ld + j], s);
        gTb[idx] = s;
        W2[i * ld + j] = gC0b[idx] + gC0b[j * n + i];  // C0_bar + C0_bar'
    }
    GSYNC();
    if (g.qfull) {
        // C0 = R Q R':  R_bar = C0_bar R Q' + C0_bar' R Q,  Q_bar = R' C0_bar R      (W1 <- C0_bar R, Pf <- C0_bar' R)
        GFOR(idx, n * k) {
            const int i = idx / k, c = idx - i * k;
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < n; ++j) {
                s1 = fma(gC0b[i * n + j], Rs[j * k + c], s1);
                s2 = fma(gC0b[j * n + i], Rs[j * k + c], s2);
            }
            W1[i * ld + c] = s1;
            Pf[i * ld + c] = s2;
        }
        GSYNC();
        GFOR(idx, n * k) {
            const int i = idx / k, c = idx - i * k;
            double s = 0.0;
            for (int c2 = 0; c2 < k; ++c2) s = fma(W1[i * ld + c2], Qs[c * k + c2], fma(Pf[i * ld + c2], Qs[c2 * k + c], s));
            gRb[idx] = s;
        }
        GFOR(idx, k * k) {
            const int a_ = idx / k, b_ = idx - a_ * k;
            // symmetrised: Q is a covariance, and how dll/dQ splits between Q_ab and Q_ba depends on how the filter is extended to
            // non-symmetric covariances (the Joseph form and the rank-p form extend it differently); only the symmetric part is defined
            double s = 0.0;
            for (int i = 0; i < n; ++i) s = fma(Rs[i * k + a_], W1[i * ld + b_], fma(Rs[i * k + b_], W1[i * ld + a_], s));
            g.qfull_bar[(size_t)draw * k * k + idx] = 0.5 * s;
        }
    } else {
    // R_bar = (C0_bar + C0_bar') R Q;  q_bar_c = sum_ij R_ic C0_bar_ij R_jc = (1/2) sum_i R_ic ((C0_bar + C0_bar') R)_ic
    GFOR(idx, n * k) {
        const int i = idx / k, c = idx - i * k;
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(W2[i * ld + j], Rs[j * k + c], s);
        W1[i * ld + c] = s;
        gRb[idx] = s * qs[c];
    }
    GSYNC();
    GFOR(c, k) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s = fma(Rs[i * k + c], W1[i * ld + c], s);
        s *= 0.5;
        if (g.sigma_inputs) s *= 2.0 * g.qdiag[(size_t)draw * g.q_stride + c];
        g.q_bar[(size_t)draw * k + c] = s;
    }
    }
    GFOR(c, p) {
        if (g.h_bar) {
            double s = hb[c];
            if (g.sigma_inputs) s *= 2.0 * (g.hdiag ? g.hdiag[(size_t)draw * g.h_stride + c] : 0.0);
            g.h_bar[(size_t)draw * p + c] = s;
        }
        if (g.d_bar) g.d_bar[(size_t)draw * p + c] = db[c];
    }
    if (g.Z_bar) GFOR(idx, p * n) g.Z_bar[(size_t)draw * p * n + idx] = Zb[idx];
    if (G_TID == 0) {
        g.ll[draw] = ll;
        g.status[draw] = status;
    }
    GSYNC();
}

// -----------------------------------------------------------------------------------------------------------------
struct PolicyAdjointArgs {
    const double* A;  // [N][n][n]  (unused by the arithmetic: A_bar = S does not depend on A; kept for the reference's signature)
    const double* B;
    const double* C;
    const double* D;      // [N][n][k] or NULL
    const double* T;      // [N][n][n]
    const double* R;      // [N][n][k] or NULL
    const double* T_bar;  // [N][n][n]
    const double* R_bar;  // [N][n][k] or NULL
    long long N;
    int n, k, max_iter;
    double* A_bar;    // [N][n][n]
    double* B_bar;    // [N][n][n]
    double* C_bar;    // [N][n][n]
    double* D_bar;    // [N][n][k] or NULL
    int32_t* status;  // [N] or NULL: GECON_ST_SINGULAR when C T + B is singular (outputs NaN)
};

GHH size_t policy_adjoint_smem_doubles(int n, int nt) { return (size_t)6 * n * ldim(n) + nt + 8; }

GHD void policy_adjoint_draw(const PolicyAdjointArgs& g, long long draw, double* sm, int* s_int) {
    const int n = g.n, k = (g.R_bar && g.D && g.R) ? g.k : 0, ld = ldim(n);
    const int tile = n * ld;
    double* X0 = sm;  // W -> W^-1
    double* X1 = X0 + tile;
    double* X2 = X1 + tile;
    double* X3 = X2 + tile;
    double* X4 = X3 + tile;  // S
    double* X5 = X4 + tile;  // G
    double* s_red = X5 + tile;
    const size_t o = (size_t)draw * n * n;
    const double *gB = g.B + o, *gC = g.C + o, *gT = g.T + o, *gTb = g.T_bar + o;
    double *oA = g.A_bar + o, *oB = g.B_bar + o, *oC = g.C_bar + o;
    // W = C T + B  (X1), identity (X0)
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        double s = gB[idx];
        for (int kk = 0; kk < n; ++kk) s = fma(gC[i * n + kk], gT[kk * n + j], s);
        X1[i * ld + j] = s;
        X0[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    GSYNC();
    // Gauss-Jordan with partial pivoting on [X1 | X0] -> X0 = W^-1
    bool singular = false;
    for (int c = 0; c < n; ++c) {
        if (G_TID == 0) {
            int piv = c;
            double best = fabs(X1[c * ld + c]);
            for (int i = c + 1; i < n; ++i) {
                const double x = fabs(X1[i * ld + c]);
                if (x > best) {
                    best = x;
                    piv = i;
                }
            }
            s_int[0] = piv;
            s_int[1] = (best > 0.0 && best <= 1.7e308) ? 1 : 0;
        }
        GSYNC();
        const int piv = s_int[0];
        if (!s_int[1]) {
            singular = true;
            break;
        }
        if (piv != c) {
            GFOR(j, n) {
                double t = X1[c * ld + j];
                X1[c * ld + j] = X1[piv * ld + j];
                X1[piv * ld + j] = t;
                t = X0[c * ld + j];
                X0[c * ld + j] = X0[piv * ld + j];
                X0[piv * ld + j] = t;
            }
            GSYNC();
        }
        const double inv = 1.0 / X1[c * ld + c];
        GSYNC();
        GFOR(j, n) {
            X1[c * ld + j] *= inv;
            X0[c * ld + j] *= inv;
        }
        GSYNC();
        GFOR(i, n) X2[i * ld] = (i == c) ? 0.0 : X1[i * ld + c];  // multipliers out of the way of the row updates below
        GSYNC();
        GFOR(idx, n * n) {
            const int i = idx / n, j = idx - i * n;
            if (i == c) continue;  // the pivot row is read by everyone in this phase: leave it alone (its multiplier is 0)
            const double f = X2[i * ld];
            X1[i * ld + j] = fma(-f, X1[c * ld + j], X1[i * ld + j]);
            X0[i * ld + j] = fma(-f, X0[c * ld + j], X0[i * ld + j]);
        }
        GSYNC();
    }
    if (singular) {
        const double nanv = NAN;
        GFOR(idx, n * n) oA[idx] = oB[idx] = oC[idx] = nanv;
        if (g.D_bar) GFOR(idx, n * g.k) g.D_bar[(size_t)draw * n * g.k + idx] = nanv;
        if (G_TID == 0 && g.status) g.status[draw] = ST_SINGULAR;
        GSYNC();
        return;
    }
    // selection matrix in reverse: D_bar = -W^-T R_bar, W_bar = D_bar R', B_bar = W_bar, C_bar = W_bar T', T_bar += C' W_bar
    if (k > 0) {
        const double* gRb = g.R_bar + (size_t)draw * n * k;
        const double* gR = g.R + (size_t)draw * n * k;
        GFOR(idx, n * k) {
            const int i = idx / k, c = idx - i * k;
            double s = 0.0;
            for (int kk = 0; kk < n; ++kk) s = fma(X0[kk * ld + i], gRb[kk * k + c], s);
            X1[i * ld + c] = -s;
            if (g.D_bar) g.D_bar[(size_t)draw * n * k + idx] = -s;
        }
        GSYNC();
        GFOR(idx, n * n) {
            const int i = idx / n, j = idx - i * n;
            double s = 0.0;
            for (int c = 0; c < k; ++c) s = fma(X1[i * ld + c], gR[j * k + c], s);
            X2[i * ld + j] = s;  // W_bar
        }
        GSYNC();
        GFOR(idx, n * n) {
            const int i = idx / n, j = idx - i * n;
            double s = 0.0, u = gTb[idx];
            for (int kk = 0; kk < n; ++kk) {
                s = fma(X2[i * ld + kk], gT[j * n + kk], s);   // W_bar T'
                u = fma(gC[kk * n + i], X2[kk * ld + j], u);  // T_bar + C' W_bar
            }
            oB[idx] = X2[i * ld + j];
            oC[idx] = s;
            X3[i * ld + j] = u;
        }
    } else {
        if (g.D_bar) GFOR(idx, n * g.k) g.D_bar[(size_t)draw * n * g.k + idx] = 0.0;
        GFOR(idx, n * n) {
            const int i = idx / n, j = idx - i * n;
            oB[idx] = 0.0;
            oC[idx] = 0.0;
            X3[i * ld + j] = gTb[idx];
        }
    }
    GSYNC();
    // S = Q = -W^-T T_bar_total (X4),  G = -W^-T C' (X5),  T_k = T' (X1)
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        double s = 0.0, u = 0.0;
        for (int kk = 0; kk < n; ++kk) {
            s = fma(X0[kk * ld + i], X3[kk * ld + j], s);
            u = fma(X0[kk * ld + i], gC[j * n + kk], u);
        }
        X4[i * ld + j] = -s;
        X5[i * ld + j] = -u;
        X1[i * ld + j] = gT[j * n + i];
    }
    GSYNC();
    const int cap = g.max_iter > 0 ? g.max_iter : 64;
    for (int it = 0; it < cap; ++it) {
        mm<false, false>(X2, ld, X5, ld, X4, ld, n, n, n, 1.0, 0.0);  // G S
        GSYNC();
        mm<false, false>(X3, ld, X2, ld, X1, ld, n, n, n, 1.0, 0.0);  // (G S) T_k
        mm<false, false>(X0, ld, X5, ld, X5, ld, n, n, n, 1.0, 0.0);  // G^2
        GSYNC();
        mm<false, false>(X2, ld, X1, ld, X1, ld, n, n, n, 1.0, 0.0);  // T_k^2
        GFOR(idx, n * n) {
            const int i = idx / n, j = idx - i * n;
            X4[i * ld + j] += X3[i * ld + j];
            X5[i * ld + j] = X0[i * ld + j];
        }
        GSYNC();
        GFOR(idx, n * n) X1[(idx / n) * ld + (idx % n)] = X2[(idx / n) * ld + (idx % n)];
        const double dmax = absmax(X3, ld, n, n, s_red), smax = absmax(X4, ld, n, n, s_red);
        if (!(dmax > 1e-17 * smax)) break;
    }
    GSYNC();
    // A_bar = S, B_bar += S T', C_bar += S T' T'
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        double s = 0.0;
        for (int kk = 0; kk < n; ++kk) s = fma(X4[i * ld + kk], gT[j * n + kk], s);
        X2[i * ld + j] = s;
        oA[idx] = X4[i * ld + j];
        oB[idx] += s;
    }
    GSYNC();
    GFOR(idx, n * n) {
        const int i = idx / n, j = idx - i * n;
        double s = 0.0;
        for (int kk = 0; kk < n; ++kk) s = fma(X2[i * ld + kk], gT[j * n + kk], s);
        oC[idx] += s;
    }
    if (G_TID == 0 && g.status) g.status[draw] = 0;
    GSYNC();
}

}  // namespace gecon_grad
