// Reverse mode of the perturbation solution on the fp64 tensor path: one CTA per draw, every matrix an NP x LD shared-memory tile.
//
// Same mathematics as policy_adjoint_draw (grad.cuh; the host-checkable DFMA version, which stays the executable specification and
// the fallback for n > 64): gEconpy/solvers/shared.py:12-75 (o1_policy_function_adjoints, the `pullback` of every solver Op) and
// R = -(C T + B)^-1 D in reverse.
//   W = C T + B,  X = W^-T (blocked Gauss-Jordan on W' with the identity as right-hand side)
//   D_bar = -X R_bar,  W_bar = D_bar R',  B_bar = W_bar,  C_bar = W_bar T',  T_bar_tot = T_bar + C' W_bar
//   Stein equation S = Q + G S T' (the reference's n^2 x n^2 Kronecker system, shared.py:53-71), Q = -X T_bar_tot, G = -X C', by
//   doubling:  S += G_k S T_k,  G_{k+1} = G_k^2,  T_{k+1} = T_k^2,  T_0 = T'    until max|increment| <= 1e-17 max|S|
//   A_bar = S,  B_bar += S T',  C_bar += S T' T'
// Round-2 motivation: with the Kalman adjoint on its rank-p form the DFMA version (one pivot column per barrier, 128 threads, two
// shared loads per FMA) was 17 of the gradient step's 60 ms for 32,768 medium-NK draws.
#include "common.cuh"
#include "grad_args.h"
#include "linalg.cuh"

namespace gecon {

template <int NP>
struct PaSmem {
    static constexpr size_t bytes = sizeof(double) * (7 * Cfg<NP>::TILE + 2 * NP) + sizeof(int) * (2 * NP + 8);
};

// dst[c][r] = src[r][c] for the n x m corner of a global row-major matrix (zero padded): the transposed tile
template <int NP>
__device__ __forceinline__ void tile_load_t(double* __restrict__ dst, const double* __restrict__ src, int rows, int cols, int ldg) {
    constexpr int LD = Cfg<NP>::LD;
    for (int i = threadIdx.x; i < Cfg<NP>::TILE; i += Cfg<NP>::NT) {
        const int r = i / LD, c = i - r * LD;  // element (r, c) of the tile = element (c, r) of the source
        dst[i] = (c < rows && r < cols) ? src[(size_t)c * ldg + r] : 0.0;
    }
}

template <int NP>
__device__ __forceinline__ double acc_absmax(const Acc<NP>& a) {
    double m = 0.0;
#pragma unroll
    for (int ct = 0; ct < NP / 8; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double v = fabs(a.v[ct][e]);
            if (v > m || v != v) m = v;
        }
    return m;
}

template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT) policy_adjoint_dmma_kernel(const gecon_grad::PolicyAdjointArgs g) {
    using C = Cfg<NP>;
    constexpr int LD = C::LD, NT = C::NT;
    extern __shared__ __align__(16) double sm[];
    double* Ct = sm;               // C                      -> scratch of the doubling loop
    double* Tt = Ct + C::TILE;     // T (kept to the end)
    double* X = Tt + C::TILE;      // W' -> W^-T             -> scratch
    double* Wb = X + C::TILE;      // identity / panels / W_bar -> scratch
    double* S = Wb + C::TILE;      // T_bar_tot -> S
    double* G = S + C::TILE;
    double* Tk = G + C::TILE;
    double* s_inv = Tk + C::TILE;
    double* s_red = s_inv + NP;
    int* s_piv = reinterpret_cast<int*>(s_red + NP);
    int* s_flag = s_piv + NP;
    const int n = g.n, k = (g.R_bar && g.D && g.R) ? g.k : 0;
    const int nt8 = (n + 7) >> 3, kt8 = (k + 7) >> 3, k4 = (k + 3) & ~3, n4 = (n + 3) & ~3;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    for (long long draw = blockIdx.x; draw < g.N; draw += gridDim.x) {
        const size_t o = (size_t)draw * n * n;
        const double *gB = g.B + o, *gC = g.C + o, *gT = g.T + o, *gTb = g.T_bar + o;
        double *oA = g.A_bar + o, *oB = g.B_bar + o, *oC = g.C_bar + o;
        // ---- W' = T' C' + B' (into X), identity (into Wb)
        tile_load<NP>(Ct, gC, n, n, n);
        tile_load<NP>(Tt, gT, n, n, n);
        tile_load_t<NP>(X, gB, n, n, n);
        for (int i = threadIdx.x; i < C::TILE; i += NT) Wb[i] = (i / LD == i % LD && i / LD < n) ? 1.0 : 0.0;
        __syncthreads();
        {
            Acc<NP> acc;
            acc_load<NP>(acc, X);
            gemm_acc<NP, true, true>(acc, Tt, Ct, 1.0, 0, n4);
            __syncthreads();
            acc_store<NP>(acc, X);
        }
        __syncthreads();
        const bool ok = gj_solve_blocked<NP>(X, X, Wb, Wb, 0, nt8, nullptr, nullptr, 0, 0, n, true, s_piv, s_flag);
        if (!ok) {  // singular C T + B: NaN outputs + status, never an exception
            for (int i = threadIdx.x; i < n * n; i += NT) oA[i] = oB[i] = oC[i] = qnan;
            if (g.D_bar) {
                for (int i = threadIdx.x; i < n * g.k; i += NT) g.D_bar[(size_t)draw * n * g.k + i] = qnan;
            }
            if (threadIdx.x == 0 && g.status) g.status[draw] = GECON_ST_SINGULAR;
            __syncthreads();
            continue;
        }
        tile_copy<NP>(X, Wb);  // X = W^-T
        __syncthreads();
        // ---- selection matrix in reverse
        if (k > 0) {
            const double* gRb = g.R_bar + (size_t)draw * n * k;
            const double* gR = g.R + (size_t)draw * n * k;
            tile_load<NP>(S, gRb, n, k, k);  // R_bar (n x k)
            tile_load<NP>(G, gR, n, k, k);   // R     (n x k)
            __syncthreads();
            {
                Acc<NP> acc;  // D_bar = -X R_bar  -> Tk (n x k)
                acc_zero<NP>(acc);
                gemm_acc<NP, false, false>(acc, X, S, -1.0, 0, n4, 0, kt8);
                acc_store<NP>(acc, Tk, 0, kt8);
            }
            __syncthreads();
            if (g.D_bar) tile_store<NP>(g.D_bar + (size_t)draw * n * k, Tk, n, k, k, 1.0, nullptr, nullptr);
            {
                Acc<NP> acc;  // W_bar = D_bar R'  -> Wb
                acc_zero<NP>(acc);
                gemm_acc<NP, false, true>(acc, Tk, G, 1.0, 0, k4);
                acc_store<NP>(acc, Wb);
            }
            __syncthreads();
            tile_store<NP>(oB, Wb, n, n, n, 1.0, nullptr, nullptr);  // B_bar = W_bar (S T' is added at the end)
            tile_load<NP>(S, gTb, n, n, n);
            __syncthreads();
            {
                Acc<NP> acc;  // C_bar = W_bar T'  -> G (stored to global now, S T' T' added at the end)
                acc_zero<NP>(acc);
                gemm_acc<NP, false, true>(acc, Wb, Tt, 1.0, 0, n4);
                acc_store<NP>(acc, G);
                Acc<NP> tb;   // T_bar_tot = T_bar + C' W_bar  -> S
                acc_load<NP>(tb, S);
                gemm_acc<NP, true, false>(tb, Ct, Wb, 1.0, 0, n4);
                __syncthreads();
                acc_store<NP>(tb, S);
            }
            __syncthreads();
            tile_store<NP>(oC, G, n, n, n, 1.0, nullptr, nullptr);
            __syncthreads();
        } else {
            if (g.D_bar) {
                for (int i = threadIdx.x; i < n * g.k; i += NT) g.D_bar[(size_t)draw * n * g.k + i] = 0.0;
            }
            for (int i = threadIdx.x; i < n * n; i += NT) oB[i] = oC[i] = 0.0;
            tile_load<NP>(S, gTb, n, n, n);
            __syncthreads();
        }
        // ---- Q = -X T_bar_tot (-> Wb, then S),  G = -X C',  T_0 = T'
        {
            Acc<NP> q, gg;
            acc_zero<NP>(q);
            acc_zero<NP>(gg);
            gemm_acc<NP, false, false>(q, X, S, -1.0, 0, n4);
            gemm_acc<NP, false, true>(gg, X, Ct, -1.0, 0, n4);
            __syncthreads();
            acc_store<NP>(q, S);
            acc_store<NP>(gg, G);
        }
        tile_load_t<NP>(Tk, gT, n, n, n);
        __syncthreads();
        // ---- doubling: S += G S T_k, G <- G^2, T_k <- T_k^2     (scratch: Ct = G S, X = G^2, Wb = T_k^2)
        const int cap = g.max_iter > 0 ? g.max_iter : 64;
        for (int it = 0; it < cap; ++it) {
            {
                Acc<NP> a;
                acc_zero<NP>(a);
                gemm_acc<NP, false, false>(a, G, S, 1.0, 0, n4);
                acc_store<NP>(a, Ct);
            }
            __syncthreads();
            Acc<NP> inc, s_new, g2, t2;
            acc_zero<NP>(inc);
            gemm_acc<NP, false, false>(inc, Ct, Tk, 1.0, 0, n4);
            acc_load<NP>(s_new, S);
#pragma unroll
            for (int ct = 0; ct < NP / 8; ++ct) {
                s_new.v[ct][0] += inc.v[ct][0];
                s_new.v[ct][1] += inc.v[ct][1];
            }
            acc_zero<NP>(g2);
            acc_zero<NP>(t2);
            gemm_acc<NP, false, false>(g2, G, G, 1.0, 0, n4);
            gemm_acc<NP, false, false>(t2, Tk, Tk, 1.0, 0, n4);
            const double dmax = block_max<NP>(acc_absmax<NP>(inc), s_red);   // (barriers inside: every read of S, G, T_k is done)
            const double smax = block_max<NP>(acc_absmax<NP>(s_new), s_red);
            acc_store<NP>(s_new, S);
            acc_store<NP>(g2, G);
            acc_store<NP>(t2, Tk);
            __syncthreads();
            if (!(dmax > 1e-17 * smax)) break;
        }
        // ---- A_bar = S,  B_bar += S T',  C_bar += S T' T'
        tile_store<NP>(oA, S, n, n, n, 1.0, nullptr, nullptr);
        {
            Acc<NP> a;
            acc_zero<NP>(a);
            gemm_acc<NP, false, true>(a, S, Tt, 1.0, 0, n4);
            acc_store<NP>(a, Ct);
        }
        __syncthreads();
        {
            Acc<NP> a;
            acc_zero<NP>(a);
            gemm_acc<NP, false, true>(a, Ct, Tt, 1.0, 0, n4);
            acc_store<NP>(a, X);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n * n; i += NT) {
            const int r = i / n, c = i - r * n;
            oB[i] += Ct[r * LD + c];
            oC[i] += X[r * LD + c];
        }
        if (threadIdx.x == 0 && g.status) g.status[draw] = 0;
        __syncthreads();
    }
}

template <int NP>
static int launch_pa(const gecon_grad::PolicyAdjointArgs& g, cudaStream_t st) {
    int grid = 0;
    int rc = persistent_grid(policy_adjoint_dmma_kernel<NP>, Cfg<NP>::NT, PaSmem<NP>::bytes, g.N, &grid, nullptr);
    if (rc) return rc;
    policy_adjoint_dmma_kernel<NP><<<grid, Cfg<NP>::NT, PaSmem<NP>::bytes, st>>>(g);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

// n <= 64 and k <= padded n: the tensor-path kernel; returns GECON_E_UNSUPPORTED_SIZE otherwise (the caller falls back to the DFMA kernel)
int policy_adjoint_dmma_launch(const gecon_grad::PolicyAdjointArgs& g, cudaStream_t st) {
    const int np = round_up8(g.n);
    GECON_DISPATCH_NP(np, return launch_pa<NP_>(g, st));
    return 0;
}

}  // namespace gecon
