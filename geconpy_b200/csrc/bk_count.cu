// Batched Blanchard-Kahn eigenvalue count, one CTA per parameter draw.
//
// Reference quantity (gEconpy/model/perturbation.py:448-505,586-625 check_bk_condition_pt; pencil assembly as in
// gEconpy/solvers/gensys.py:568-614; count rule gEconpy/pytensorf/real_eig.py:31-36 + perturbation.py:618-620):
//     Gamma0 = [[B, C], [-I, 0]],  Gamma1 = [[A, 0], [0, I]],  rows/cols sel = {0..n-1} U {n + lead_idx},
//     G = -Gamma0[sel, sel] + 1e-8 I,  M = G^-1 Gamma1[sel, sel],  n_unstable = #{ |eig(M)| > 1 }.
//
// The count is computed without an eigen-solver.  The Cayley transform N = (Gamma1 - G)(Gamma1 + G)^-1 is similar to
// (M - I)(M + I)^-1, which maps |lambda| > 1 to Re > 0 (infinite eigenvalues of the pencil go to +1, zeros to -1), and
//     #{Re eig(N) > 0} = (m + trace sign(N)) / 2,
// with sign(N) from the determinant-scaled Newton iteration S <- (mu S + (mu S)^-1) / 2.  Only Gauss-Jordan inverses
// and elementwise updates are needed, i.e. the same in-CTA primitives as the cycle-reduction kernel.  The iteration
// works on N' (trace and sign commute with transposition), which is what one solve with (Gamma1 + G)' produces.
// N' is balanced first (Parlett-Reinsch scaling, eig.cuh).  The sign iteration is only trusted when it is well conditioned: a trace that does not settle on an integer (an eigenvalue on
// or next to the unit circle), or a sign matrix with entries above 1e6 (nearly parallel stable / unstable invariant subspaces:
// the Newton iteration's error grows like eps ||S||^2, and wide priors produce pencils with ||S|| ~ 1e9 ... 1e100), sends the
// draw to the FALLBACK: warp 0 of the CTA runs the Hessenberg + Francis QR eigenvalue routine (eig.cuh) on N' and counts
// Re(mu) > 0 directly -- the backward-stable route, on a matrix with O(1) entries, which agrees with QZ on the pencil where
// dgeev on M (entries 1e8) is itself only good to 1e-4.  Only a draw on which that fails too (singular Gamma1 + G, QR sweep
// not converging) is flagged GECON_ST_BK_INCONCLUSIVE instead of being guessed.
#include "common.cuh"
#include "eig.cuh"
#include "linalg.cuh"

namespace gecon {

template <int NP>
struct BkSmem {
    static constexpr size_t bytes = sizeof(double) * (3 * Cfg<NP>::TILE + 2 * NP) + sizeof(int) * (3 * NP + 8);
};

// Resident CTAs per SM the register allocator must leave room for.  The kernel is bound by the barriers of its Gauss-Jordan solves, so
// co-resident CTAs pay directly; shared memory (three tiles) allows 8 / 5 / 3 / 2 / 2 CTAs at NP = 32 / 40 / 48 / 56 / 64, and without a
// bound ptxas takes 168-180 registers (2 CTAs at NP = 40).  Measured on the wide-prior population (m = 38, NP = 40): see DESIGN 3.3.
template <int NP>
constexpr int bk_min_ctas() {
    return NP <= 24 ? 1 : NP <= 40 ? 4 : NP <= 48 ? 3 : NP <= 64 ? 2 : 1;
}

template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT, bk_min_ctas<NP>()) bk_count_kernel(const gecon_bk_args p, const gecon_compact_jac cj, double* __restrict__ scratch) {
    using C = Cfg<NP>;
    constexpr int LD = C::LD, NT = C::NT;
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* W = S + C::TILE;
    double* X = W + C::TILE;
    double* s_inv = X + C::TILE;
    double* s_red = s_inv + NP;
    int* s_piv = reinterpret_cast<int*>(s_red + NP);
    int* s_sel = s_piv + NP;
    int* s_flag = s_sel + NP;  // [NP + 1]

    const int n = p.n, nl = p.n_lead, m = n + nl;
    const int cap = p.max_iter > 0 ? p.max_iter : 60;
    for (int i = threadIdx.x; i < m; i += NT) s_sel[i] = (i < n) ? i : n + p.lead_idx[i - n];
    __syncthreads();

    for (long long draw = blockIdx.x; draw < p.N; draw += gridDim.x) {
        if (p.accumulate && (p.status[draw] & p.skip_mask)) continue;  // uniform: already decided upstream
        const double *gA, *gB, *gC;
        if (cj.vals) {  // compact Jacobian: expand A, B, C of this draw into the CTA's dense scratch
            double* sc = scratch + (size_t)blockIdx.x * (size_t)3 * n * n;
            expand_compact(sc, cj, draw, n, 0, 0, 2);
            gA = sc;
            gB = sc + (size_t)n * n;
            gC = sc + (size_t)2 * n * n;
        } else {
            gA = p.A + (size_t)draw * n * n;
            gB = p.B + (size_t)draw * n * n;
            gC = p.C + (size_t)draw * n * n;
        }
        // W = (Gamma1 + G)',  X = (Gamma1 - G)'   (element [r][c] of the transpose = element (c, r))
        auto assemble = [&]() {
            for (int i = threadIdx.x; i < C::TILE; i += NT) {
                const int r = i / LD, c = i - r * LD;
                double wv = 0.0, xv = 0.0;
                if (r < m && c < m) {
                    const int R = s_sel[c], Cc = s_sel[r];  // (row, col) in the 2n x 2n pencil
                    double g0, g1;
                    if (R < n) {
                        g0 = (Cc < n) ? gB[R * n + Cc] : gC[R * n + (Cc - n)];
                        g1 = (Cc < n) ? gA[R * n + Cc] : 0.0;
                    } else {
                        g0 = (Cc == R - n) ? -1.0 : 0.0;
                        g1 = (Cc == R) ? 1.0 : 0.0;
                    }
                    const double g = -g0 + ((r == c) ? 1e-8 : 0.0);
                    wv = g1 + g;
                    xv = g1 - g;
                }
                W[i] = wv;
                X[i] = xv;
            }
        };
        assemble();
        __syncthreads();
        const int mt = (m + 7) >> 3;
        bool ok = gj_solve_blocked<NP>(W, W, X, X, 0, mt, nullptr, nullptr, 0, 0, m, true, s_piv, s_flag, s_inv);
        // Balance N' (diagonal similarity by powers of two: spectrum and trace of the sign unchanged).  Wide priors produce pencils
        // whose rows differ by 30 orders of magnitude; unbalanced, the sign matrix of such an N has entries of 1e9 ... 1e100 and
        // the Newton iteration is useless (numpy emulation on the nk_wide population: reliable on 64 % of the finite draws without
        // balancing, 100 % with it).
        if (ok && threadIdx.x < 32) warp_balance_scale(X, LD, m, (int)threadIdx.x);
        __syncthreads();
        tile_copy<NP>(S, X);
        __syncthreads();

        bool settled = false;
        int it = 0;
        double smax_last = 0.0;
        while (ok && it < cap) {
            ++it;
            tile_copy<NP>(W, S);
            for (int i = threadIdx.x; i < C::TILE; i += NT) {
                const int r = i / LD, c = i - r * LD;
                X[i] = (r == c && r < m) ? 1.0 : 0.0;
            }
            __syncthreads();
            ok = gj_solve_blocked<NP>(W, W, X, X, 0, mt, nullptr, nullptr, 0, 0, m, true, s_piv, s_flag, s_inv);
            if (!ok) break;
            // determinant scaling: mu = |det S|^(-1/m) = exp(mean log |1/pivot|)
            double lg = ((int)threadIdx.x < m) ? log(fabs(s_inv[threadIdx.x])) : 0.0;
            lg = block_sum<NP>(lg, s_red);
            double mu = exp(lg / m);
            if (!(mu > 1e-150 && mu < 1e150)) mu = 1.0;
            const double hm = 0.5 * mu, hi = 0.5 / mu;
            double dmax = 0.0, smax = 0.0;
            for (int i = threadIdx.x; i < C::TILE; i += NT) {
                const double so = S[i];
                const double sn = hm * so + hi * X[i];
                S[i] = sn;
                const double d = fabs(sn - so), a = fabs(sn);
                if (d > dmax || d != d) dmax = d;
                if (a > smax || a != a) smax = a;
            }
            dmax = block_max<NP>(dmax, s_red);
            smax = block_max<NP>(smax, s_red);
            smax_last = smax;
            if (dmax != dmax || smax != smax) {
                ok = false;
                break;
            }
            if (dmax <= 1e-6 * (1.0 + smax)) {
                settled = true;
                break;
            }
        }
        __syncthreads();
        double tr = ((int)threadIdx.x < m) ? S[threadIdx.x * LD + threadIdx.x] : 0.0;
        tr = block_sum<NP>(tr, s_red);
        const double cnt = 0.5 * (m + tr);
        const double rc = rint(cnt);
        int nu = -1;
        if (ok && settled && fabs(cnt - rc) < 0.05 && rc >= 0.0 && rc <= (double)m && smax_last <= 1e6) {
            nu = (int)rc;
        } else {
            // ---- fallback: eigenvalues of N' by the QR routine (warp 0), count Re(mu) > 0
            assemble();
            __syncthreads();
            const bool ok2 = gj_solve_blocked<NP>(W, W, X, X, 0, mt, nullptr, nullptr, 0, 0, m, true, s_piv, s_flag, s_inv);
            __syncthreads();
            if (threadIdx.x < 32) {
                const int lane = threadIdx.x;
                int res = -1;
                bool finite = ok2;
                if (ok2) {
                    for (int i = lane; i < m * m; i += 32) finite = finite && (fabs(X[(i / m) * LD + (i % m)]) <= 1.7e308);
                }
                finite = __all_sync(0xffffffffu, finite);
                if (finite) {
                    double* ort = W;  // the solve has destroyed W: scratch
                    double* wr = W + NP;
                    double* wi = W + 2 * NP;
                    if (warp_real_eig(X, LD, m, 1, ort, wr, wi, lane)) {
                        __syncwarp();
                        int c = 0;
                        bool fin = true;
                        for (int i = lane; i < m; i += 32) {
                            c += (wr[i] > 0.0) ? 1 : 0;
                            fin = fin && (fabs(wr[i]) <= 1.7e308) && (fabs(wi[i]) <= 1.7e308);
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
                        if (__all_sync(0xffffffffu, fin)) res = c;
                    }
                }
                if (lane == 0) s_flag[0] = res;
            }
            __syncthreads();
            nu = s_flag[0];
        }
        if (threadIdx.x == 0) {
            int st = 0;
            if (nu >= 0) {
                if (nu != nl) st |= GECON_ST_BK;
            } else {
                st |= GECON_ST_BK | GECON_ST_BK_INCONCLUSIVE;
            }
            if (p.n_unstable) p.n_unstable[draw] = nu;
            p.status[draw] = p.accumulate ? (p.status[draw] | st) : st;
        }
        __syncthreads();
    }
}

template <int NP>
static int launch_bk(const gecon_bk_args& a, cudaStream_t st) {
    int grid = 0;
    int rc = persistent_grid(bk_count_kernel<NP>, Cfg<NP>::NT, BkSmem<NP>::bytes, a.N, &grid, nullptr);
    if (rc) return rc;
    gecon_compact_jac cj{};
    double* scratch = nullptr;
    if (a.compact) {
        cj = *a.compact;
        GECON_CUDA(cudaMallocAsync((void**)&scratch, sizeof(double) * (size_t)grid * (size_t)3 * a.n * a.n, st));
    }
    bk_count_kernel<NP><<<grid, Cfg<NP>::NT, BkSmem<NP>::bytes, st>>>(a, cj, scratch);
    g_launch_count++;
    const cudaError_t le = cudaGetLastError();
    if (scratch) cudaFreeAsync(scratch, st);
    GECON_CUDA(le);
    return 0;
}

int bk_kernel_info(int m, int* ctas, int* smem, int* threads) {
    const int np = round_up8(m);
    GECON_DISPATCH_NP_WIDE(np, {
        int grid = 0;
        int rc = persistent_grid(bk_count_kernel<NP_>, Cfg<NP_>::NT, BkSmem<NP_>::bytes, 1 << 30, &grid, ctas);
        if (rc) return rc;
        *smem = (int)BkSmem<NP_>::bytes;
        *threads = Cfg<NP_>::NT;
    });
    return 0;
}

static int check_bk_args(const gecon_bk_args* a) {
    if (!a || a->struct_size != sizeof(gecon_bk_args)) {
        set_last_error("gecon_bk_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if ((!a->compact && (!a->A || !a->B || !a->C)) || (a->compact && (!a->compact->vals || !a->compact->table)) || !a->status || a->N < 0 || a->n < 1 || a->n_lead < 0 || a->n_lead > a->n || (a->n_lead > 0 && !a->lead_idx)) {
        set_last_error("gecon_bk_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    return 0;
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_bk_count_batched(const gecon_bk_args* args, void* stream) {
    int rc = check_bk_args(args);
    if (rc) return rc;
    if (args->N == 0) return 0;
    const int np = round_up8(args->n + args->n_lead);
    GECON_DISPATCH_NP_WIDE(np, return launch_bk<NP_>(*args, (cudaStream_t)stream));
    return 0;
}

extern "C" int gecon_bk_count_host(const gecon_bk_args* a) {
    int rc = check_bk_args(a);
    if (rc) return rc;
    if (a->compact) {
        set_last_error("gecon_bk_count_host: compact Jacobians are a device-entry-point feature");
        return GECON_E_BADARG;
    }
    if (a->N == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, bm = N * n * n * 8;
    DevBuf dA, dB, dC, dL, dU, dS;
    gecon_bk_args d = *a;
    GECON_CUDA(dA.alloc(bm));
    GECON_CUDA(dB.alloc(bm));
    GECON_CUDA(dC.alloc(bm));
    GECON_CUDA(dL.alloc((size_t)a->n_lead * 4));
    GECON_CUDA(dS.alloc(N * 4));
    GECON_CUDA(cudaMemcpy(dA.p, a->A, bm, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dB.p, a->B, bm, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dC.p, a->C, bm, cudaMemcpyHostToDevice));
    if (a->n_lead) GECON_CUDA(cudaMemcpy(dL.p, a->lead_idx, (size_t)a->n_lead * 4, cudaMemcpyHostToDevice));
    if (a->accumulate) GECON_CUDA(cudaMemcpy(dS.p, a->status, N * 4, cudaMemcpyHostToDevice));
    d.A = dA.as<double>();
    d.B = dB.as<double>();
    d.C = dC.as<double>();
    d.lead_idx = dL.as<int32_t>();
    d.status = dS.as<int32_t>();
    if (a->n_unstable) {
        GECON_CUDA(dU.alloc(N * 4));
        if (a->accumulate) GECON_CUDA(cudaMemcpy(dU.p, a->n_unstable, N * 4, cudaMemcpyHostToDevice));  // skipped draws keep theirs
        d.n_unstable = dU.as<int32_t>();
    }
    rc = gecon_bk_count_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(a->status, d.status, N * 4, cudaMemcpyDeviceToHost));
    if (a->n_unstable) GECON_CUDA(cudaMemcpy(a->n_unstable, d.n_unstable, N * 4, cudaMemcpyDeviceToHost));
    return 0;
}
