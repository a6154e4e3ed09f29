// Batched cycle reduction + selection matrix + policy residual, one CTA per parameter draw.
//
// Restates gEconpy/solvers/cycle_reduction.py:127-183 (_cycle_reduction_core; authoritative flags),
// gEconpy/solvers/shared.py:74-75 (R = -(C T + B)^-1 D), gEconpy/model/statespace.py:213 (residual) and
// gEconpy/solvers/backward_looking.py:8-133 (C == NULL).  See DESIGN.md "cr_solve".
//
// Shared memory: 7 tiles (A0, A1, A2, A1hat, W, X0, X2).  Per iteration: copy A1->W, A0->X0, A2->X2; one
// Gauss-Jordan solve for [X0 | X2] = A1^-1 [A0 | A2]; four DMMA products; norms.  The non-zero column ranges of
// A0 (= lag columns) and A2 (= lead columns) are invariant under the iteration and are used to skip the zero blocks
// of the right-hand sides and of the products.
#include "common.cuh"
#include "linalg.cuh"

namespace gecon {

template <int NP>
__device__ __forceinline__ void acc_sub(Acc<NP>& a, const Acc<NP>& b) {
#pragma unroll
    for (int ct = 0; ct < NP / 8; ++ct) {
        a.v[ct][0] -= b.v[ct][0];
        a.v[ct][1] -= b.v[ct][1];
    }
}

// NP = 64: seven 34 KB tiles do not fit in 227 KB, so A1hat (touched once per iteration, by the warp that owns each
// strip) lives in an L2-resident global workspace, one tile per CTA, allocated stream-ordered by the launcher.
template <int NP>
constexpr bool cr_a1h_global() {
    return NP > 56 || NP == 48;
}
// NP = 48 (round 2): seven 20 KB tiles allow ONE six-warp CTA per SM; with A1hat and A2 (read as an A operand by two products, as a
// right-hand-side source by the first block step and by one norm per iteration; written once) in the per-CTA global workspace, five
// tiles = 100 KB and two CTAs share an SM, each hiding the other's barriers.
template <int NP>
constexpr bool cr_a2_global() {
    return NP == 48;
}

template <int NP>
struct CrSmem {
    static constexpr int TILES = 7 - (cr_a1h_global<NP>() ? 1 : 0) - (cr_a2_global<NP>() ? 1 : 0);
    static constexpr int WS_TILES = (cr_a1h_global<NP>() ? 1 : 0) + (cr_a2_global<NP>() ? 1 : 0);
    static constexpr size_t bytes = sizeof(double) * (TILES * Cfg<NP>::TILE + 8 * NP) + sizeof(int) * (4 * NP + 8);
};

// resident CTAs per SM that shared memory allows (7 tiles per CTA): the register allocator must not get in the way
template <int NP>
constexpr int cr_min_ctas() {
    // (NP = 16: 8 / 10 / 12 CTAs per SM measured 4.60 / 4.53 / 4.67 ms on the RBC workload)
    return NP <= 8 ? 16 : NP <= 16 ? 10 : NP <= 24 ? 5 : NP <= 32 ? 3 : NP <= 48 ? 2 : 1;
}

template <int NP>
__global__ void __launch_bounds__(Cfg<NP>::NT, cr_min_ctas<NP>()) cr_solve_kernel(const gecon_cr_args p, double* __restrict__ ws,
                                                                                  const gecon_compact_jac cj, double* __restrict__ scratch) {
    using C = Cfg<NP>;
    constexpr int LD = C::LD;
    extern __shared__ __align__(16) double sm[];
    double* A0 = sm;
    double* A1 = A0 + C::TILE;
    double* wsc = ws + (size_t)blockIdx.x * CrSmem<NP>::WS_TILES * C::TILE;  // this CTA's tiles in the L2-resident workspace
    double* A2 = cr_a2_global<NP>() ? wsc + (cr_a1h_global<NP>() ? C::TILE : 0) : A1 + C::TILE;
    double* nxt_sm = A1 + (cr_a2_global<NP>() ? 1 : 2) * C::TILE;
    double* A1h = cr_a1h_global<NP>() ? wsc : nxt_sm;
    double* W = nxt_sm + (cr_a1h_global<NP>() ? 0 : 1) * C::TILE;
    double* X0 = W + C::TILE;
    double* X2 = X0 + C::TILE;
    double* s_red = X2 + C::TILE;  // [8 NP]: two norm1_fast buffers
    int* s_piv = reinterpret_cast<int*>(s_red + 8 * NP);
    int* s_perm = s_piv + NP;
    int* s_lead = s_perm + NP;
    int* s_i = s_lead + NP;
    int* s_flag = s_i + 4;  // [NP + 1]
    for (int i = threadIdx.x; i < NP; i += C::NT) s_piv[i] = 0;

    const int n = p.n, k = p.k;
    const int no = (p.unperm && p.n_out > 0) ? p.n_out : n;  // rows/cols written out (sub-block gather when < n)
    if (p.unperm) {
        for (int i = threadIdx.x; i < no; i += C::NT) s_perm[i] = p.unperm[i];
    }
    const int* perm = p.unperm ? s_perm : nullptr;
    const int nl = p.lead_idx ? p.n_lead : 0;
    for (int i = threadIdx.x; i < nl; i += C::NT) s_lead[i] = p.lead_idx[i];

    for (long long draw = blockIdx.x; draw < p.N; draw += gridDim.x) {
        const double *gA, *gB, *gC, *gD;
        if (cj.vals) {  // compact Jacobian: expand this draw into the CTA's dense scratch (L2-resident) and read that
            double* sc = scratch + (size_t)blockIdx.x * ((size_t)3 * n * n + (size_t)n * k);
            expand_compact(sc, cj, draw, n, k, 0, 3);
            gA = sc;
            gB = sc + (size_t)n * n;
            gC = (cj.off[3] > cj.off[2]) ? sc + (size_t)2 * n * n : nullptr;  // no lead entries at all: backward-looking system
            gD = (p.R && cj.off[4] > cj.off[3]) ? sc + (size_t)3 * n * n : nullptr;
        } else {
            gA = p.A + (size_t)draw * n * n;
            gB = p.B + (size_t)draw * n * n;
            gC = p.C ? p.C + (size_t)draw * n * n : nullptr;
            gD = p.D ? p.D + (size_t)draw * n * k : nullptr;
        }
        if (!cj.vals) {   // the next draw of this CTA: pull its A, B, C (freshly written by the Jacobian kernel, i.e. in HBM) into L2 now, so
            // that its tile loads ~150 us from now do not wait on DRAM (12.5 % of the stall samples were those loads)
            const long long nxt = draw + gridDim.x;
            if (nxt < p.N) {
                const size_t bytes = (size_t)n * n * sizeof(double);
                for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)C::NT * 128) {
                    prefetch_l2(reinterpret_cast<const char*>(p.A + (size_t)nxt * n * n) + off);
                    prefetch_l2(reinterpret_cast<const char*>(p.B + (size_t)nxt * n * n) + off);
                    if (p.C) prefetch_l2(reinterpret_cast<const char*>(p.C + (size_t)nxt * n * n) + off);
                }
            }
        }

        tile_load<NP>(A0, gA, n, n, n);
        tile_load<NP>(A1, gB, n, n, n);
        tile_load<NP>(A1h, gB, n, n, n);
        if (gC) tile_load<NP>(A2, gC, n, n, n);
        else tile_zero<NP>(A2);
        __syncthreads();

        int lo0, hi0, lo2, hi2;
        nonzero_col_range<NP>(A0, n, s_i, lo0, hi0);
        nonzero_col_range<NP>(A2, n, s_i, lo2, hi2);
        // GEMM ranges: k in multiples of 4, output column tiles in multiples of 8
        const int k0lo = lo0 & ~3, k0hi = (hi0 + 3) & ~3, k2lo = lo2 & ~3, k2hi = (hi2 + 3) & ~3;
        const int c0lo = lo0 >> 3, c0hi = (hi0 + 7) >> 3, c2lo = lo2 >> 3, c2hi = (hi2 + 7) >> 3;

        int status = 0;
        bool converged = false;
        int it = 0;
        double a0n = 0.0, a2n = 0.0, a1n = 0.0;

        if (gC) {
            while (it < p.max_iter) {
                ++it;
                // [X0 | X2] = A1^-1 [A0 | A2], read from the live tiles, written to the scratch tiles; solution row j is
                // left in row s_piv[j], which the four products below read through
                const bool ok = gj_solve_blocked<NP>(A1, W, A0, X0, c0lo, c0hi, A2, X2, c2lo, c2hi, n, false, s_piv, s_flag);
                if (!ok) {  // LAPACK: singular U -> inf/NaN in getrs; the norm test below then stops the loop
                    tile_nanfill<NP>(X0, n, n);
                    tile_nanfill<NP>(X2, n, n);
                    __syncthreads();
                }
                {
                    Acc<NP> m20, t;
                    acc_zero(m20);
                    gemm_acc_bmap<NP>(m20, A2, X0, s_piv, 1.0, k2lo, k2hi, ok ? c0lo : 0, ok ? c0hi : C::CT);
                    acc_load<NP>(t, A1);
                    gemm_acc_bmap<NP>(t, A0, X2, s_piv, -1.0, k0lo, k0hi, ok ? c2lo : 0, ok ? c2hi : C::CT);
                    acc_sub(t, m20);
                    acc_store<NP>(t, A1);
                    acc_load<NP>(t, A1h);
                    acc_sub(t, m20);
                    acc_store<NP>(t, A1h);
                }
                Acc<NP> m00, m22;
                acc_zero(m00);
                acc_zero(m22);
                gemm_acc_bmap<NP>(m00, A0, X0, s_piv, -1.0, k0lo, k0hi, ok ? c0lo : 0, ok ? c0hi : C::CT);
                gemm_acc_bmap<NP>(m22, A2, X2, s_piv, -1.0, k2lo, k2hi, ok ? c2lo : 0, ok ? c2hi : C::CT);
                __syncthreads();  // every warp is done reading A0, A2
                acc_store<NP>(m00, A0);
                acc_store<NP>(m22, A2);
                __syncthreads();
                a0n = norm1_fast<NP>(A0, s_red);
                if (a0n < p.tol) {
                    if (p.scan_semantics) {  // the scan twin tests ||A0||_1 only (cycle_reduction.py:268-273)
                        converged = true;
                        break;
                    }
                    a2n = norm1_fast<NP>(A2, s_red + 4 * NP);
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
            if (!converged) {
                status |= GECON_ST_CR_NOT_CONVERGED;
                if (p.norms) {  // diagnostics of the numpy twin's failure tuple (cycle_reduction.py:101-109)
                    __syncthreads();  // norm1_fast has no trailing barrier
                    a2n = norm1<NP>(A2, n, s_red);
                    a1n = norm1<NP>(A1, n, s_red);
                }
            }
        }

        // ---- T = -A1hat^-1 A (cycle_reduction.py:181), or T = -B^-1 A for backward-looking models; 0 if not converged.
        // The iterated A0, A1, A2 are dead from here on: the A0 tile takes A again, A1 takes B, A2 takes C.
        tile_load<NP>(A0, gA, n, n, n);
        tile_zero<NP>(X0);
        __syncthreads();
        if (converged || !gC || p.scan_semantics) {
            const bool ok = gj_solve_blocked<NP>(A1h, W, A0, X0, c0lo, c0hi, nullptr, nullptr, 0, 0, n, true, s_piv, s_flag);  // backward looking: A1h == B
            if (!ok) {
                status |= GECON_ST_SINGULAR;
                tile_nanfill<NP>(X0, n, n);
            } else {
                for (int i = threadIdx.x; i < C::TILE; i += C::NT) X0[i] = -X0[i];
            }
        }
        double* Tt = X0;
        __syncthreads();

        // ---- CT = C T (A1hat tile), W = B + CT, R = -W^-1 D (shared.py:74-75)
        if (gC) tile_load<NP>(A2, gC, n, n, n);
        tile_load<NP>(A1, gB, n, n, n);
        __syncthreads();
        {
            Acc<NP> ct;
            acc_zero(ct);
            if (gC) gemm_acc<NP, false, false>(ct, A2, Tt, 1.0, (status & GECON_ST_SINGULAR) ? 0 : k2lo, (status & GECON_ST_SINGULAR) ? NP : k2hi);
            acc_store<NP>(ct, A1h);
        }
        __syncthreads();
        // One solve with W for both right-hand sides: D (-> R) and, for the Blanchard-Kahn certificate, the non-zero
        // columns of C (-> W^-1 C = -F, left in the A2 tile).
        const bool want_cert = p.lead_idx && gC && converged && !(status & (GECON_ST_SINGULAR | GECON_ST_CR_NAN));
        bool have_F = false;
        if ((gD && p.R) || want_cert) {
            for (int i = threadIdx.x; i < C::TILE; i += C::NT) W[i] = A1[i] + A1h[i];
            const int kd = (gD && p.R) ? k : 0;
            if (kd) tile_load<NP>(X2, gD, n, k, k);
            __syncthreads();
            const bool ok = gj_solve_blocked<NP>(W, W, X2, X2, 0, (kd + 7) >> 3, A2, A2, want_cert ? c2lo : 0, want_cert ? c2hi : 0, n, true, s_piv, s_flag);
            if (!ok) {
                status |= GECON_ST_SINGULAR;
                if (kd) tile_nanfill<NP>(X2, n, k);
                __syncthreads();
            }
            have_F = ok && want_cert;
            if (kd) tile_store<NP>(p.R + (size_t)draw * (p.r_stride ? (size_t)p.r_stride : (size_t)no * k), X2, no, k, k, -1.0, perm, nullptr);
        }

        // ---- residual sum((A + B T + (C T) T)^2) in solver order (statespace.py:213)
        {
            Acc<NP> e;
            acc_load<NP>(e, A0);
            gemm_acc<NP, false, false>(e, A1, Tt, 1.0);
            gemm_acc<NP, false, false>(e, A1h, Tt, 1.0);
            double ss = 0.0;
#pragma unroll
            for (int ct = 0; ct < C::CT; ++ct) ss += e.v[ct][0] * e.v[ct][0] + e.v[ct][1] * e.v[ct][1];
            const double resid = block_sum<NP>(ss, s_red);
            if (p.resid_tol > 0.0 && !(resid < p.resid_tol)) status |= GECON_ST_RESID;
            if (threadIdx.x == 0) {
                if (p.resid) p.resid[draw] = resid;
                if (p.n_iter) p.n_iter[draw] = it;
                if (p.norms) {
                    p.norms[3 * draw] = a0n;
                    p.norms[3 * draw + 1] = a2n;
                    p.norms[3 * draw + 2] = a1n;
                }
                p.status[draw] = p.accumulate ? (p.status[draw] | status) : status;
            }
        }
        tile_store<NP>(p.T + (size_t)draw * (p.t_stride ? (size_t)p.t_stride : (size_t)no * no), Tt, no, no, p.t_ld ? p.t_ld : no, 1.0, perm, perm);
        __syncthreads();

        // ---- Blanchard-Kahn certificate (see gecon_cr_args.lead_idx): rho(T) < 1 and rho(F_LL) < 1 by repeated squaring.
        // Z1 = the lag-column block of T (its other columns are zero), Z2 = (W^-1 C)[lead][:, lead]; ping-pong tiles.
        if (p.lead_idx) {
            bool certified = false;
            if (have_F) {
                const int s1 = hi0 - lo0;
                double* Z1 = W;
                double* Z1b = A1;
                double* Z2 = A0;
                double* Z2b = A1h;
                tile_zero<NP>(Z1);
                tile_zero<NP>(Z2);
                __syncthreads();
                for (int idx = threadIdx.x; idx < s1 * s1; idx += C::NT) {
                    const int a = idx / s1, b = idx - a * s1;
                    Z1[a * LD + b] = Tt[(lo0 + a) * LD + lo0 + b];
                }
                for (int idx = threadIdx.x; idx < nl * nl; idx += C::NT) {
                    const int a = idx / nl, b = idx - a * nl;
                    Z2[a * LD + b] = A2[s_lead[a] * LD + s_lead[b]];
                }
                __syncthreads();
                const int k1 = (s1 + 3) & ~3, c1 = (s1 + 7) >> 3, k2 = (nl + 3) & ~3, c2 = (nl + 7) >> 3;
                bool ok1 = (s1 == 0), ok2 = (nl == 0);
                for (int sq = 0; sq <= 14; ++sq) {
                    // the norms are only inspected after every second squaring (and at the start): a certificate
                    // found one squaring late costs one small product, a norm costs two barriers
                    const bool look = (sq == 0) || (sq & 1) == 0 || sq == 14;
                    const double n1 = (ok1 || !look) ? (ok1 ? 0.0 : 2.0) : norm1<NP>(Z1, s1, s_red);
                    const double n2 = (ok2 || !look) ? (ok2 ? 0.0 : 2.0) : norm1<NP>(Z2, nl, s_red);
                    ok1 = ok1 || (n1 < 1.0);
                    ok2 = ok2 || (n2 < 1.0);
                    if (ok1 && ok2) {
                        certified = true;
                        break;
                    }
                    if (!(n1 < 1e100) || !(n2 < 1e100) || sq == 14) break;  // growing powers / NaN: leave it to bk_count
                    Acc<NP> q1, q2;
                    acc_zero(q1);
                    acc_zero(q2);
                    const int wp = threadIdx.x >> 5;  // strips beyond the blocks hold nothing
                    if (!ok1 && wp < c1) gemm_acc<NP, false, false>(q1, Z1, Z1, 1.0, 0, k1, 0, c1);
                    if (!ok2 && wp < c2) gemm_acc<NP, false, false>(q2, Z2, Z2, 1.0, 0, k2, 0, c2);
                    if (!ok1 && wp < c1) acc_store<NP>(q1, Z1b, 0, c1);
                    if (!ok2 && wp < c2) acc_store<NP>(q2, Z2b, 0, c2);
                    __syncthreads();
                    if (!ok1) {
                        double* t = Z1;
                        Z1 = Z1b;
                        Z1b = t;
                    }
                    if (!ok2) {
                        double* t = Z2;
                        Z2 = Z2b;
                        Z2b = t;
                    }
                }
            }
            if (threadIdx.x == 0) {
                if (certified) p.status[draw] |= GECON_ST_BK_CERTIFIED;
                if (p.n_unstable) p.n_unstable[draw] = certified ? nl : -1;
            }
            __syncthreads();
        }

        // ---- residual norms of gEcon's state/jumper representation (perturbation.py:287-380), as solvability_check
        // reports them: entries of T, R below trunc_tol are zeroed first; state columns = columns of the truncated T
        // with a surviving entry;  nd = ||(A + B T~ + C T~ T~)[:, states]||_F,  ns = ||B R~ + C T~ R~ + D||_F.
        if (p.solv_norms) {
            const double tt = p.trunc_tol;
            for (int i = threadIdx.x; i < C::TILE; i += C::NT) {
                const double v = Tt[i];
                W[i] = (fabs(v) < tt) ? 0.0 : v;
            }
            if (gC) tile_load<NP>(A2, gC, n, n, n);
            else tile_zero<NP>(A2);
            tile_load<NP>(A1, gB, n, n, n);
            tile_load<NP>(A1h, gA, n, n, n);
            __syncthreads();
            if ((int)threadIdx.x < NP) {  // state-column flags of the truncated T
                int any = 0;
                if ((int)threadIdx.x < n)
                    for (int i = 0; i < n; ++i) any |= (W[i * LD + threadIdx.x] != 0.0);
                s_piv[threadIdx.x] = any;
            }
            {
                Acc<NP> ct;
                acc_zero(ct);
                gemm_acc<NP, false, false>(ct, A2, W, 1.0);
                acc_store<NP>(ct, A0);
            }
            __syncthreads();
            double nd, ns = 0.0;
            {
                Acc<NP> e;
                acc_load<NP>(e, A1h);
                gemm_acc<NP, false, false>(e, A1, W, 1.0);
                gemm_acc<NP, false, false>(e, A0, W, 1.0);
                const int q2 = 2 * (threadIdx.x & 3);
                double ss = 0.0;
#pragma unroll
                for (int ct = 0; ct < C::CT; ++ct) {
                    if (s_piv[ct * 8 + q2]) ss += e.v[ct][0] * e.v[ct][0];
                    if (s_piv[ct * 8 + q2 + 1]) ss += e.v[ct][1] * e.v[ct][1];
                }
                nd = sqrt(block_sum<NP>(ss, s_red));
            }
            if (gD && p.R) {
                for (int i = threadIdx.x; i < C::TILE; i += C::NT) {
                    const double v = -X2[i];  // X2 holds -R
                    A1h[i] = (fabs(v) < tt) ? 0.0 : v;
                }
                tile_load<NP>(A2, gD, n, k, k);
                __syncthreads();
                Acc<NP> e;
                acc_load<NP>(e, A2);
                gemm_acc<NP, false, false>(e, A1, A1h, 1.0);
                gemm_acc<NP, false, false>(e, A0, A1h, 1.0);
                double ss = 0.0;
#pragma unroll
                for (int ct = 0; ct < C::CT; ++ct) ss += e.v[ct][0] * e.v[ct][0] + e.v[ct][1] * e.v[ct][1];
                ns = sqrt(block_sum<NP>(ss, s_red));
            }
            if (threadIdx.x == 0) {
                p.solv_norms[2 * draw] = nd;
                p.solv_norms[2 * draw + 1] = ns;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------- host
static int check_cr_args(const gecon_cr_args* a) {
    if (!a || a->struct_size != sizeof(gecon_cr_args)) {
        set_last_error("gecon_cr_args: bad struct_size");
        return GECON_E_BADARG;
    }
    const bool cmp = a->compact != nullptr;
    if (cmp && (!a->compact->vals || !a->compact->table || a->compact->stride < a->compact->off[4] || a->compact->off[0] != 0)) {
        set_last_error("gecon_cr_args: malformed compact Jacobian descriptor");
        return GECON_E_BADARG;
    }
    if ((!cmp && (!a->A || !a->B)) || !a->T || !a->status || a->N < 0 || a->n < 1 || a->k < 0 || (!cmp && a->R && !a->D)) {
        set_last_error("gecon_cr_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    if (a->n_out < 0 || a->n_out > a->n || (a->n_out > 0 && !a->unperm)) {
        set_last_error("gecon_cr_args: n_out = %d needs 0 <= n_out <= n and an index list in unperm", a->n_out);
        return GECON_E_BADARG;
    }
    {
        const int64_t no = (a->unperm && a->n_out > 0) ? a->n_out : a->n;
        if ((a->t_ld && a->t_ld < no) || (a->t_stride && a->t_stride < no * (a->t_ld ? a->t_ld : no)) || (a->r_stride && a->r_stride < no * a->k)) {
            set_last_error("gecon_cr_args: t_ld / t_stride / r_stride smaller than the block they hold");
            return GECON_E_BADARG;
        }
    }
    if (a->lead_idx && (a->n_lead < 0 || a->n_lead > a->n)) {
        set_last_error("gecon_cr_args: n_lead = %d out of range", a->n_lead);
        return GECON_E_BADARG;
    }
    if (a->lag_lo < 0 || a->lag_hi < a->lag_lo || a->lag_hi > a->n || a->lead_lo < 0 || a->lead_hi < a->lead_lo || a->lead_hi > a->n) {
        set_last_error("gecon_cr_args: column ranges [%d, %d), [%d, %d) out of order or beyond n", a->lag_lo, a->lag_hi, a->lead_lo, a->lead_hi);
        return GECON_E_BADARG;
    }
    if (a->k > round_up8(a->n)) {
        set_last_error("gecon_cr_args: k = %d exceeds the padded state dimension", a->k);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

template <int NP>
static int launch_cr(const gecon_cr_args& a, cudaStream_t st) {
    int grid = 0;
    int rc = persistent_grid(cr_solve_kernel<NP>, Cfg<NP>::NT, CrSmem<NP>::bytes, a.N, &grid, nullptr, "GECON_CR_CTAS_PER_SM");
    if (rc) return rc;
    double *ws = nullptr, *scratch = nullptr;
    if (CrSmem<NP>::WS_TILES > 0) GECON_CUDA(cudaMallocAsync((void**)&ws, sizeof(double) * (size_t)grid * CrSmem<NP>::WS_TILES * Cfg<NP>::TILE, st));
    gecon_compact_jac cj{};
    if (a.compact) {
        cj = *a.compact;
        GECON_CUDA(cudaMallocAsync((void**)&scratch, sizeof(double) * (size_t)grid * ((size_t)3 * a.n * a.n + (size_t)a.n * a.k), st));
    }
    cr_solve_kernel<NP><<<grid, Cfg<NP>::NT, CrSmem<NP>::bytes, st>>>(a, ws, cj, scratch);
    g_launch_count++;
    const cudaError_t le = cudaGetLastError();
    if (ws) cudaFreeAsync(ws, st);
    if (scratch) cudaFreeAsync(scratch, st);
    GECON_CUDA(le);
    return 0;
}

int cr_kernel_info(int n, int* ctas, int* smem, int* threads) {
    const int np = round_up8(n);
    GECON_DISPATCH_NP(np, {
        int grid = 0;
        int rc = persistent_grid(cr_solve_kernel<NP_>, Cfg<NP_>::NT, CrSmem<NP_>::bytes, 1 << 30, &grid, ctas);
        if (rc) return rc;
        *smem = (int)CrSmem<NP_>::bytes;
        *threads = Cfg<NP_>::NT;
    });
    return 0;
}

}  // namespace gecon

using namespace gecon;

// One warp per draw (cr_warp.cuh) when the system is forward looking, fits a lane per row and the caller does not ask for
// the diagnostics only the CTA-per-draw kernel produces.  c_out = packed width in column tiles, rg = packed ranges.
// GECON_CR_KERNEL=cta|warp overrides the choice (warp: wherever it is eligible).
static bool cr_warp_eligible(const gecon_cr_args& a, cw_ranges* rg, int* c_out) {
    const int np = round_up8(a.n);
    const bool has_c = a.compact ? (a.compact->off[3] > a.compact->off[2]) : (a.C != nullptr);
    if (!has_c || np > 32 || a.solv_norms) return false;
    const char* e = getenv("GECON_CR_KERNEL");
    if (e && strcmp(e, "cta") == 0) return false;
    const bool hint = (a.lag_hi > a.lag_lo) || (a.lead_hi > a.lead_lo);
    rg->o0 = hint ? (a.lag_lo & ~1) : 0;
    rg->w0 = hint ? a.lag_hi - rg->o0 : a.n;
    rg->o2 = hint ? (a.lead_lo & ~1) : 0;
    rg->w2 = hint ? a.lead_hi - rg->o2 : a.n;
    const int kd = ((a.D || a.compact) && a.R) ? a.k : 0;
    const int nl = a.lead_idx ? a.n_lead : 0;
    int w = rg->w0 > rg->w2 ? rg->w0 : rg->w2;
    w = kd > w ? kd : w;
    w = nl > w ? nl : w;
    int c = (w + 7) / 8;
    if (c < 1) c = 1;
    if (8 * c > np) return false;
    *c_out = c;
    // measured (r02): 33.8 vs 52.8 ms per 262,144 draws at NP = 24, 26.1 vs 52.3 ms per 131,072 at NP = 32, 2.5 vs 4.5 ms per 65,536 at NP = 16
    return true;
}

static int launch_cr_warp(const gecon_cr_args& a, const cw_ranges& rg, int c, cudaStream_t st, int* info) {
    switch (round_up8(a.n)) {
        case 8: return cr_warp_launch_np8(a, rg, c, st, info);
        case 16: return cr_warp_launch_np16(a, rg, c, st, info);
        case 24: return cr_warp_launch_np24(a, rg, c, st, info);
        case 32: return cr_warp_launch_np32(a, rg, c, st, info);
        default: break;
    }
    set_last_error("cr_warp: padded dimension %d > 32", round_up8(a.n));
    return GECON_E_UNSUPPORTED_SIZE;
}

extern "C" int gecon_cr_check_args(const gecon_cr_args* args) { return check_cr_args(args); }

extern "C" int gecon_cr_solve_batched(const gecon_cr_args* args, void* stream) {
    int rc = check_cr_args(args);
    if (rc) return rc;
    if (args->N == 0) return 0;
    const int np = round_up8(args->n);
    {
        cw_ranges rg;
        int c = 0;
        if (cr_warp_eligible(*args, &rg, &c)) return launch_cr_warp(*args, rg, c, (cudaStream_t)stream, nullptr);
    }
    GECON_DISPATCH_NP(np, return launch_cr<NP_>(*args, (cudaStream_t)stream));
    return 0;
}

extern "C" int gecon_cr_solve_host(const gecon_cr_args* args) {
    int rc = check_cr_args(args);
    if (rc) return rc;
    if (args->N == 0) return 0;
    if (args->t_stride || args->r_stride || args->t_ld || args->compact) {
        set_last_error("gecon_cr_solve_host: strided outputs and compact Jacobians are device-entry-point features");
        return GECON_E_BADARG;
    }
    const size_t N = (size_t)args->N, n = args->n, k = args->k;
    const size_t no = (args->unperm && args->n_out > 0) ? (size_t)args->n_out : n;
    const size_t bm = N * n * n * sizeof(double), bd = N * n * k * sizeof(double);
    const size_t bmo = N * no * no * sizeof(double), bdo = N * no * k * sizeof(double);
    DevBuf dA, dB, dC, dD, dT, dR, dSt, dIt, dRes, dNo, dPerm, dLead, dNu, dSn;
    gecon_cr_args d = *args;
    GECON_CUDA(dA.alloc(bm));
    GECON_CUDA(dB.alloc(bm));
    GECON_CUDA(dT.alloc(bmo));
    GECON_CUDA(dSt.alloc(N * sizeof(int32_t)));
    GECON_CUDA(cudaMemcpy(dA.p, args->A, bm, cudaMemcpyHostToDevice));
    GECON_CUDA(cudaMemcpy(dB.p, args->B, bm, cudaMemcpyHostToDevice));
    if (args->accumulate) GECON_CUDA(cudaMemcpy(dSt.p, args->status, N * sizeof(int32_t), cudaMemcpyHostToDevice));
    d.A = dA.as<double>();
    d.B = dB.as<double>();
    d.T = dT.as<double>();
    d.status = dSt.as<int32_t>();
    if (args->C) {
        GECON_CUDA(dC.alloc(bm));
        GECON_CUDA(cudaMemcpy(dC.p, args->C, bm, cudaMemcpyHostToDevice));
        d.C = dC.as<double>();
    }
    if (args->D) {
        GECON_CUDA(dD.alloc(bd));
        GECON_CUDA(cudaMemcpy(dD.p, args->D, bd, cudaMemcpyHostToDevice));
        d.D = dD.as<double>();
    }
    if (args->R) {
        GECON_CUDA(dR.alloc(bdo));
        d.R = dR.as<double>();
    }
    if (args->n_iter) {
        GECON_CUDA(dIt.alloc(N * sizeof(int32_t)));
        d.n_iter = dIt.as<int32_t>();
    }
    if (args->resid) {
        GECON_CUDA(dRes.alloc(N * sizeof(double)));
        d.resid = dRes.as<double>();
    }
    if (args->norms) {
        GECON_CUDA(dNo.alloc(3 * N * sizeof(double)));
        d.norms = dNo.as<double>();
    }
    if (args->unperm) {
        GECON_CUDA(dPerm.alloc(no * sizeof(int32_t)));
        GECON_CUDA(cudaMemcpy(dPerm.p, args->unperm, no * sizeof(int32_t), cudaMemcpyHostToDevice));
        d.unperm = dPerm.as<int32_t>();
    }
    if (args->lead_idx) {
        GECON_CUDA(dLead.alloc((size_t)args->n_lead * sizeof(int32_t)));
        GECON_CUDA(cudaMemcpy(dLead.p, args->lead_idx, (size_t)args->n_lead * sizeof(int32_t), cudaMemcpyHostToDevice));
        d.lead_idx = dLead.as<int32_t>();
    }
    if (args->n_unstable) {
        GECON_CUDA(dNu.alloc(N * sizeof(int32_t)));
        d.n_unstable = dNu.as<int32_t>();
    }
    if (args->solv_norms) {
        GECON_CUDA(dSn.alloc(2 * N * sizeof(double)));
        d.solv_norms = dSn.as<double>();
    }
    rc = gecon_cr_solve_batched(&d, nullptr);
    if (rc) return rc;
    if (args->solv_norms) GECON_CUDA(cudaMemcpy(args->solv_norms, d.solv_norms, 2 * N * sizeof(double), cudaMemcpyDeviceToHost));
    if (args->n_unstable) GECON_CUDA(cudaMemcpy(args->n_unstable, d.n_unstable, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(args->T, d.T, bmo, cudaMemcpyDeviceToHost));
    if (args->R) GECON_CUDA(cudaMemcpy(args->R, d.R, bdo, cudaMemcpyDeviceToHost));
    GECON_CUDA(cudaMemcpy(args->status, d.status, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (args->n_iter) GECON_CUDA(cudaMemcpy(args->n_iter, d.n_iter, N * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (args->resid) GECON_CUDA(cudaMemcpy(args->resid, d.resid, N * sizeof(double), cudaMemcpyDeviceToHost));
    if (args->norms) GECON_CUDA(cudaMemcpy(args->norms, d.norms, 3 * N * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}
