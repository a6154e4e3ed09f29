/* gecon_b200.h -- C ABI of libgecon_b200.so: the B200 (sm_100a, fp64) implementation of gEconpy's per-draw
 * estimation hot path, batched over parameter draws.
 *
 * Reference interfaces replaced (paths relative to the gEconpy source tree):
 *   gecon_cr_solve_*      gEconpy/solvers/cycle_reduction.py:127-183 (_cycle_reduction_core), :23-114 (numpy twin),
 *                         gEconpy/solvers/shared.py:74-75 (R = -(C T + B)^-1 D), gEconpy/model/statespace.py:213
 *                         (policy residual), gEconpy/solvers/backward_looking.py:8-133 (C == 0 special case)
 *   gecon_bk_count_*      gEconpy/model/perturbation.py:448-505,586-625 (check_bk_condition_pt),
 *                         gEconpy/solvers/gensys.py:568-614 (pencil assembly), gEconpy/pytensorf/real_eig.py:31-36
 *   gecon_dlyap_*         pt.linalg.solve_discrete_lyapunov call at gEconpy/model/statespace.py:814-815
 *   gecon_kalman_ll_*     PyMCStateSpace.build_statespace_graph call at gEconpy/model/statespace.py:1151-1157
 *                         (pymc_extras StandardFilter), with Q/H/Z/d as built at statespace.py:240-296,334-388,800-812
 *   gecon_solve_*         the np.linalg.solve / _solve_gen call sites (cycle_reduction.py:181,396; backward_looking.py)
 *
 * Conventions
 *   - All matrices are IEEE fp64, C-contiguous (row-major), batched on a leading draw axis:
 *     A[N][n][n], D[N][n][k], T[N][n][n], R[N][n][k].  Inputs are never written.
 *   - Entry points ending in `_batched` take DEVICE pointers and a CUDA stream (a cudaStream_t passed as
 *     void*; NULL = default stream) and return immediately after the launch.  Entry points ending in `_host` take
 *     HOST pointers, copy in, launch, copy out and synchronise.
 *   - The return value is 0 on success, a negative GECON_E_* for bad arguments, or a positive cudaError_t.
 *     gecon_get_last_error() returns a message for the calling thread's last failure.
 *   - Numerical failure is never an error code: it is reported per draw in `status` (bit field below), exactly as
 *     the reference reports it through flags / NaN fills / -inf potentials.
 *   - Supported sizes: 1 <= n <= 64 (Blanchard-Kahn pencils and the general solve up to 88; the gradient path's filter up to 48),
 *     0 <= k <= n, 1 <= p <= 8.
 */
#ifndef GECON_B200_H
#define GECON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GECON_ABI_VERSION 3

/* argument errors */
#define GECON_E_BADARG (-1)
#define GECON_E_UNSUPPORTED_SIZE (-2)
#define GECON_E_NO_DEVICE (-3)

/* per-draw status bits (0 = everything fine) */
#define GECON_ST_CR_NOT_CONVERGED 0x001  /* cycle reduction hit max_iter or a NaN norm: T = 0 (cycle_reduction.py:181-183) */
#define GECON_ST_CR_NAN 0x002            /* the A0 norm became NaN */
#define GECON_ST_SINGULAR 0x004          /* a linear solve met a zero / non-finite pivot: output NaN-filled */
#define GECON_ST_RESID 0x008             /* sum((A + B T + C T T)^2) >= resid_tol (statespace.py:1210-1215) */
#define GECON_ST_BK 0x010                /* n_unstable != n_forward (perturbation.py:612-625) */
#define GECON_ST_BK_INCONCLUSIVE 0x020   /* eigenvalue count could not be resolved (eigenvalue on the unit circle) */
#define GECON_ST_LYAP 0x040              /* doubling iteration for P0 did not converge (rho(T) >= 1) */
#define GECON_ST_NOT_PD 0x080            /* innovation covariance F_t not positive definite at some t */
#define GECON_ST_LL_NONFINITE 0x100      /* log-likelihood is NaN or +-inf */
#define GECON_ST_JAC_NONFINITE 0x200     /* steady state / Jacobian evaluation produced a non-finite entry */
#define GECON_ST_SKIPPED 0x400           /* draw skipped because status_in & gate_mask != 0 */
#define GECON_ST_BK_CERTIFIED 0x800      /* informational: n_unstable == n_forward proved by the solver kernel */

/* ---------------------------------------------------------------------------------------------------------------
 * Cycle reduction:  A + B T + C T T = 0,  R = -(C T + B)^-1 D,  resid = sum((A + B T + C T T)^2).
 * A = d/dx_{t-1}, B = d/dx_t, C = d/dx_{t+1}, D = d/deps  (the REFERENCE's naming, perturbation.py:42-46).
 * Follows _cycle_reduction_core: T = 0 and GECON_ST_CR_NOT_CONVERGED unless ||A0||_1 < tol and ||A2||_1 < tol
 * within max_iter iterations.  If C is NULL the system is backward looking: T = -B^-1 A, R = -B^-1 D.
 * ------------------------------------------------------------------------------------------------------------- */
/* Compact Jacobian: the structural non-zeros of A, B, C, D only (a per-model generated kernel writes them, see the end of this
 * file), so that the dense matrices never exist in HBM: medium NK 112 doubles per draw instead of 1,824.  When a solver entry
 * point is given one, its A, B, C, D pointers are ignored. */
typedef struct gecon_compact_jac {
    const double* vals;   /* [N][stride] device: entry e of draw i at vals[i * stride + e] */
    int64_t stride;       /* doubles between consecutive draws (>= nnz) */
    const int32_t* table; /* [nnz] device: row << 16 | col of entry e (rows / columns in solver order) */
    int32_t off[5];       /* entries of A are [off[0], off[1]), of B [off[1], off[2]), of C [off[2], off[3]), of D [off[3], off[4]) */
    int32_t reserved;
} gecon_compact_jac;

typedef struct gecon_cr_args {
    size_t struct_size;  /* sizeof(gecon_cr_args) */
    const double* A;     /* [N][n][n] */
    const double* B;     /* [N][n][n] */
    const double* C;     /* [N][n][n] or NULL */
    const double* D;     /* [N][n][k] or NULL (then R is not computed) */
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t max_iter;
    int32_t accumulate;     /* 0: status is overwritten, else new bits are OR-ed into the existing value */
    double tol;
    double resid_tol;       /* sets GECON_ST_RESID when resid >= resid_tol or NaN; <= 0 disables */
    const int32_t* unperm;  /* [n] or NULL: outputs are T[unperm][:, unperm], R[unperm] (statespace.py:217-220) */
    double* T;              /* [N][n][n] out */
    double* R;              /* [N][n][k] out or NULL */
    int32_t* status;        /* [N] out (overwritten) */
    int32_t* n_iter;        /* [N] out or NULL: iterations executed */
    double* resid;          /* [N] out or NULL */
    double* norms;          /* [N][3] out or NULL: final ||A0||_1, ||A2||_1, ||A1||_1 (the last two are always
                               evaluated when the iteration did not converge) */
    int32_t n_out;          /* 0: outputs are n x n / n x k.  > 0 (needs unperm): unperm has n_out <= n entries and the
                               outputs are the SUB-BLOCKS T[unperm][:, unperm] (n_out x n_out), R[unperm] (n_out x k).
                               Used to hand the Kalman kernel only the variables the likelihood depends on (lagged
                               variables + observed variables): an exact restriction, see DESIGN.md */
    int32_t n_lead;         /* entries of lead_idx */
    const int32_t* lead_idx; /* [n_lead] or NULL.  When given, converged draws also get a Blanchard-Kahn CERTIFICATE:
                               with W = C T + B the pencil's finite spectrum is eig(T) U {1/mu : mu in eig(F)},
                               F = -W^-1 C restricted to the lead columns, so rho(T) < 1 and rho(F) < 1 (each proved by
                               a power of the matrix having 1-norm < 1, repeated squaring) imply n_unstable == n_lead.
                               Certified draws get GECON_ST_BK_CERTIFIED; the others are left to gecon_bk_count_* */
    int32_t* n_unstable;    /* [N] out or NULL: n_lead for certified draws, -1 otherwise */
    double* solv_norms;     /* [N][2] out or NULL: (norm_deterministic, norm_stochastic) of solvability_check
                               (statistics/perturbation_diagnostics.py:105-161, model/perturbation.py:287-380) */
    double trunc_tol;       /* entries of T, R below this are zeroed before the norms (the reference's `tol`) */
    /* Strided outputs (device entry point only; 0 = dense).  They let the solver write T, R straight into the top-left
     * block of a pre-initialised AUGMENTED transition / selection pair [[T, 0], [F, C]], [[R], [0]] (cumulator and
     * observation-lag states of gEconpy/model/statespace.py:598-723): the constant rows are filled once and never
     * touched, so augmentation costs no extra pass over HBM. */
    int64_t t_stride;       /* doubles between consecutive draws of T (0: n_out * n_out) */
    int64_t r_stride;       /* doubles between consecutive draws of R (0: n_out * k) */
    int32_t t_ld;           /* leading dimension of a draw's T (0: n_out) */
    /* Structural column ranges (all four 0: none given).  A = d/dx_{t-1} has non-zero entries only in the columns of lagged
     * variables and C = d/dx_{t+1} only in those of lead variables, and both sets are contiguous in the reference's solver
     * order (gEconpy/model/perturbation.py:130-158, model/model.py:199-250).  When lag_hi > lag_lo or lead_hi > lead_lo the
     * caller PROMISES that A (C) is zero outside columns [lag_lo, lag_hi) ([lead_lo, lead_hi)): entries outside are not
     * read, and the one-warp-per-draw kernel keeps A0, A2, X0, X2 packed to those ranges (csrc/cr_warp.cuh). */
    int32_t lag_lo;
    int32_t lag_hi;
    int32_t lead_lo;
    int32_t lead_hi;
    int32_t scan_semantics; /* 0: _cycle_reduction_core (converged = ||A0||_1 < tol AND ||A2||_1 < tol; T = 0 otherwise).
                               1: the scan twin (cycle_reduction.py:246-294): the iteration stops as soon as ||A0||_1 < tol (A2 is
                               not tested), T = -A1hat^-1 A is ALWAYS computed (no zeroing), n_iter = the steps actually taken
                               (the `n_steps` output of scan_cycle_reduction); GECON_ST_CR_NOT_CONVERGED then only says that
                               max_iter passed without ||A0||_1 < tol */
    const gecon_compact_jac* compact; /* HOST pointer to the descriptor, or NULL.  Device entry point only */
} gecon_cr_args;

int gecon_cr_solve_batched(const gecon_cr_args* args, void* stream);
int gecon_cr_solve_host(const gecon_cr_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Blanchard-Kahn count on the regularised Sims pencil of check_bk_condition_pt:
 *   Gamma0 = [[B, C], [-I, 0]], Gamma1 = [[A, 0], [0, I]], rows/cols {0..n-1} U {n + lead_idx[j]},
 *   M = (-Gamma0_sel + 1e-8 I)^-1 Gamma1_sel,  n_unstable = #{|eig(M)| > 1},  ok = (n_unstable == n_lead).
 * The count is obtained without forming eigenvalues: #{|lambda| > 1} = (m + trace sign(N)) / 2 with
 * N = (Gamma1_sel - G)(Gamma1_sel + G)^-1, G = -Gamma0_sel + 1e-8 I  (Cayley transform + matrix sign function).
 * Sets GECON_ST_BK / GECON_ST_BK_INCONCLUSIVE in status (OR-ed into the existing value when accumulate != 0).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gecon_bk_args {
    size_t struct_size;
    const double* A;
    const double* B;
    const double* C;
    int64_t N;
    int32_t n;
    int32_t n_lead;
    const int32_t* lead_idx; /* [n_lead] column positions (in the order of A,B,C) of the structural lead variables */
    int32_t accumulate;      /* 0: status is overwritten, else OR-ed */
    int32_t max_iter;        /* Newton iterations of the sign function (<= 0: default 60) */
    int32_t* n_unstable;     /* [N] out or NULL (-1 when inconclusive) */
    int32_t* status;         /* [N] in/out */
    int32_t skip_mask;       /* with accumulate != 0: draws with status & skip_mask are left untouched (e.g.
                                GECON_ST_BK_CERTIFIED | GECON_ST_JAC_NONFINITE) */
    int32_t reserved0;
    const gecon_compact_jac* compact; /* HOST pointer to the descriptor, or NULL (then A, B, C are read).  Device entry point only */
} gecon_bk_args;

int gecon_bk_count_batched(const gecon_bk_args* args, void* stream);
int gecon_bk_count_host(const gecon_bk_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Discrete Lyapunov equation P = T P T' + R diag(q) R' by Smith doubling.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gecon_dlyap_args {
    size_t struct_size;
    const double* T;     /* [N][n][n] */
    const double* R;     /* [N][n][k] */
    const double* qdiag; /* shock variances: [N][k] (q_stride = k) or [k] shared (q_stride = 0) */
    int64_t q_stride;
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t max_iter; /* <= 0: default 64 */
    int32_t accumulate;
    double* P;        /* [N][n][n] out */
    int32_t* status;  /* [N] */
    int32_t* n_iter;  /* [N] out or NULL */
} gecon_dlyap_args;

int gecon_dlyap_batched(const gecon_dlyap_args* args, void* stream);
int gecon_dlyap_host(const gecon_dlyap_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Kalman-filter log-likelihood (pymc_extras StandardFilter semantics: update -> jitter -> predict, Joseph form,
 * missing observations masked out of Z and H, a0 = 0, P0 = dlyap(T, R Q R') unless P0 is given):
 *   x_t = T x_{t-1} + R eps_t, eps ~ N(0, diag(q));   y_t = d + Z x_t + eta_t, eta ~ N(0, diag(h)).
 * Z is either dense (p x n, shared by all draws or one per draw: z_stride) or a selector given by obs_idx
 * (Z[a][obs_idx[a]] = 1).
 * Kernels behind it (same arithmetic; the Joseph update in its rank-p form P - K (P Z' + jitter K)'): one THREAD per draw for a
 * selector Z with n <= 4, p <= 2; one WARP per draw for a selector Z with n <= 31 (n <= 23 when T_obs p is too large for four
 * warps' tiles next to Y); one CTA per draw otherwise (dense Z, full shock covariance, n <= 64).  p <= 8.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gecon_kalman_args {
    size_t struct_size;
    const double* T;     /* [N][n][n] */
    const double* R;     /* [N][n][k] */
    const double* qdiag; /* shock variances, [N][k] or shared [k] */
    int64_t q_stride;    /* k or 0 */
    const double* hdiag; /* measurement-error variances, [N][p] or shared [p]; NULL = 0 */
    int64_t h_stride;    /* p or 0 */
    const double* Z;     /* [p][n] shared or [N][p][n] (z_stride = p n), or NULL when obs_idx is given */
    const int32_t* obs_idx; /* [p] or NULL */
    const double* d;     /* observation intercept: [N][p] (d_stride = p), shared [p] (0), or NULL = 0 */
    int64_t d_stride;
    const double* Y;     /* [Tobs][p] shared by all draws; NaN or missing_fill marks a missing entry */
    const double* P0;    /* [N][n][n] or NULL (computed in-kernel by doubling) */
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t p;
    int32_t Tobs;
    double jitter;        /* cov_jitter, reference default 1e-8 */
    double missing_fill;  /* reference default -9999.0 */
    int32_t mvn_const_mode; /* 0: -0.5 * (p log 2pi + ...) per step; 1: bare log 2pi (older pymc_extras) */
    int32_t lyap_max_iter;  /* <= 0: default 64 */
    const int32_t* status_in; /* [N] or NULL */
    int32_t gate_mask;    /* draws with status_in & gate_mask get ll = -inf and are skipped (the reference's
                             pm.Potential(-inf) gates, statespace.py:1206-1215) */
    int32_t sigma_inputs; /* 0: qdiag/hdiag hold variances; 1: they hold standard deviations (sigma_<shock>,
                             error_sigma_<state>: statespace.py:255-258,800-810) and are squared in the kernel */
    double* ll;           /* [N] out */
    int32_t* status;      /* [N] out: status_in | new bits (may alias status_in) */
    double* ll_t;         /* [N][Tobs] out or NULL: per-observation log-likelihood */
    int64_t z_stride;     /* 0: Z is shared by all draws; p * n: Z is [N][p][n], one design matrix per draw (parameter-
                             dependent observation equations, statespace.py:299-331) */
    const double* qfull;  /* full shock covariance Q ("state_cov" of full_shock_covariance=True, statespace.py:245-249):
                             [N][k][k] (qfull_stride = k k) or shared [k][k] (0); when given, qdiag is ignored (may be NULL)
                             and the state dimension runs on the CTA-per-draw kernel */
    int64_t qfull_stride;
    int32_t h_count;      /* 0: hdiag holds p entries per draw.  > 0: only the first h_count entries are read, the others are 0
                             (lets hdiag point INTO a wider parameter vector: error variances of the first h_count observables) */
    int32_t t_cols;       /* 0: T is dense.  > 0: the caller's promise that only the first t_cols columns of T can be non-zero (T = -A1hat^-1 A
                             has non-zero columns only at the lagged variables; BatchedStateSpace orders the filter variables [lagged |
                             observed only]): the warp-per-draw kernel then runs the k-loops of T [P+ | a+] and W T' over those columns only.
                             The columns beyond MUST hold zeros (the set-up, P0 = dlyap(T, R Q R'), still reads all of T).  The other
                             kernels ignore it */
    int32_t mask_intercept; /* 0 (upstream StandardFilter as restated there): the intercept d is NOT masked at missing
                               entries, v_i = 0 - d_i there.  1: missing entries have v_i = 0 (d masked like Z and H), the
                               convention of a filter that drops missing rows.  tests/golden/make_kalman_goldens.py records which
                               one the installed pymc_extras follows */
    int32_t reserved2;
} gecon_kalman_args;

int gecon_kalman_ll_batched(const gecon_kalman_args* args, void* stream);
int gecon_kalman_ll_host(const gecon_kalman_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched general solve X = M^-1 RHS with partial pivoting (NaN-filled + GECON_ST_SINGULAR on failure),
 * and a batched product C = alpha * op(A) * op(B) (exercises the in-CTA DMMA GEMM; used by tests).
 * ------------------------------------------------------------------------------------------------------------- */
int gecon_solve_batched(const double* M, const double* RHS, int64_t N, int32_t n, int32_t m, double* X, int32_t* status,
                        void* stream);
int gecon_solve_host(const double* M, const double* RHS, int64_t N, int32_t n, int32_t m, double* X, int32_t* status);
int gecon_gemm_batched(const double* A, const double* B, int64_t N, int32_t n, int32_t trans_a, int32_t trans_b, double alpha,
                       double* C, void* stream);
int gecon_gemm_host(const double* A, const double* B, int64_t N, int32_t n, int32_t trans_a, int32_t trans_b, double alpha,
                    double* C);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched linear propagation X_t = T X_{t-1} + R E_t, t = 0..L-1, X an n x m panel per draw:
 *   simulate            gEconpy/model/simulate.py:171-183 (_simulate_linear_system): X0 = NULL, E = shock trajectories,
 *                       m = number of trajectories; out[0] = R e_0
 *   impulse responses   gEconpy/model/simulate.py:201-318: m = k, E_0 = diag(shock sizes), E_t = 0 afterwards
 *   autocovariances     gEconpy/model/statistics/covariance.py:133-161 (_compute_autocovariance_matrix):
 *                       X0 = Sigma, start_at_x0 = 1, R = E = NULL; out[h] = T^h Sigma
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gecon_propagate_args {
    size_t struct_size;
    const double* T;   /* [N][n][n] */
    const double* R;   /* [N][n][k] or NULL (no shocks) */
    const double* X0;  /* [N][n][m] state before t = 0, or NULL (zero) */
    const double* E;   /* shocks: [L][k][m] shared by all draws (e_stride = 0) or [N][L][k][m] (e_stride = L k m); NULL = none */
    int64_t e_stride;
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t m;
    int32_t L;
    int32_t start_at_x0; /* 1: out[0] = X0 (+ R E_0) instead of T X0 (+ R E_0) */
    int32_t reserved0;
    double* out;       /* [N][L][n][m] */
} gecon_propagate_args;

int gecon_propagate_batched(const gecon_propagate_args* args, void* stream);
int gecon_propagate_host(const gecon_propagate_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Gradient path.
 *
 * gecon_kalman_grad_*: the log-likelihood of gecon_kalman_ll_* (same arguments, Z shared or a selector) AND its gradient
 * with respect to T, R, the shock / measurement-error scales and the observation intercept, by a reverse sweep over the
 * stored predicted moments (P0 = dlyap(T, R Q R') is differentiated through a second doubling).  Replaces the graph
 * pytensor differentiates behind PyMCStateSpace.build_statespace_graph (gEconpy/model/statespace.py:812-820,1151-1157).
 * With sigma_inputs != 0, q_bar / h_bar are derivatives with respect to the standard deviations.
 * Sizes: n <= 48, k <= n, p <= 8.  Gated draws (status_in & gate_mask) get ll = -inf and zero gradients.
 * Shock covariance: diagonal (qdiag -> q_bar) or full (qfull -> qfull_bar; full_shock_covariance=True, statespace.py:245-249).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct gecon_kalman_grad_args {
    size_t struct_size;
    const double* T;
    const double* R;
    const double* qdiag;
    int64_t q_stride;
    const double* hdiag;
    int64_t h_stride;
    const double* Z;         /* [p][n] shared or [N][p][n] (z_stride), or NULL when obs_idx is given */
    const int32_t* obs_idx;
    const double* d;
    int64_t d_stride;
    const double* Y;
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t p;
    int32_t Tobs;
    double jitter;
    double missing_fill;
    int32_t mvn_const_mode;
    int32_t lyap_max_iter;
    const int32_t* status_in;
    int32_t gate_mask;
    int32_t sigma_inputs;
    double* ll;       /* [N] out */
    int32_t* status;  /* [N] out */
    double* T_bar;    /* [N][n][n] out: dll/dT */
    double* R_bar;    /* [N][n][k] out: dll/dR */
    double* q_bar;    /* [N][k] out: dll/dq (variances) or dll/dsigma */
    double* h_bar;    /* [N][p] out or NULL */
    double* d_bar;    /* [N][p] out or NULL */
    int64_t z_stride; /* 0: Z shared by all draws; p * n: one design matrix per draw */
    double* Z_bar;    /* [N][p][n] out or NULL: dll/dZ (dense design matrices; the adjoint of observation equations) */
    int32_t mask_intercept; /* as in gecon_kalman_args */
    int32_t reserved2;
    const double* qfull;    /* full shock covariance Q, as in gecon_kalman_args ([N][k][k] with qfull_stride = k k, or shared [k][k]
                               with 0); when given, qdiag / q_bar are ignored (may be NULL) and the derivative goes to qfull_bar */
    int64_t qfull_stride;
    double* qfull_bar;      /* [N][k][k] out: dll/dQ, symmetrised ((G + G') / 2, G = R' C0_bar R): Q is a covariance, only the
                               symmetric part of its derivative is defined */
} gecon_kalman_grad_args;

int gecon_kalman_grad_batched(const gecon_kalman_grad_args* args, void* stream);
int gecon_kalman_grad_host(const gecon_kalman_grad_args* args);

/* gecon_policy_adjoint_*: reverse mode of the perturbation solution.  Given T_bar (and optionally R_bar) returns the
 * adjoints of A, B, C (and D):
 *   o1_policy_function_adjoints(A, B, C, T, T_bar) -> [A_bar, B_bar, C_bar]   gEconpy/solvers/shared.py:12-71
 *   (the pullback of CycleReductionWrapper / GensysWrapper / scan_cycle_reduction, cycle_reduction.py:212-213,
 *   gensys.py:668-676), plus R = -(C T + B)^-1 D (shared.py:74-75) in reverse when R_bar, R and D are given.
 * The reference solves an n^2 x n^2 Kronecker system for the multipliers S; here W' S + C' S T' = -T_bar (W = C T + B)
 * is solved as the Stein equation S = Q + G S T' by doubling.  Singular W: NaN outputs + GECON_ST_SINGULAR.  n <= 64. */
typedef struct gecon_policy_adjoint_args {
    size_t struct_size;
    const double* A;      /* [N][n][n] (not read: A_bar = S does not depend on A; kept for the reference's signature) */
    const double* B;
    const double* C;
    const double* D;      /* [N][n][k] or NULL */
    const double* T;
    const double* R;      /* [N][n][k] or NULL */
    const double* T_bar;  /* [N][n][n] */
    const double* R_bar;  /* [N][n][k] or NULL */
    int64_t N;
    int32_t n;
    int32_t k;
    int32_t max_iter;     /* doubling steps, <= 0: 64 */
    int32_t reserved0;
    double* A_bar;
    double* B_bar;
    double* C_bar;
    double* D_bar;        /* [N][n][k] or NULL */
    int32_t* status;      /* [N] or NULL */
} gecon_policy_adjoint_args;

int gecon_policy_adjoint_batched(const gecon_policy_adjoint_args* args, void* stream);
int gecon_policy_adjoint_host(const gecon_policy_adjoint_args* args);

/* ---------------------------------------------------------------------------------------------------------------
 * Eigenvalues of batched real general matrices M[N][m][m] (balancing, Householder Hessenberg reduction, Francis double-shift
 * QR; one warp per matrix, m <= 160).  Replaces the numpy.linalg.eig call of RealEig.perform (gEconpy/pytensorf/real_eig.py:
 * 29-36) and, applied to M = (-Gamma0_sel + 1e-8 I)^-1 Gamma1_sel, compute_bk_eigenvalues_pt (gEconpy/model/perturbation.py:
 * 448-505).  re / im [N][m] come out in deflation order (complex pairs adjacent, positive imaginary part first); callers
 * sort by modulus as the reference does.  status: 0, GECON_ST_LL_NONFINITE (non-finite input) or GECON_ST_BK_INCONCLUSIVE
 * (the QR iteration did not converge); outputs are NaN-filled then.  balance != 0: scale by powers of two first (as dgeev).
 * ------------------------------------------------------------------------------------------------------------- */
int gecon_real_eig_batched(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status,
                           void* stream);
int gecon_real_eig_host(const double* M, int64_t N, int32_t m, int32_t balance, double* re, double* im, int32_t* status);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused theta -> log-likelihood (`gecon_model_<hash>_loglik`; the per-draw path of SURVEY 3.3,
 * gEconpy/model/statespace.py:725-820, 1139-1215): per chunk of draws, on the caller's stream, the generated model kernel
 * writes the COMPACT Jacobian, the solver kernels consume it, the filter reads the shock / measurement-error scales in
 * place from the parameter vector.  No dense A, B, C, D, no memsets, no copy kernels; workspace from the stream-ordered
 * allocator.  Plain state spaces only: selector observation matrix, diagonal shock covariance, no state augmentation, no
 * observation equations, no steady-state intercept (the Python pipeline handles those).
 * Every generated model library exports
 *     int gecon_model_loglik(gecon_pipeline_args* args, void* stream);
 * which fills in `jacobian`, `nz_table`, `nz_off`, `nnz`, `n`, `k`, `n_theta`, `col_ranges`, `cr_solve` (and `lead_idx` /
 * `n_lead` when they are NULL / 0 and check_bk != 0) from its own tables and calls gecon_loglik_pipeline.
 * ------------------------------------------------------------------------------------------------------------- */
typedef int (*gecon_jacobian_compact_fn)(const double* theta, int64_t theta_stride, int64_t N, double* vals, double* xss,
                                         int32_t* status, void* stream);

typedef int (*gecon_cr_solve_fn)(const gecon_cr_args* args, void* stream); /* the contract of gecon_cr_solve_batched */
typedef int (*gecon_kalman_ll_fn)(const gecon_kalman_args* args, void* stream); /* the contract of gecon_kalman_ll_batched */

typedef struct gecon_pipeline_args {
    size_t struct_size;
    gecon_jacobian_compact_fn jacobian; /* the model's gecon_model_jacobian_compact */
    const int32_t* nz_table;  /* HOST [nnz]: row << 16 | col of the structural non-zeros, grouped A, B, C, D */
    const int32_t* nz_off;    /* HOST [5] */
    int32_t nnz;
    int32_t n;                /* model variables */
    int32_t k;                /* shocks */
    int32_t n_theta;          /* free parameters: the leading columns of a parameter row */
    int32_t n_err;            /* measurement-error sigmas: they follow the k shock sigmas and belong to the FIRST n_err observables
                                 (gEconpy/model/statespace.py:800-808) */
    int32_t p;                /* observables */
    int32_t n_filter;         /* variables handed to the filter (lagged + observed), <= n */
    int32_t n_lead;
    const int32_t* filter_vars; /* HOST [n_filter]: their positions in solver order */
    const int32_t* obs_idx;     /* HOST [p]: positions of the observables WITHIN filter_vars */
    const int32_t* lead_idx;    /* HOST [n_lead]: structural lead variables (solver order) */
    int32_t col_ranges[4];    /* lag_lo, lag_hi, lead_lo, lead_hi (gecon_cr_args) */
    const double* theta;      /* DEVICE [N][theta_stride]: free parameters | sigma_<shock> (k) | error_sigma (n_err) | ... */
    int64_t theta_stride;
    int64_t N;
    const double* Y;          /* DEVICE [Tobs][p] */
    int32_t Tobs;
    int32_t max_iter;
    double tol;               /* cycle reduction */
    double solver_tol;        /* residual gate (statespace.py:1210-1215) */
    double jitter;
    double missing_fill;
    int32_t mvn_const_mode;
    int32_t mask_intercept;
    int32_t gate_mask;
    int32_t check_bk;         /* 0: no Blanchard-Kahn check; 1: count every uncertified draw (status bits as the reference's); 2: "gate only" --
                                 draws that gate_mask already rejects are not counted (same ll, their BK bit stays unset) */
    int32_t scan_semantics;
    int32_t timing;           /* != 0: CUDA events around every kernel; gecon_pipeline_stage_ms then returns the per-stage totals of
                                 the calling thread's last call (jacobian, cr_solve, bk_count, kalman_ll).  Synchronises the stream */
    int64_t chunk;            /* draws per launch (<= 0: 65,536) */
    double* ll;               /* DEVICE [N] out */
    int32_t* status;          /* DEVICE [N] out */
    int32_t* n_iter;          /* DEVICE [N] out or NULL */
    gecon_cr_solve_fn cr_solve; /* NULL: gecon_cr_solve_batched.  gecon_model_loglik sets it to the model library's own
                                 gecon_model_cr_solve -- the warp-per-draw solver compiled with this model's n and column ranges as
                                 compile-time constants (csrc/cr_warp_spec.cu); same contract, same results, falls back to the generic
                                 entry point for arguments it was not built for (GECON_CR_SPEC=0 forces the generic kernel) */
    gecon_kalman_ll_fn kalman_ll; /* NULL: gecon_kalman_ll_batched.  Else the filter compiled for this configuration's (filter
                                 dimension, observables) pair -- gecon_kalman_ll_spec of a csrc/kalman_spec.cu build; same contract, same
                                 results, falls back to the generic entry point by itself (GECON_KF_SPEC=0 forces the generic kernel) */
} gecon_pipeline_args;

int gecon_loglik_pipeline(const gecon_pipeline_args* args, void* stream);
int gecon_pipeline_stage_ms(float* ms4);

/* What the per-model solver builds (gecon_model_cr_solve, csrc/cr_warp_spec.cu) and the per-configuration filter builds
 * (gecon_kalman_ll_spec, csrc/kalman_spec.cu) link against: the argument validation of gecon_cr_solve_* / gecon_kalman_ll_* (0 or
 * GECON_E_*; no device work), and the padded dimension at which gecon_kalman_ll_batched would run the one-warp-per-draw kernel on
 * these arguments (0: it would run the thread-per-draw or the CTA-per-draw kernel). */
int gecon_cr_check_args(const gecon_cr_args* args);
int gecon_kalman_check_args(const gecon_kalman_args* args);
int gecon_kalman_warp_np(const gecon_kalman_args* args);

/* library / device information */
int gecon_abi_version(void);
/* measured fp64 peaks of the current device in TFLOP/s: register-resident DFMA chains and mma.sync.m8n8k4.f64 chains (the
 * denominators of the roofline fractions; a few milliseconds, synchronises the device) */
int gecon_fp64_peak(double* dfma_tflops, double* dmma_tflops);
int gecon_device_count(void);
const char* gecon_get_last_error(void);
/* resident CTAs per SM and dynamic shared memory (bytes) of a kernel for state dimension n:
 * which = 0 cr_solve, 1 kalman_ll (needs p, Tobs), 2 bk_count (n = pencil size), 3 dlyap */
int gecon_kernel_info(int32_t which, int32_t n, int32_t p, int32_t Tobs, int32_t* ctas_per_sm, int32_t* smem_bytes,
                      int32_t* threads);
/* counts kernel launches made by this library in the calling process (bench.py's gpu_launches) */
int64_t gecon_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Per-model generated libraries (geconpy_b200/model/codegen.py emits one per model spec) export:
 *   int gecon_model_info(int32_t* n, int32_t* k, int32_t* n_theta);
 *   int gecon_model_jacobian_batched(const double* theta, int64_t N, double* A, double* B, double* C, double* D,
 *                                    double* xss, int32_t* status, void* stream);
 *   int gecon_model_jacobian_compact(const double* theta, int64_t theta_stride, int64_t N, double* vals, double* xss,
 *                                    int32_t* status, void* stream);          (structural non-zeros only)
 *   int gecon_model_structure(int32_t* nnz, const int32_t** table, const int32_t** off, int32_t* col_ranges,
 *                             int32_t* n_lead, const int32_t** lead_idx);
 *   int gecon_model_loglik(gecon_pipeline_args* args, void* stream);          (the fused entry point above)
 * theta is [N][n_theta]; A,B,C,D come out in the reference's permuted solver order (perturbation.py:130-158).
 * Replaces the compiled pytensor function f(*ss, *params) -> [A,B,C,D] (model.py:1647-1664, build.py:681-695).
 * ------------------------------------------------------------------------------------------------------------- */
typedef int (*gecon_model_jacobian_fn)(const double* theta, int64_t N, double* A, double* B, double* C, double* D, double* xss,
                                       int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GECON_B200_H */
