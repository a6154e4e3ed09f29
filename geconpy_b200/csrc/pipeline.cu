// The fused theta -> log-likelihood entry point: one C call per population, no intermediate dense Jacobians.
//
// SURVEY.md section 8(b) lists `gecon_model_<hash>_loglik(dTheta, N, dY, ..., dLL, dStatus, stream)`; section 3.3 is the path it
// replaces (one compiled logp evaluation per draw: parameters -> steady state -> A, B, C, D -> cycle reduction -> R, residual ->
// Blanchard-Kahn -> P0 -> Kalman filter -> gates; gEconpy/model/statespace.py:725-820, 1139-1215).  Every generated model
// library exports `gecon_model_loglik`, a thin wrapper that fills in its own Jacobian kernel and structure tables and calls
// gecon_loglik_pipeline below.  Per chunk of draws, on the caller's stream:
//   1  generated kernel: theta (read in place, strided) -> COMPACT Jacobian = the structural non-zeros of A, B, C, D only
//   2  cr_solve (one warp per draw for n <= 24): tiles built by scattering the compact entries into shared memory; writes
//      the filter's T[U][:, U], R[U] blocks (U = lagged + observed variables)
//   3  bk_count for the draws the solver kernel could not certify (usually none)
//   4  kalman_ll: shock / measurement-error scales read in place from the parameter vector (q_stride, h_stride, h_count)
// No memsets, no copy kernels, nothing but (theta, Y) in and (ll, status) out: medium NK moves 0.9 KB (compact Jacobian) +
// 1.1 KB (T, R) per draw through L2 instead of 14.6 KB + 1.1 KB through HBM.  Workspace comes from the stream-ordered allocator.
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace gecon {

static thread_local float t_stage_ms[4] = {0.f, 0.f, 0.f, 0.f};

struct StageTimer {
    bool on;
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<int> stage;
    StageTimer(bool on_, cudaStream_t st_) : on(on_), st(st_) {}
    void begin(int s) {
        if (!on) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        ev.push_back(a);
        ev.push_back(b);
        stage.push_back(s);
    }
    void end() {
        if (on) cudaEventRecord(ev.back(), st);
    }
    void collect() {  // synchronises on the last event: only with timing requested
        if (!on) return;
        for (int s = 0; s < 4; ++s) t_stage_ms[s] = 0.f;
        if (!ev.empty()) cudaEventSynchronize(ev.back());
        for (size_t i = 0; i < stage.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
            t_stage_ms[stage[i]] += ms;
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

// Device copies of the index tables, keyed by (device, content): a configuration uploads them once (a blocking copy) and every
// later call finds them here, so the pipeline never synchronises the stream.  Never freed (a few hundred bytes per configuration).
static int cached_tables(const std::vector<int32_t>& h, int32_t** out) {
    static std::mutex mu;
    static std::map<std::pair<int, std::vector<int32_t>>, int32_t*> cache;
    int dev = 0;
    GECON_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, h);
    auto it = cache.find(key);
    if (it == cache.end()) {
        int32_t* d = nullptr;
        GECON_CUDA(cudaMalloc((void**)&d, sizeof(int32_t) * h.size()));
        GECON_CUDA(cudaMemcpy(d, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
        it = cache.emplace(std::move(key), d).first;
    }
    *out = it->second;
    return 0;
}

static int check_pipeline(const gecon_pipeline_args* a) {
    if (!a || a->struct_size != sizeof(gecon_pipeline_args)) {
        set_last_error("gecon_pipeline_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (!a->jacobian || !a->nz_table || !a->nz_off || !a->theta || !a->Y || !a->ll || !a->status || !a->filter_vars || !a->obs_idx || a->N < 0 ||
        a->n < 1 || a->k < 1 || a->p < 1 || a->n_filter < 1 || a->n_filter > a->n || a->Tobs < 0 || a->nnz < 1 ||
        a->theta_stride < a->n_theta + a->k + a->n_err || a->n_err > a->p || (a->check_bk && a->n_lead > 0 && !a->lead_idx)) {
        set_last_error("gecon_pipeline_args: null pointer or inconsistent dimensions");
        return GECON_E_BADARG;
    }
    if (a->n > 64 || a->p > 8) {
        set_last_error("gecon_pipeline_args: unsupported size n = %d (max 64), p = %d (max 8)", a->n, a->p);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    return 0;
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_loglik_pipeline(const gecon_pipeline_args* a, void* stream) {
    int rc = check_pipeline(a);
    if (rc) return rc;
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t chunk = a->chunk > 0 ? a->chunk : 65536;
    const int64_t nc = a->N < chunk ? a->N : chunk;
    const int n = a->n, k = a->k, p = a->p, nf = a->n_filter, nl = a->check_bk ? a->n_lead : 0;
    // ---- workspace (stream-ordered): compact Jacobian, T, R, per-draw scalars, the small index tables
    const size_t b_vals = sizeof(double) * (size_t)nc * a->nnz, b_T = sizeof(double) * (size_t)nc * nf * nf, b_R = sizeof(double) * (size_t)nc * nf * k;
    const size_t b_i = sizeof(int32_t) * (size_t)nc;
    const size_t total = b_vals + b_T + b_R + sizeof(double) * (size_t)nc + 3 * b_i + 256;
    // the small index tables (structure of the model and of the configuration): uploaded once per distinct content, cached
    const size_t n_idx = (size_t)a->nnz + nf + p + (nl > 0 ? nl : 1);
    std::vector<int32_t> h(n_idx, 0);
    memcpy(h.data(), a->nz_table, sizeof(int32_t) * a->nnz);
    memcpy(h.data() + a->nnz, a->filter_vars, sizeof(int32_t) * nf);
    memcpy(h.data() + a->nnz + nf, a->obs_idx, sizeof(int32_t) * p);
    if (nl > 0) memcpy(h.data() + a->nnz + nf + p, a->lead_idx, sizeof(int32_t) * nl);
    int32_t* d_idx = nullptr;
    rc = cached_tables(h, &d_idx);
    if (rc) return rc;
    // T = -A1hat^-1 A has non-zero columns only at the lagged variables [lag_lo, lag_hi): when the filter variables come
    // [lagged ... | others ...] (BatchedStateSpace orders them so), the filter is told how many leading columns of its T can be
    // non-zero (gecon_kalman_args.t_cols) and skips the k-steps of its two products with T beyond them
    int t_cols = 0;
    if (a->col_ranges[1] > a->col_ranges[0]) {
        auto lagged = [&](int v) { return v >= a->col_ranges[0] && v < a->col_ranges[1]; };
        int t = 0;
        while (t < nf && lagged(a->filter_vars[t])) ++t;
        bool rest_outside = true;
        for (int i = t; i < nf; ++i) rest_outside = rest_outside && !lagged(a->filter_vars[i]);
        if (rest_outside && t > 0 && t < nf) t_cols = t;
    }
    int32_t* d_table = d_idx;
    int32_t* d_fv = d_table + a->nnz;
    int32_t* d_obs = d_fv + nf;
    int32_t* d_lead = d_obs + p;
    char* ws = nullptr;
    keep_mempool();
    GECON_CUDA(cudaMallocAsync((void**)&ws, total, st));
    char* cur = ws;
    auto take = [&](size_t bytes) {
        char* r = cur;
        cur += (bytes + 15) & ~(size_t)15;
        return r;
    };
    double* d_vals = (double*)take(b_vals);
    double* d_T = (double*)take(b_T);
    double* d_R = (double*)take(b_R);
    double* d_resid = (double*)take(sizeof(double) * (size_t)nc);
    int32_t* d_st = (int32_t*)take(b_i);
    int32_t* d_it = (int32_t*)take(b_i);
    int32_t* d_nu = (int32_t*)take(b_i);
    StageTimer tm(a->timing != 0, st);
    for (int64_t lo = 0; lo < a->N && rc == 0; lo += nc) {
        const int64_t cnt = (a->N - lo) < nc ? (a->N - lo) : nc;
        const double* th = a->theta + (size_t)lo * a->theta_stride;
        // 1 ---- compact Jacobian
        tm.begin(0);
        rc = a->jacobian(th, a->theta_stride, cnt, d_vals, nullptr, d_st, stream);
        tm.end();
        g_launch_count++;
        if (rc) {
            set_last_error("model Jacobian kernel failed: %s", cudaGetErrorString((cudaError_t)rc));
            break;
        }
        gecon_compact_jac cj{};
        cj.vals = d_vals;
        cj.stride = a->nnz;
        cj.table = d_table;
        for (int q = 0; q < 5; ++q) cj.off[q] = a->nz_off[q];
        // 2 ---- cycle reduction, R, residual, Blanchard-Kahn certificate; writes the filter's blocks
        gecon_cr_args cr{};
        cr.struct_size = sizeof(cr);
        cr.N = cnt;
        cr.n = n;
        cr.k = k;
        cr.max_iter = a->max_iter;
        cr.accumulate = 1;
        cr.tol = a->tol;
        cr.resid_tol = a->solver_tol;
        cr.unperm = d_fv;
        cr.n_out = nf;
        cr.T = d_T;
        cr.R = d_R;
        cr.status = d_st;
        cr.n_iter = d_it;
        cr.resid = d_resid;
        cr.n_lead = nl;
        cr.lead_idx = nl > 0 ? d_lead : nullptr;
        cr.n_unstable = d_nu;
        cr.lag_lo = a->col_ranges[0];
        cr.lag_hi = a->col_ranges[1];
        cr.lead_lo = a->col_ranges[2];
        cr.lead_hi = a->col_ranges[3];
        cr.scan_semantics = a->scan_semantics;
        cr.compact = &cj;
        tm.begin(1);
        rc = a->cr_solve ? a->cr_solve(&cr, stream) : gecon_cr_solve_batched(&cr, stream);
        tm.end();
        if (rc) break;
        // 3 ---- exact Blanchard-Kahn count for the draws without a certificate
        if (nl > 0) {
            gecon_bk_args bk{};
            bk.struct_size = sizeof(bk);
            bk.N = cnt;
            bk.n = n;
            bk.n_lead = nl;
            bk.lead_idx = d_lead;
            bk.accumulate = 1;
            bk.n_unstable = d_nu;
            bk.status = d_st;
            // check_bk == 2 ("gate only"): a draw the gate already rejects (no convergence, residual, ...) is not counted -- its
            // log-likelihood is -inf either way, only its Blanchard-Kahn BIT stays unset.  On a population where half the draws
            // fail the count is most of the step; samplers that only consume the gate ask for this.
            bk.skip_mask = GECON_ST_BK_CERTIFIED | GECON_ST_JAC_NONFINITE |
                           (a->check_bk == 2 ? (a->gate_mask & ~(GECON_ST_BK | GECON_ST_BK_INCONCLUSIVE)) : 0);
            bk.compact = &cj;
            tm.begin(2);
            rc = gecon_bk_count_batched(&bk, stream);
            tm.end();
            if (rc) break;
        }
        // 4 ---- P0 + Kalman filter + gates; the scales are read in place from the parameter vector
        gecon_kalman_args kf{};
        kf.struct_size = sizeof(kf);
        kf.T = d_T;
        kf.R = d_R;
        kf.qdiag = th + a->n_theta;
        kf.q_stride = a->theta_stride;
        kf.hdiag = a->n_err > 0 ? th + a->n_theta + k : nullptr;
        kf.h_stride = a->theta_stride;
        kf.h_count = a->n_err;
        kf.obs_idx = d_obs;
        kf.Y = a->Y;
        kf.N = cnt;
        kf.n = nf;
        kf.k = k;
        kf.p = p;
        kf.Tobs = a->Tobs;
        kf.jitter = a->jitter;
        kf.missing_fill = a->missing_fill;
        kf.mvn_const_mode = a->mvn_const_mode;
        kf.status_in = d_st;
        kf.gate_mask = a->gate_mask;
        kf.sigma_inputs = 1;
        kf.ll = a->ll + lo;
        kf.status = a->status + lo;
        kf.mask_intercept = a->mask_intercept;
        kf.t_cols = t_cols;
        tm.begin(3);
        rc = a->kalman_ll ? a->kalman_ll(&kf, stream) : gecon_kalman_ll_batched(&kf, stream);
        tm.end();
        if (rc) break;
        if (a->n_iter) {
            const cudaError_t e = cudaMemcpyAsync(a->n_iter + lo, d_it, sizeof(int32_t) * (size_t)cnt, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) rc = fail_cuda(e, "n_iter copy");
        }
    }
    tm.collect();
    cudaFreeAsync(ws, st);
    return rc;
}

extern "C" int gecon_pipeline_stage_ms(float* ms4) {
    if (!ms4) return GECON_E_BADARG;
    for (int s = 0; s < 4; ++s) ms4[s] = t_stage_ms[s];
    return 0;
}
