// Batched linear propagation  X_t = T X_{t-1} + R E_t  over parameter draws: the device backend of the reference's
// posterior-batched simulation, impulse responses and autocovariances:
//   gEconpy/model/simulate.py:171-183  _simulate_linear_system   out[0] = R e_0, out[t] = T out[t-1] + R e_t
//   gEconpy/model/simulate.py:201-318  impulse_response_function one trajectory per shock (m = k columns)
//   gEconpy/model/statistics/covariance.py:133-161 _compute_autocovariance_matrix  Gamma_h = T^h Sigma (X_0 = Sigma, no shocks)
// One CTA per draw; T, R and the two n x m state panels live in shared memory; plain DFMA (this path is bandwidth
// bound on the output stream [N][L][n][m], not on the n^2 m flops per step).
#include "common.cuh"

namespace gecon {

__global__ void __launch_bounds__(256) propagate_kernel(const gecon_propagate_args p) {
    extern __shared__ __align__(16) double sm[];
    const int n = p.n, k = p.k, m = p.m, L = p.L;
    const int ldt = n | 1;  // odd leading dimension: column walks of T are conflict-free
    double* Ts = sm;                    // [n][ldt]
    double* Rs = Ts + n * ldt;          // [n][k]
    double* X0 = Rs + n * k;            // [n][m]
    double* X1 = X0 + n * m;            // [n][m]
    double* Es = X1 + n * m;            // [k][m]
    for (long long draw = blockIdx.x; draw < p.N; draw += gridDim.x) {
        const double* gT = p.T + (size_t)draw * n * n;
        for (int i = threadIdx.x; i < n * n; i += blockDim.x) Ts[(i / n) * ldt + (i % n)] = gT[i];
        if (p.R)
            for (int i = threadIdx.x; i < n * k; i += blockDim.x) Rs[i] = p.R[(size_t)draw * n * k + i];
        for (int i = threadIdx.x; i < n * m; i += blockDim.x) X0[i] = p.X0 ? p.X0[(size_t)draw * n * m + i] : 0.0;
        __syncthreads();
        double* cur = X0;
        double* nxt = X1;
        double* out = p.out + (size_t)draw * L * n * m;
        for (int t = 0; t < L; ++t) {
            const double* gE = (p.E && p.R) ? p.E + (size_t)(p.e_stride ? draw * p.e_stride : 0) + (size_t)t * k * m : nullptr;
            if (gE) {
                for (int i = threadIdx.x; i < k * m; i += blockDim.x) Es[i] = gE[i];
                __syncthreads();
            }
            const bool keep = (t == 0 && p.start_at_x0);  // out[0] = X0 (+ R e_0): the autocovariance at lag 0
            for (int idx = threadIdx.x; idx < n * m; idx += blockDim.x) {
                const int i = idx / m, c = idx - i * m;
                double acc = 0.0;
                if (keep) {
                    acc = cur[idx];
                } else if (t > 0 || p.X0) {
                    for (int j = 0; j < n; ++j) acc = fma(Ts[i * ldt + j], cur[j * m + c], acc);
                }
                if (gE)
                    for (int s = 0; s < k; ++s) acc = fma(Rs[i * k + s], Es[s * m + c], acc);
                nxt[idx] = acc;
                out[(size_t)t * n * m + idx] = acc;
            }
            __syncthreads();
            double* tmp = cur;
            cur = nxt;
            nxt = tmp;
        }
    }
}

static int check_prop_args(const gecon_propagate_args* a) {
    if (!a || a->struct_size != sizeof(gecon_propagate_args)) {
        set_last_error("gecon_propagate_args: bad struct_size");
        return GECON_E_BADARG;
    }
    if (!a->T || !a->out || a->N < 0 || a->n < 1 || a->k < 0 || a->m < 1 || a->L < 0 || (a->E && !a->R) || (a->start_at_x0 && !a->X0)) {
        set_last_error("gecon_propagate_args: null pointer or bad dimension");
        return GECON_E_BADARG;
    }
    return 0;
}

static size_t prop_smem(const gecon_propagate_args& a) {
    const size_t n = a.n, k = a.k, m = a.m;
    return sizeof(double) * (n * (n | 1) + n * k + 2 * n * m + k * m);
}

}  // namespace gecon

using namespace gecon;

extern "C" int gecon_propagate_batched(const gecon_propagate_args* a, void* stream) {
    int rc = check_prop_args(a);
    if (rc) return rc;
    if (a->N == 0 || a->L == 0) return 0;
    const size_t smem = prop_smem(*a);
    if (smem > 227 * 1024) {
        set_last_error("gecon_propagate: n = %d, k = %d, m = %d do not fit in shared memory (%zu bytes)", a->n, a->k, a->m, smem);
        return GECON_E_UNSUPPORTED_SIZE;
    }
    const int work = a->n * a->m;
    const int threads = work >= 256 ? 256 : ((work + 31) / 32) * 32;
    int grid = 0;
    rc = persistent_grid(propagate_kernel, threads, smem, a->N, &grid, nullptr);
    if (rc) return rc;
    propagate_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(*a);
    g_launch_count++;
    GECON_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int gecon_propagate_host(const gecon_propagate_args* a) {
    int rc = check_prop_args(a);
    if (rc) return rc;
    if (a->N == 0 || a->L == 0) return 0;
    const size_t N = (size_t)a->N, n = a->n, k = a->k, m = a->m, L = a->L;
    DevBuf dT, dR, dX, dE, dO;
    gecon_propagate_args d = *a;
    GECON_CUDA(dT.alloc(N * n * n * 8));
    GECON_CUDA(cudaMemcpy(dT.p, a->T, N * n * n * 8, cudaMemcpyHostToDevice));
    d.T = dT.as<double>();
    if (a->R) {
        GECON_CUDA(dR.alloc(N * n * k * 8));
        GECON_CUDA(cudaMemcpy(dR.p, a->R, N * n * k * 8, cudaMemcpyHostToDevice));
        d.R = dR.as<double>();
    }
    if (a->X0) {
        GECON_CUDA(dX.alloc(N * n * m * 8));
        GECON_CUDA(cudaMemcpy(dX.p, a->X0, N * n * m * 8, cudaMemcpyHostToDevice));
        d.X0 = dX.as<double>();
    }
    if (a->E) {
        const size_t ne = (a->e_stride ? N : 1) * L * k * m;
        GECON_CUDA(dE.alloc(ne * 8));
        GECON_CUDA(cudaMemcpy(dE.p, a->E, ne * 8, cudaMemcpyHostToDevice));
        d.E = dE.as<double>();
    }
    GECON_CUDA(dO.alloc(N * L * n * m * 8));
    d.out = dO.as<double>();
    rc = gecon_propagate_batched(&d, nullptr);
    if (rc) return rc;
    GECON_CUDA(cudaMemcpy(a->out, d.out, N * L * n * m * 8, cudaMemcpyDeviceToHost));
    return 0;
}
