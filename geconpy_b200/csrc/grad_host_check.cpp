// TEST INFRASTRUCTURE: compiles the per-draw routines of grad.cuh as plain C++ (one "thread", barriers vanish) so that
// the CPU test-suite can check their arithmetic against oracle/adjoints.py without a GPU.  Never loaded by the product.
//   g++ -O2 -shared -fPIC -DGECON_HOST_CHECK -I include grad_host_check.cpp -o libgecon_grad_hostcheck.so
#ifndef GECON_HOST_CHECK
#define GECON_HOST_CHECK
#endif
#include <vector>

#include "../../include/gecon_b200.h"
#include "grad.cuh"
#include "grad_args.h"

extern "C" int gecon_kalman_grad_hostcheck(const gecon_kalman_grad_args* a) {
    gecon_grad::KalmanGradArgs g = gecon_grad::to_internal(*a);
    std::vector<double> sm(gecon_grad::kalman_grad_smem_doubles(g.n, g.k, g.p, 1));
    std::vector<double> traj((size_t)g.Tobs * gecon_grad::kalman_grad_traj_stride(g.n, g.p) + 1), c0b(2 * (size_t)g.n * g.n);
    g.traj = traj.data();
    g.c0bar_ws = c0b.data();
    for (long long i = 0; i < g.N; ++i) gecon_grad::kalman_grad_draw(g, i, 0, sm.data());
    return 0;
}

extern "C" int gecon_policy_adjoint_hostcheck(const gecon_policy_adjoint_args* a) {
    gecon_grad::PolicyAdjointArgs g = gecon_grad::to_internal(*a);
    std::vector<double> sm(gecon_grad::policy_adjoint_smem_doubles(g.n, 1));
    int s_int[4];
    for (long long i = 0; i < g.N; ++i) gecon_grad::policy_adjoint_draw(g, i, sm.data(), s_int);
    return 0;
}
