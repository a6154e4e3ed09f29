// public C-ABI argument structs (include/gecon_b200.h) -> the internal argument structs of grad.cuh
#pragma once
#include "../../include/gecon_b200.h"
#include "grad.cuh"

namespace gecon_grad {

inline KalmanGradArgs to_internal(const gecon_kalman_grad_args& a) {
    KalmanGradArgs g{};
    g.T = a.T, g.R = a.R, g.qdiag = a.qdiag, g.q_stride = a.q_stride, g.hdiag = a.hdiag, g.h_stride = a.h_stride;
    g.Z = a.Z, g.z_stride = a.z_stride, g.obs_idx = a.obs_idx, g.d = a.d, g.d_stride = a.d_stride, g.Y = a.Y;
    g.N = a.N, g.n = a.n, g.k = a.k, g.p = a.p, g.Tobs = a.Tobs;
    g.jitter = a.jitter, g.missing_fill = a.missing_fill, g.mvn_const_mode = a.mvn_const_mode, g.lyap_max_iter = a.lyap_max_iter;
    g.status_in = a.status_in, g.gate_mask = a.gate_mask, g.sigma_inputs = a.sigma_inputs;
    g.mask_intercept = a.mask_intercept;
    g.qfull = a.qfull, g.qfull_stride = a.qfull_stride, g.qfull_bar = a.qfull_bar;
    g.ll = a.ll, g.status = a.status, g.T_bar = a.T_bar, g.R_bar = a.R_bar, g.q_bar = a.q_bar, g.h_bar = a.h_bar, g.d_bar = a.d_bar, g.Z_bar = a.Z_bar;
    g.traj = nullptr, g.c0bar_ws = nullptr;
    return g;
}

inline PolicyAdjointArgs to_internal(const gecon_policy_adjoint_args& a) {
    PolicyAdjointArgs g{};
    g.A = a.A, g.B = a.B, g.C = a.C, g.D = a.D, g.T = a.T, g.R = a.R, g.T_bar = a.T_bar, g.R_bar = a.R_bar;
    g.N = a.N, g.n = a.n, g.k = a.k, g.max_iter = a.max_iter;
    g.A_bar = a.A_bar, g.B_bar = a.B_bar, g.C_bar = a.C_bar, g.D_bar = a.D_bar, g.status = a.status;
    return g;
}

}  // namespace gecon_grad
