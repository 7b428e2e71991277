// params_host.hpp -- host-side helpers shared by the C ABI (igt_abi.cu) and the CPU test harness
// (tests/hostsim): default igt_params and the igt_params -> DevParams<T> conversion.
#pragma once
#include <cstring>
#include "../../include/igt_mpc.h"
#include "solver_core.cuh"

namespace igt {

template <typename T>
inline void fill_dev_params(const igt_params &p, DevParams<T> &d)
{
    memset(&d, 0, sizeof(d));
    d.N = p.N; d.n_rk = p.n_rk; d.n_cinf = p.n_cinf; d.max_iter = p.max_iter; d.n_alpha = p.n_alpha;
    d.second_order = p.second_order; d.n_layers = 0; d.stall_iter = p.stall_iter; d.stall_rp = T(p.stall_rp);
    d.max_trials = p.max_trials > 0 ? p.max_trials : (1 << 30);
    d.dt = T(p.dt); d.h = T(p.dt / p.n_rk); d.l_r = T(p.l_r); d.lsum = T(p.l_r + p.l_f);
    d.inv_lr = T(1.0 / p.l_r); d.rho = T(p.l_r / (p.l_f + p.l_r));
    d.v_min = T(p.v_min); d.v_max = T(p.v_max); d.a_min = T(p.a_min); d.a_max = T(p.a_max);
    d.df_max = T(p.df_max); d.ey_lim = T(p.ey_lim); d.da_max = T(p.da_max); d.ddf_max = T(p.ddf_max);
    d.d_min = T(p.d_min); d.w_u = T(p.w_u);
    d.tol = T(p.tol); d.tol_rp = T(p.tol_rp); d.tol_comp = T(p.tol_comp); d.mu0 = T(p.mu0);
    d.mu_floor = T(p.mu_floor); d.kappa_eps = T(p.kappa_eps); d.kappa_mu = T(p.kappa_mu);
    d.theta_mu = T(p.theta_mu); d.y_init_min = T(p.y_init_min); d.tau_min = T(p.tau_min);
    d.mu0_warm = T(p.mu0_warm); d.y_init_min_warm = T(p.y_init_min_warm); d.alpha_safety = T(0.99);
    d.reg_min = T(p.reg_min); d.reg_up = T(p.reg_up); d.reg_down = T(p.reg_down); d.reg_max = T(p.reg_max);
    d.reg_jump = T(p.reg_jump);
    d.eps_phi = T(p.eps_phi); d.gamma_theta = T(p.gamma_theta); d.theta_small = T(p.theta_small);
    d.acc_tol = T(p.acc_tol); d.acc_rp = T(p.acc_rp); d.acc_comp = T(p.acc_comp); d.x0_tol = T(p.x0_tol);
    for (int m = 0; m < p.n_cinf; m++) {
        d.cinf_A[m][0] = T(p.cinf_A[m][0]); d.cinf_A[m][1] = T(p.cinf_A[m][1]); d.cinf_b[m] = T(p.cinf_b[m]);
    }
}


inline int default_params(igt_params *p, int precision)
{
    if (!p) return IGT_EINVAL;
    memset(p, 0, sizeof(*p));
    p->N = 40; p->n_rk = 4; p->dt = 0.1; p->l_r = 2.235; p->l_f = 2.235;
    p->v_min = 0.0; p->v_max = 5.0; p->a_min = -4.0; p->a_max = 3.0; p->df_max = 1.0; p->ey_lim = 0.2;
    p->da_max = 0.1 * 0.9; p->ddf_max = 0.1 * 0.7; p->d_min = 5.6; p->w_u = 0.05;
    p->n_cinf = 0;
    p->mu0 = 0.3; p->kappa_eps = 10.0; p->kappa_mu = 0.2; p->theta_mu = 1.5; p->y_init_min = 0.3;
    p->tau_min = 0.99; p->reg_min = 1e-4; p->reg_up = 10.0; p->reg_down = 10.0; p->reg_max = 1e10; p->reg_jump = 1.1;
    p->gamma_theta = 1e-6; p->max_iter = 60; p->n_alpha = 6; p->second_order = 1;
    p->mu0_warm = 1e-3; p->y_init_min_warm = 1e-2;   // see igt_mpc.h: softer than round 1's (1e-4, 1e-3)
    p->stall_iter = 16; p->stall_rp = 1e-2; p->max_trials = 0;
    p->precision = precision;
    p->acc_tol = 1e-3; p->acc_rp = 1e-6; p->acc_comp = 1e-4; p->x0_tol = 1e-6;
    if (precision == IGT_PREC_F64) {
        p->tol = 1e-6; p->tol_rp = 1e-8; p->tol_comp = 1e-7; p->mu_floor = 1e-8;
        p->eps_phi = 1e-12; p->theta_small = 1e-10;
    } else {
        p->tol = 1e-3; p->tol_rp = 2e-5; p->tol_comp = 1e-4; p->mu_floor = 2e-5;
        p->eps_phi = 1e-6; p->theta_small = 1e-4;
    }
    return IGT_OK;
}


}  // namespace igt
