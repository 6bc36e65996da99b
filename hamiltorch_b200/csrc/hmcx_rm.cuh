// hmcx_rm.cuh -- definitions shared by the RMHMC kernels: hmcx_rmhmc.cu (one thread / one warp pair per chain, D <= 16)
// and hmcx_rmhmc_cta.cu (one CTA per chain, metric and eigenvectors in shared memory, D <= 64).
#pragma once
#include "hmcx_common.cuh"

namespace hmcx {

struct RmTarget {
    int kind, D;
    float log_norm, inv_var_v;
    const float* mean;
    const float* ivar;
    const float* prec;        // GAUSS_FULL: [D,D] row-major
};

struct RmCfg {
    int softabs;          // Metric.SOFTABS (1) or Metric.HESSIAN (0)
    int jacdiag;          // Metric.JACOBIAN_DIAG (:100-106): G = diag((d log p / d theta_i)^2), no eigen-decomposition
    float alpha, jitter;  // softabs_const, jitter scale (jitter < 0: none)
    float pi_term;        // D*log(2*pi) in fp32 as samplers.py:712
};

}  // namespace hmcx
