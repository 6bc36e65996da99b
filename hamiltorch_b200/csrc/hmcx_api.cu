// hmcx_api.cu -- the extern "C" surface declared in include/hmcx.h; validates and dispatches on target kind.
#include <cstdlib>
#include "hmcx_common.cuh"

namespace hmcx {
int elem_leapfrog(const hmcx_target_t*, const hmcx_mass_t*, const float*, const float*, const float*, int, int, int,
                  float*, float*, float*, float*, cudaStream_t);
int elem_hamiltonian(const hmcx_target_t*, const hmcx_mass_t*, const float*, const float*, int, int, float*,
                     uint8_t*, cudaStream_t);
int elem_gibbs(const hmcx_mass_t*, const hmcx_rng_t*, int, int, int, int64_t, float*, cudaStream_t);
int elem_hmc_run(const hmcx_target_t*, const hmcx_mass_t*, const hmcx_rng_t*, const hmcx_nuts_t*, const float*,
                 float*, float*, int, int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*,
                 int, float*, const hmcx_sink_t*, cudaStream_t);
int coupled_leapfrog(const hmcx_target_t*, const hmcx_mass_t*, const float*, const float*, const float*, int, int, int,
                     float*, float*, float*, float*, cudaStream_t);
int coupled_hamiltonian(const hmcx_target_t*, const hmcx_mass_t*, const float*, const float*, int, int, float*, uint8_t*,
                        cudaStream_t);
size_t dense_rmhmc_workspace_floats(int, int);
int dense_rmhmc_run(const hmcx_target_t*, const hmcx_rmhmc_t*, const hmcx_const_metric_t*, const hmcx_rng_t*,
                    const float*, float*, const float*, int, int, int, int, int, int, int, float*, uint8_t*, uint8_t*,
                    float*, int32_t*, float*, cudaStream_t);
int mlp_split_run(const hmcx_target_t*, const hmcx_mass_t*, const hmcx_rng_t*, const hmcx_nuts_t*, int, const float*,
                  float*, float*, int, int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*,
                  cudaStream_t, const float*, float*, float*);
int mlp_leapfrog(const hmcx_target_t*, const hmcx_mass_t*, const hmcx_rng_t*, int, double, const float*, const float*, float*,
                 int, int, int, float*, float*, cudaStream_t);
int mlp_grad_log_prob(const hmcx_target_t*, const float*, int, int, int, float*, float*, cudaStream_t);
int mlp_predict(const hmcx_target_t*, const float*, int, int, float*, float*, cudaStream_t);
int small_hmc_run(const hmcx_target_t*, const hmcx_mass_t*, const hmcx_rng_t*, const hmcx_nuts_t*, const float*, float*,
                  float*, int, int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*,
                  cudaStream_t);
int gemm_nt_tf32x3(const float*, const float*, float*, int, int, int, cudaStream_t);
size_t dense_workspace_floats(int, int, int);
int dense_hmc_run(const hmcx_target_t*, const hmcx_mass_t*, const hmcx_rng_t*, const hmcx_nuts_t*, const float*, float*,
                  float*, int, int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*, float*,
                  cudaStream_t);
int rmhmc_run(const hmcx_target_t*, const hmcx_rmhmc_t*, const hmcx_rng_t*, const float*, float*, const float*, int,
              int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*, cudaStream_t);
size_t mlp_packed_x_floats(const hmcx_target_t*);
int mlp_pack_x(const hmcx_target_t*, float*, cudaStream_t);
int rmhmc_cta_run(const hmcx_target_t*, const hmcx_rmhmc_t*, const hmcx_rng_t*, const float*, float*, const float*, int,
                  int, int, int, int, int, int, float*, uint8_t*, uint8_t*, float*, int32_t*, const float*, float*, float*,
                  float*, float*, int, float*, cudaStream_t);
}  // namespace hmcx

static inline bool is_elem(const hmcx_target_t* t) {
    return t && (t->kind == HMCX_TARGET_GAUSS_ISO || t->kind == HMCX_TARGET_GAUSS_DIAG);
}

extern "C" {

int hmcx_abi_version(void) { return HMCX_ABI_VERSION; }

size_t hmcx_hmc_workspace_bytes(const hmcx_target_t* target, const hmcx_mass_t* mass, int32_t C, int32_t ld) {
    const bool full_mass = mass && mass->kind == HMCX_MASS_FULL;
    if (is_elem(target) && !full_mass && ld > 4096) return (size_t)C * (size_t)ld * sizeof(float);
    if (is_elem(target) && !full_mass) return (size_t)C * sizeof(float);      // log p(q_cur) carried between windows of iterations
    if (target && target->dim > 16 &&
        (target->kind == HMCX_TARGET_GAUSS_FULL || (full_mass && is_elem(target))))
        return hmcx::dense_workspace_floats(C, target->dim, full_mass ? 1 : 0) * sizeof(float);
    return 0;
}

const char* hmcx_status_string(int status) {
    switch (status) {
        case HMCX_OK: return "ok";
        case HMCX_ERR_INVALID_ARG: return "invalid argument";
        case HMCX_ERR_UNSUPPORTED: return "unsupported target / mass / dimension combination";
        case HMCX_ERR_CUDA: return "CUDA launch error";
        default: return "unknown status";
    }
}

int hmcx_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass, const float* q_in, const float* p_in,
                  const float* eps, int32_t C, int32_t ld, int32_t L, float* q_out, float* p_out, float* q_traj,
                  float* p_traj, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (is_elem(target) && !(mass && mass->kind == HMCX_MASS_FULL))
        return hmcx::elem_leapfrog(target, mass, q_in, p_in, eps, C, ld, L, q_out, p_out, q_traj, p_traj,
                                   (cudaStream_t)stream);
    // coupled gradient (GAUSS_FULL, FUNNEL) or full mass matrix: one CTA per chain, any D (hmcx_coupled.cu)
    return hmcx::coupled_leapfrog(target, mass, q_in, p_in, eps, C, ld, L, q_out, p_out, q_traj, p_traj,
                                  (cudaStream_t)stream);
}

int hmcx_hamiltonian(const hmcx_target_t* target, const hmcx_mass_t* mass, const float* q, const float* p,
                     int32_t C, int32_t ld, float* H_out, uint8_t* flags_out, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (is_elem(target) && !(mass && mass->kind == HMCX_MASS_FULL))
        return hmcx::elem_hamiltonian(target, mass, q, p, C, ld, H_out, flags_out, (cudaStream_t)stream);
    return hmcx::coupled_hamiltonian(target, mass, q, p, C, ld, H_out, flags_out, (cudaStream_t)stream);
}

int hmcx_gibbs(const hmcx_mass_t* mass, const hmcx_rng_t* rng, int32_t D, int32_t C, int32_t ld, int64_t iter,
               float* p_out, void* stream) {
    return hmcx::elem_gibbs(mass, rng, D, C, ld, iter, p_out, (cudaStream_t)stream);
}

int hmcx_hmc_run(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                 const hmcx_nuts_t* nuts, const float* q_init, float* q_cur, float* eps, int32_t C, int32_t ld,
                 int32_t L, int32_t num_samples, int32_t burn, int32_t iter_begin, int32_t iter_end,
                 float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                 int32_t* num_rejected, int32_t tuning, float* workspace, void* stream) {
    return hmcx_hmc_run_sink(target, mass, rng, nuts, q_init, q_cur, eps, C, ld, L, num_samples, burn, iter_begin,
                             iter_end, samples_out, accept_out, diverged_out, ham_out, num_rejected, tuning, workspace,
                             nullptr, stream);
}

int hmcx_hmc_run_sink(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                      const hmcx_nuts_t* nuts, const float* q_init, float* q_cur, float* eps, int32_t C, int32_t ld,
                      int32_t L, int32_t num_samples, int32_t burn, int32_t iter_begin, int32_t iter_end,
                      float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                      int32_t* num_rejected, int32_t tuning, float* workspace, const hmcx_sink_t* sink, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (sink && sink->thin < 1) return HMCX_ERR_INVALID_ARG;
    if (sink && sink->thin == 1 && !sink->sum && !sink->sumsq) sink = nullptr;
    const bool full_mass = mass && mass->kind == HMCX_MASS_FULL;
    if (is_elem(target) && !full_mass)
        return hmcx::elem_hmc_run(target, mass, rng, nuts, q_init, q_cur, eps, C, ld, L, num_samples, burn,
                                  iter_begin, iter_end, samples_out, accept_out, diverged_out, ham_out,
                                  num_rejected, tuning, workspace, sink, (cudaStream_t)stream);
    if (sink) return HMCX_ERR_UNSUPPORTED;
    if (target->dim > 16 && (target->kind == HMCX_TARGET_GAUSS_FULL || (full_mass && is_elem(target))))
        // dense target and / or full mass matrix at scale: tcgen05 GEMMs over all chains per leapfrog step (hmcx_tc.cu)
        return hmcx::dense_hmc_run(target, mass, rng, nuts, q_init, q_cur, eps, C, ld, L, num_samples, burn,
                                   iter_begin, iter_end, samples_out, accept_out, diverged_out, ham_out,
                                   num_rejected, workspace, (cudaStream_t)stream);
    if (is_elem(target) || target->kind == HMCX_TARGET_GAUSS_FULL || target->kind == HMCX_TARGET_FUNNEL)
        // coupled gradient or full mass matrix: thread-per-chain kernel, D <= 16 (samplers.py:293-294, :811-812, :198-199)
        return hmcx::small_hmc_run(target, mass, rng, nuts, q_init, q_cur, eps, C, ld, L, num_samples, burn,
                                   iter_begin, iter_end, samples_out, accept_out, diverged_out, ham_out,
                                   num_rejected, (cudaStream_t)stream);
    if (target->kind == HMCX_TARGET_MLP)      // un-split Bayesian NN == sample_model (samplers.py:1261)
        return hmcx::mlp_split_run(target, mass, rng, nuts, HMCX_SCHEME_PLAIN, q_init, q_cur, eps, C, ld, L,
                                   num_samples, burn, iter_begin, iter_end, samples_out, accept_out, diverged_out,
                                   ham_out, num_rejected, (cudaStream_t)stream, nullptr, nullptr, nullptr);
    return HMCX_ERR_UNSUPPORTED;
}

int hmcx_split_run(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                   const hmcx_nuts_t* nuts, int32_t scheme, const float* q_init, float* q_cur, float* eps, int32_t C,
                   int32_t ld, int32_t L, int32_t num_samples, int32_t burn, int32_t iter_begin, int32_t iter_end,
                   float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                   int32_t* num_rejected, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (target->kind != HMCX_TARGET_MLP) return HMCX_ERR_UNSUPPORTED;
    return hmcx::mlp_split_run(target, mass, rng, nuts, scheme, q_init, q_cur, eps, C, ld, L, num_samples, burn,
                               iter_begin, iter_end, samples_out, accept_out, diverged_out, ham_out, num_rejected,
                               (cudaStream_t)stream, nullptr, nullptr, nullptr);
}

int hmcx_split_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng, int32_t scheme,
                        double step_size, const float* q_in, const float* p_in, float* eps, int32_t C, int32_t ld, int32_t L,
                        float* q_traj, float* p_traj, void* stream) {
    if (!target || !rng) return HMCX_ERR_INVALID_ARG;
    if (target->kind != HMCX_TARGET_MLP) return HMCX_ERR_UNSUPPORTED;
    return hmcx::mlp_leapfrog(target, mass, rng, scheme, step_size, q_in, p_in, eps, C, ld, L, q_traj, p_traj,
                              (cudaStream_t)stream);
}

int hmcx_rmhmc_run(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng, const float* q_init,
                   float* q_cur, const float* eps, int32_t C, int32_t ld, int32_t L, int32_t num_samples,
                   int32_t burn, int32_t iter_begin, int32_t iter_end, float* samples_out, uint8_t* accept_out,
                   uint8_t* diverged_out, float* ham_out, int32_t* num_rejected, void* stream) {
    // metric / eigenvectors in shared memory, one CTA per chain (hmcx_rmhmc_cta.cu); HMCX_RMHMC_FORCE_CTA=1 sends the small
    // problems there too (tests run every golden chain through both kernels)
    const char* force = getenv("HMCX_RMHMC_FORCE_CTA");
    if (target && (target->dim > 16 || (force && force[0] == '1')))
        return hmcx::rmhmc_cta_run(target, cfg, rng, q_init, q_cur, eps, C, ld, L, num_samples, burn, iter_begin, iter_end,
                                   samples_out, accept_out, diverged_out, ham_out, num_rejected, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, 0, nullptr, (cudaStream_t)stream);
    return hmcx::rmhmc_run(target, cfg, rng, q_init, q_cur, eps, C, ld, L, num_samples, burn, iter_begin, iter_end,
                           samples_out, accept_out, diverged_out, ham_out, num_rejected, (cudaStream_t)stream);
}

int hmcx_rmhmc_leapfrog(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng, const float* q_in,
                        const float* p_in, const float* eps, int32_t C, int32_t ld, int32_t L, float* q_traj,
                        float* p_traj, float* q_copy_out, float* p_copy_out, uint8_t* flags_out, void* stream) {
    if (!q_in || !p_in || !q_traj || !p_traj) return HMCX_ERR_INVALID_ARG;
    // one "iteration" with the momentum given and no MH
    return hmcx::rmhmc_cta_run(target, cfg, rng, q_in, nullptr, eps, C, ld, L, 1, 0, 0, 1, nullptr, nullptr, flags_out,
                               nullptr, nullptr, p_in, q_traj, p_traj, q_copy_out, p_copy_out, 0, nullptr,
                               (cudaStream_t)stream);
}

int hmcx_rmhmc_hamiltonian(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng, const float* q,
                           const float* p, int32_t C, int32_t ld, float* H_out, uint8_t* flags_out, void* stream) {
    if (!q || !p || !H_out) return HMCX_ERR_INVALID_ARG;
    return hmcx::rmhmc_cta_run(target, cfg, rng, q, nullptr, nullptr, C, ld, 1, 1, 0, 0, 1, nullptr, nullptr, flags_out,
                               nullptr, nullptr, p, nullptr, nullptr, nullptr, nullptr, 1, H_out, (cudaStream_t)stream);
}

size_t hmcx_rmhmc_dense_workspace_bytes(int32_t C, int32_t D) {
    return hmcx::dense_rmhmc_workspace_floats(C, D) * sizeof(float);
}

int hmcx_rmhmc_dense_run(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_const_metric_t* metric,
                         const hmcx_rng_t* rng, const float* q_init, float* q_cur, const float* eps, int32_t C,
                         int32_t ld, int32_t L, int32_t num_samples, int32_t burn, int32_t iter_begin, int32_t iter_end,
                         float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                         int32_t* num_rejected, float* workspace, void* stream) {
    return hmcx::dense_rmhmc_run(target, cfg, metric, rng, q_init, q_cur, eps, C, ld, L, num_samples, burn, iter_begin,
                                 iter_end, samples_out, accept_out, diverged_out, ham_out, num_rejected, workspace,
                                 (cudaStream_t)stream);
}

int hmcx_gemm_nt_tf32x3(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K, void* stream) {
    return hmcx::gemm_nt_tf32x3(A, B, D, M, N, K, (cudaStream_t)stream);
}

int hmcx_grad_log_prob(const hmcx_target_t* target, const float* q, int32_t C, int32_t ld, int32_t split,
                       float* grad_out, float* log_prob_out, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (target->kind == HMCX_TARGET_MLP)
        return hmcx::mlp_grad_log_prob(target, q, C, ld, split, grad_out, log_prob_out, (cudaStream_t)stream);
    return HMCX_ERR_UNSUPPORTED;
}

int hmcx_mlp_predict(const hmcx_target_t* target, const float* samples, int32_t S, int32_t ld, float* pred_out,
                     float* log_prob_out, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (target->kind != HMCX_TARGET_MLP) return HMCX_ERR_UNSUPPORTED;
    return hmcx::mlp_predict(target, samples, S, ld, pred_out, log_prob_out, (cudaStream_t)stream);
}

size_t hmcx_mlp_packed_x_bytes(const hmcx_target_t* target) {
    if (!target || target->kind != HMCX_TARGET_MLP) return 0;
    return hmcx::mlp_packed_x_floats(target) * sizeof(float);
}

int hmcx_mlp_pack_x(const hmcx_target_t* target, float* packed_out, void* stream) {
    if (!target) return HMCX_ERR_INVALID_ARG;
    if (target->kind != HMCX_TARGET_MLP) return HMCX_ERR_UNSUPPORTED;
    return hmcx::mlp_pack_x(target, packed_out, (cudaStream_t)stream);
}

int hmcx_copy_rows_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                         void* stream) {
    if (!dst || !src || width > dpitch || width > spitch) return HMCX_ERR_INVALID_ARG;
    if (width == 0 || height == 0) return HMCX_OK;
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, (cudaStream_t)stream) == cudaSuccess
               ? HMCX_OK : HMCX_ERR_CUDA;
}

}  // extern "C"
