/*
 * hmcx.h -- C ABI of libhmcx.so, the B200 (sm_100a) batched-chain Hamiltonian Monte Carlo engine.
 *
 * Drop-in boundary for the hot path of AdamCobb/hamiltorch (reference, pure Python): the reference has no
 * FFI of its own, its boundary is the Python function surface.  Each entry point below
 * names the reference function (hamiltorch/samplers.py, file:line) whose arithmetic it replaces for a whole
 * batch of C independent chains; hamiltorch_b200/_native.py binds them with ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers + sizes + an opaque cudaStream_t (void*); the library never
 *     allocates, never synchronises, keeps no global state; calls are re-entrant given distinct buffers.
 *   - state arrays are fp32 row-major (C, ld): row c = chain c, ld >= D, ld % 4 == 0, 16-byte aligned,
 *     pad elements are ignored on input and written as 0.
 *   - return value: HMCX_OK or a negative HMCX_ERR_* code (hmcx_status_string()).  Numerical divergence is
 *     NOT an error: like the reference's LogProbError -> reject (samplers.py:1045) it is reported per chain
 *     and iteration in `diverged_out`.
 *   - fp32 arithmetic follows the reference's operation order with fused multiply-add contraction disabled,
 *     so that elementwise state is bit-identical to the reference's PyTorch-CPU path; reductions (the
 *     Hamiltonian sums) differ from torch.dot only in summation order.
 */
#ifndef HMCX_H
#define HMCX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMCX_ABI_VERSION 7

#define HMCX_MLP_TC_AUTO 0
#define HMCX_MLP_TC_OFF  1

/* status codes */
#define HMCX_OK                0
#define HMCX_ERR_INVALID_ARG  -1
#define HMCX_ERR_UNSUPPORTED  -2
#define HMCX_ERR_CUDA         -3

/* hmcx_target_t.kind -- the log-densities the kernels can differentiate (hamiltorch_b200/targets.py) */
#define HMCX_TARGET_GAUSS_ISO   0   /* log p = -0.5*sum(x*x) + log_norm                               */
#define HMCX_TARGET_GAUSS_DIAG  1   /* log p = -0.5*sum((x-mean)^2*inv_var) + log_norm                 */
#define HMCX_TARGET_GAUSS_FULL  2   /* log p = -0.5*(x-mean).P(x-mean) + log_norm                      */
#define HMCX_TARGET_FUNNEL      3   /* Neal's funnel, notebooks/hamiltorch_log_prob_examples.ipynb c22 */
#define HMCX_TARGET_MLP         4   /* define_model_log_prob, samplers.py:1093-1201 (regression, MLP)   */

/* hmcx_mass_t.kind -- the inv_mass argument of sample()/leapfrog() (samplers.py:283-296) */
#define HMCX_MASS_NONE  0
#define HMCX_MASS_DIAG  1           /* inv_mass 1-D, samplers.py:296, :814, gibbs :201                 */
#define HMCX_MASS_FULL  2           /* inv_mass 2-D, samplers.py:294, :812, gibbs :199                 */

/* hmcx_rng_t.mode */
#define HMCX_RNG_INJECTED 0         /* host supplies the reference's own random stream (parity mode)   */
#define HMCX_RNG_PHILOX   1         /* in-kernel Philox4x32-10 keyed by (seed, chain, iteration)       */

/* Dense-stack Bayesian NN target == the closure define_model_log_prob builds (samplers.py:1093-1201) for a
 * Linear/activation stack with model_loss='regression', plus its data-split form (define_split_model_log_prob,
 * :1203-1258): split m owns data rows [split_begin[m], split_begin[m+1]) and divides the prior by prior_scale.
 * Flat parameter layout = util.flatten (util.py:121-136): per Linear the (out,in) row-major weight, then the bias. */
#define HMCX_MLP_MAX_LAYERS 8
#define HMCX_MLP_MAX_SPLITS 64
#define HMCX_ACT_NONE 0
#define HMCX_ACT_RELU 1
#define HMCX_ACT_TANH 2
#define HMCX_ACT_SIGMOID 3
#define HMCX_LOSS_REGRESSION            0   /* 'regression'                      samplers.py:1182-1184          */
#define HMCX_LOSS_BINARY                1   /* 'binary_class_linear_output'      :1170-1172 (BCE with logits)   */
#define HMCX_LOSS_MULTICLASS            2   /* 'multi_class_linear_output'       :1173-1177 (cross entropy, sum) */
#define HMCX_LOSS_MULTICLASS_LOGSOFTMAX 3   /* 'multi_class_log_softmax_output'  :1179-1180 (log-softmax output +
                                               nll_loss with its default MEAN reduction)                        */

typedef struct hmcx_mlp {
    int32_t num_layers;                            /* number of Linear layers                                  */
    int32_t widths[HMCX_MLP_MAX_LAYERS + 1];       /* n_0 (inputs) ... n_L (outputs)                           */
    int32_t activation[HMCX_MLP_MAX_LAYERS];       /* applied after layer l (last must be NONE)                */
    int32_t loss;
    float   tau_out;                               /* likelihood precision (samplers.py:1184)                  */
    float   prior_scale;                           /* l_prior / prior_scale (:1199); = num_splits when split   */
    /* Gaussian prior per parameter tensor i = (W_0, b_0, W_1, b_1, ...), constants carrying the reference's fp32
     * roundings of torch.distributions.Normal(0, tau_i^-1/2).log_prob (:1143, :1156):                           */
    float   prior_two_var[2 * HMCX_MLP_MAX_LAYERS];    /* 2*scale_i^2                                          */
    float   prior_log_scale[2 * HMCX_MLP_MAX_LAYERS];  /* log(scale_i)                                         */
    float   prior_grad_coef[2 * HMCX_MLP_MAX_LAYERS];  /* (1/prior_scale)/(2*scale_i^2): d prior/dw = -(coef*2w) */
    const float* x;                                /* [num_rows, n_0] device; NULL = sample the prior (:1160)    */
    const float* y;                                /* [num_rows, n_L] device (regression / binary) or [num_rows]
                                                      class indices stored as float (multi-class)              */
    int32_t num_rows;
    int32_t num_splits;                            /* M >= 1                                                    */
    int32_t split_begin[HMCX_MLP_MAX_SPLITS + 1];
    int32_t cluster_size;                          /* CTAs (SMs) cooperating on one chain: 0 = automatic, 1 / 2 / 4 =
                                                      pinned (bit-reproducibility across chain counts)          */
    int32_t tensor_cores;                          /* HMCX_MLP_TC_AUTO: first-layer GEMMs on tcgen05 (3xTF32, fp32-level
                                                      accuracy) when the stack is n0 -> 128 -> nL with n0 in
                                                      {16,32,48,64}, nL <= 4; HMCX_MLP_TC_OFF: fp32 SIMT tiles  */
    const float* x_packed;                         /* device buffer of hmcx_mlp_packed_x_bytes() filled by hmcx_mlp_pack_x():
                                                      x as ready-made tcgen05 operands (tf32 hi | lo, both GEMM layouts),
                                                      one bulk TMA copy per tile.  NULL: fp32 SIMT tiles               */
} hmcx_mlp_t;

typedef struct hmcx_target {
    int32_t kind;
    int32_t dim;                    /* D                                                               */
    const float* mean;              /* [D] device, GAUSS_DIAG / GAUSS_FULL (NULL = zeros)              */
    const float* inv_var;           /* [D] device, GAUSS_DIAG                                          */
    const float* prec;              /* [D,D] device row-major symmetric, GAUSS_FULL                    */
    float log_norm;                 /* additive constant of log p                                      */
    float funnel_inv_var_v;         /* FUNNEL: 1/sigma_v^2                                             */
    const hmcx_mlp_t* mlp;          /* MLP: HOST pointer to the network / data description             */
} hmcx_target_t;

typedef struct hmcx_mass {
    int32_t kind;
    const float* inv_mass;          /* DIAG: [D]; FULL: [D,D] row-major                                */
    const float* mass_factor;       /* DIAG: sqrt(1/inv_mass) [D] (gibbs :201);
                                       FULL: lower Cholesky factor of inverse(inv_mass) [D,D] (:199)   */
} hmcx_mass_t;

typedef struct hmcx_rng {
    int32_t mode;
    uint64_t seed;                  /* PHILOX key                                                      */
    uint64_t chain_offset;          /* PHILOX: global id of local chain 0 (multi-GPU sharding)         */
    const float* normals;           /* INJECTED: standard normals [iter_end-iter_begin, C, ld]         */
    const float* log_uniforms;      /* INJECTED: log(U) of the MH test [iter_end-iter_begin, C]        */
    const int32_t* perms;           /* INJECTED, SPLITTING_RAND only: randperm(M) per trajectory (:550)
                                       [iter_end-iter_begin, C, M]                                     */
    const float* uniforms;          /* INJECTED, RMHMC with jitter: the torch.rand(D) draws of fisher()
                                       (:115) in call order [iter_end-iter_begin, C, uniforms_per_iter, ld]   */
    int32_t uniforms_per_iter;      /* explicit integrator: 8*L+3 (gibbs, H, 8 per step, H_new)        */
} hmcx_rng_t;

/* Dual-averaging step-size adaptation ("HMC_NUTS"), samplers.py:629-674, per chain.
 * `table` holds, for t = 1..burn+1, the five Python-double constants the reference evaluates each call:
 *   {1-1/(t+10), 1/(t+10), sqrt(t)/0.05, t^-0.75, 1-t^-0.75}            (row-major [burn+1][5], device) */
typedef struct hmcx_nuts {
    int32_t enabled;
    double  desired_accept_rate;
    double  mu;                     /* float(log(10*eps0)) evaluated in fp32 as samplers.py:664        */
    const double* table;
    double* h_bar;                  /* [C] in/out, running H_t (starts at 0, samplers.py:938)          */
    double* eps_bar;                /* [C] in/out, running eps_bar (starts at 1, samplers.py:939)      */
    const float* eps_schedule;      /* optional [num_samples, C]: use THIS step size in iteration n instead of
                                       the adapted one ("teacher forcing": parity tests replay the reference's
                                       schedule because adaptation amplifies fp32 summation-order noise)       */
    float* eps_trace;               /* optional [C, num_samples]: the step size the kernel's own adaptation
                                       yields after iteration n (i.e. for iteration n+1)                       */
    double step_size_init;          /* the Python-double step size the run starts from (0 = not given; read even when
                                       enabled == 0).  The splitting integrators divide the DOUBLE step size before the
                                       product with the fp32 tensor rounds it (step_size/K_div, samplers.py:513, :558):
                                       while a chain's fp32 step size still equals (float)step_size_init the drift
                                       coefficient is (float)(step_size_init / K); an adapted step size is an fp32
                                       value in the reference too (:668) and is divided as such.                   */
} hmcx_nuts_t;

int         hmcx_abi_version(void);
const char* hmcx_status_string(int status);

/*
 * hmcx_leapfrog == samplers.leapfrog, plain-HMC branch (samplers.py:269-304) for C chains at once.  Targets GAUSS_ISO /
 * GAUSS_DIAG / GAUSS_FULL / FUNNEL, mass none / diagonal / full, any D (element-wise cases: the streaming HBM-roofline
 * kernel; coupled gradients or a full mass matrix: one CTA per chain with the state in shared memory).
 *   q_in, p_in   [C, ld]   start state (not modified)
 *   eps          [C]       per-chain step size
 *   q_out, p_out [C, ld]   state after L steps, p_out with the half-step correction of :302 applied
 *   q_traj, p_traj          optional (NULL to skip) [L, C, ld]: the L intermediate clones the reference returns
 *                           (ret_params / ret_momenta, :299-300; p_traj[L-1] is the corrected one)
 */
int hmcx_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass,
                  const float* q_in, const float* p_in, const float* eps,
                  int32_t C, int32_t ld, int32_t L,
                  float* q_out, float* p_out, float* q_traj, float* p_traj, void* stream);

/*
 * hmcx_hamiltonian == samplers.hamiltonian, sampler=HMC (samplers.py:779-815); same target / mass coverage as
 * hmcx_leapfrog.
 *   H_out [C]; flags_out [C] (optional) = 1 where log p is non-finite (the reference raises LogProbError, :783-785)
 */
int hmcx_hamiltonian(const hmcx_target_t* target, const hmcx_mass_t* mass,
                     const float* q, const float* p, int32_t C, int32_t ld,
                     float* H_out, uint8_t* flags_out, void* stream);

/*
 * hmcx_gibbs == samplers.gibbs, sampler=HMC (samplers.py:185-202), PHILOX mode only: p ~ N(0, M) for C chains
 * and iteration index `iter` (the same stream hmcx_hmc_run consumes for that iteration).
 */
int hmcx_gibbs(const hmcx_mass_t* mass, const hmcx_rng_t* rng, int32_t D, int32_t C, int32_t ld,
               int64_t iter, float* p_out, void* stream);

/*
 * hmcx_hmc_run == the sample() loop (samplers.py:954-1067) for sampler in {HMC, HMC_NUTS}: per iteration
 * gibbs -> hamiltonian -> leapfrog -> hamiltonian -> acceptance/MH -> bookkeeping (-> adaptation), as ONE
 * persistent kernel launch that advances iterations [iter_begin, iter_end) of C chains.
 *   q_init      [C, ld]  params_init (read only; needed again for the reference's first-post-burn-reject quirk)
 *   q_cur       [C, ld]  in/out: current `params`; must equal q_init when iter_begin == 0
 *   eps         [C]      in/out: per-chain step size (changes only when nuts->enabled)
 *   samples_out [C, num_samples-burn, ld]  slot 0 = params_init (:959), slot n-burn = iteration n > burn
 *   accept_out / diverged_out  optional [C, num_samples] (uint8); ham_out optional [C, num_samples, 2] = (H_old, H_new)
 *   num_rejected optional [C] int32 in/out counter (:961, :1016, :1046)
 *   workspace    hmcx_hmc_workspace_bytes() bytes of device scratch (NULL when that is 0).  For element-wise targets with
 *                ld <= 4096 these are C floats that carry log p(q_cur) from one window of iterations to the next: pass the
 *                SAME buffer to the launches [0, a), [a, b), ... of a run and they reproduce the single launch [0, S) bit for
 *                bit; with NULL a window recomputes log p(q_cur) (a reduction that associates differently from the loop's:
 *                H_old of its first iteration may differ in the last bit)
 *   tuning       0 = automatic register geometry (one float4 per thread for D <= 2560, two above); 1 = one float4 per
 *                thread; 2 / 4 = that many float4 groups per thread; 21 / 22 = one / two float2 groups per thread
 *                (tests and tuning sweeps; element-wise state and the random stream never depend on it)
 */
int hmcx_hmc_run(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                 const hmcx_nuts_t* nuts,
                 const float* q_init, float* q_cur, float* eps,
                 int32_t C, int32_t ld, int32_t L, int32_t num_samples, int32_t burn,
                 int32_t iter_begin, int32_t iter_end,
                 float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                 int32_t* num_rejected, int32_t tuning, float* workspace, void* stream);

/* Scratch hmcx_hmc_run needs in `workspace` (device bytes; 0 = may pass NULL).  Element-wise targets with D > 4096 are
 * advanced by a streamed form of the kernel that parks each proposal in a caller-provided (C, ld) buffer. */
size_t hmcx_hmc_workspace_bytes(const hmcx_target_t* target, const hmcx_mass_t* mass, int32_t C, int32_t ld);

/* sampler=RMHMC configuration (samplers.py:850 arguments that only this sampler reads) */
typedef struct hmcx_rmhmc {
    int32_t integrator;             /* Integrator enum value: 1 EXPLICIT (:389-462), 2 IMPLICIT (:305-387)         */
    int32_t metric;                 /* Metric enum value: 1 HESSIAN, 2 SOFTABS (:116-122), 3 JACOBIAN_DIAG (:100-106,
                                       hmcx_rmhmc_run only: the metric diag((d log p/d theta_i)^2) is never constant)  */
    float   softabs_const;          /* alpha                                                                      */
    float   jitter;                 /* scale of the uniform diagonal jitter (:113-115); < 0 = None                */
    float   pi_term;                /* D*log(2*pi) evaluated in fp32 as :712                                      */
    float   cos_2we, sin_2we;       /* cos/sin(2*explicit_binding_const*step_size) in fp32 as :435-436            */
    float   fixed_point_threshold;  /* implicit: :337, :356                                                       */
    int32_t fixed_point_max_iterations;
    int32_t jitter_max_tries;       /* NaN-gradient retries before LogProbError (:402-410)                        */
} hmcx_rmhmc_t;

/* integrators of the HMC family (samplers.py:269, :494, :548, :575) */
#define HMCX_SCHEME_PLAIN       0   /* plain leapfrog on the whole potential (sample_model)             */
#define HMCX_SCHEME_SPLIT_SYM   1   /* Integrator.SPLITTING       (:494-547)                             */
#define HMCX_SCHEME_SPLIT_RAND  2   /* Integrator.SPLITTING_RAND  (:548-571)                             */
#define HMCX_SCHEME_SPLIT_KMID  3   /* Integrator.SPLITTING_KMID  (:575-601)                             */

/*
 * hmcx_split_run == the sample() loop for a data-split potential U = sum_m U_m (sample_split_model, samplers.py:1364
 * -> sample with integrator in {SPLITTING, SPLITTING_RAND, SPLITTING_KMID}), and with HMCX_SCHEME_PLAIN the same
 * loop for the un-split Bayesian NN (sample_model, :1261).  Target must be HMCX_TARGET_MLP.  Arguments as
 * hmcx_hmc_run; the Hamiltonian sums the split log-probs (:787-796).
 */
int hmcx_split_run(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                   const hmcx_nuts_t* nuts, int32_t scheme,
                   const float* q_init, float* q_cur, float* eps,
                   int32_t C, int32_t ld, int32_t L, int32_t num_samples, int32_t burn,
                   int32_t iter_begin, int32_t iter_end,
                   float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                   int32_t* num_rejected, void* stream);

/*
 * hmcx_rmhmc_run == the sample() loop for sampler=RMHMC (samplers.py:965-1067 with gibbs :183-184, rm_hamiltonian
 * :677-736, fisher :69-127, explicit :389-462 / implicit :305-387 leapfrog).  Targets: FUNNEL, GAUSS_ISO, GAUSS_DIAG,
 * GAUSS_FULL (closed-form Hessian and third-derivative contraction), any jitter, metrics HESSIAN / SOFTABS /
 * JACOBIAN_DIAG.  dim <= 16: one thread per chain (dim == 2 with the explicit integrator -- BASELINE config 3 -- a
 * pair of warps per 32 chains that evaluates dH/dtheta and dH/dp concurrently); 16 < dim <= 64: one CTA per chain, the
 * metric assembled, eigen-decomposed (parallel-order Jacobi) and solved in shared memory.  Arguments as hmcx_hmc_run
 * (eps is read-only: the reference never adapts the step size of RMHMC).
 */
int hmcx_rmhmc_run(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng,
                   const float* q_init, float* q_cur, const float* eps,
                   int32_t C, int32_t ld, int32_t L, int32_t num_samples, int32_t burn,
                   int32_t iter_begin, int32_t iter_end,
                   float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                   int32_t* num_rejected, void* stream);

/*
 * hmcx_rmhmc_leapfrog == samplers.leapfrog with sampler=RMHMC called on its own (explicit :389-462, implicit :305-387):
 * L steps from (q_in, p_in), same targets / metrics as hmcx_rmhmc_run, dim <= 64.
 *   q_traj, p_traj   [L, C, ld]  theta and p after every step (the reference's ret_params / ret_momenta lists)
 *   q_copy_out, p_copy_out  optional [C, ld]: the explicit integrator's params_copy / momentum_copy after the last step
 *                    (second elements of the pairs it returns, :462)
 *   flags_out        [C] (uint8) 1 where the reference raises LogProbError inside the trajectory (outputs of that chain
 *                    from the failing step on are unspecified)
 * rng: jitter draws of the fisher() calls in call order from row 0 (INJECTED: uniforms [1, C, J, ld]) or Philox with
 * iteration index 0; normals / log_uniforms are not read.
 */
int hmcx_rmhmc_leapfrog(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng,
                        const float* q_in, const float* p_in, const float* eps, int32_t C, int32_t ld, int32_t L,
                        float* q_traj, float* p_traj, float* q_copy_out, float* p_copy_out, uint8_t* flags_out,
                        void* stream);

/*
 * hmcx_rmhmc_hamiltonian == samplers.hamiltonian with sampler=RMHMC (:817-829) == rm_hamiltonian (:677-736):
 * H_out[c] = -log p + 0.5 D log 2pi + 0.5 log det G + 0.5 p.G^-1 p (the caller doubles it for the explicit integrator's
 * augmented form, :822); flags_out [C] = 1 where the reference raises LogProbError.  One jitter row (row 0).
 */
int hmcx_rmhmc_hamiltonian(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_rng_t* rng,
                           const float* q, const float* p, int32_t C, int32_t ld, float* H_out, uint8_t* flags_out,
                           void* stream);

/*
 * hmcx_grad_log_prob == collect_gradients(log_prob_func(params), params) (samplers.py:33-66, :270-278) for C
 * parameter vectors: grad_out[c] = d log p_split(q[c]) / dq.  split = -1 sums all splits.  Any target kind.
 * log_prob_out optional [C].
 */
int hmcx_grad_log_prob(const hmcx_target_t* target, const float* q, int32_t C, int32_t ld, int32_t split,
                       float* grad_out, float* log_prob_out, void* stream);

/*
 * hmcx_mlp_predict == predict_model (samplers.py:1468-1562): forward pass of every sample over the target's data.
 *   samples [S, ld];  pred_out [S, num_rows, n_L];  log_prob_out [S] = ll + prior/prior_scale (:1197)
 */
int hmcx_mlp_predict(const hmcx_target_t* target, const float* samples, int32_t S, int32_t ld,
                     float* pred_out, float* log_prob_out, void* stream);

/*
 * hmcx_split_leapfrog == samplers.leapfrog called directly with Integrator.SPLITTING / SPLITTING_RAND / SPLITTING_KMID on the
 * list define_split_model_log_prob returns (samplers.py:494-603): L steps from (q_in, p_in) [C, ld]; q_traj / p_traj
 * [L, C, ld] receive params and momentum after EVERY step (the reference's ret_params / ret_momenta lists).  eps [C] =
 * (float)step_size per chain; step_size is the Python double the drifts divide before rounding (:513, :558).
 * SPLITTING_RAND takes its one randperm(M) per call (:550) from rng->perms [C, M] (INJECTED) or from the Philox PERM stream.
 */
int hmcx_split_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng, int32_t scheme,
                        double step_size, const float* q_in, const float* p_in, float* eps, int32_t C, int32_t ld,
                        int32_t L, float* q_traj, float* p_traj, void* stream);

/*
 * Packed X operands of the BNN tensor-core path (hmcx_mlp_t.x_packed).  The data matrix of define_model_log_prob
 * (samplers.py:1093-1201) never changes during a run, so its tf32 hi / lo split and the two tcgen05 operand layouts
 * (forward: rows x inputs, backward: inputs x rows) are built once per target:
 *   hmcx_mlp_packed_x_bytes   size of the buffer (0: this stack has no tensor-core form -- leave x_packed NULL)
 *   hmcx_mlp_pack_x           fills it from target->mlp->x on `stream` (x_packed itself is not read)
 */
size_t hmcx_mlp_packed_x_bytes(const hmcx_target_t* target);
int hmcx_mlp_pack_x(const hmcx_target_t* target, float* packed_out, void* stream);

/*
 * hmcx_gemm_nt_tf32x3: D[M,N] = A[M,K] . B[N,K]^T on the 5th-generation tensor cores (tcgen05.mma kind::tf32 with
 * 3xTF32 split operands => fp32-accurate, fp32 accumulation in tensor memory).  Row-major fp32 device arrays;
 * M, N multiples of 128, K a multiple of 32.  The dense contraction behind full-covariance targets / full mass
 * matrices at large D (grad log p of ALL chains = -(Q - mu) P: M = chains, N = K = D; samplers.py:294, :812).
 */
int hmcx_gemm_nt_tf32x3(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K, void* stream);

/*
 * Constant-metric RMHMC on the tensor cores.  For Gaussian targets without jitter the metric of samplers.py:69-127
 * (G = -Hessian, or its softabs map V diag(lambda coth(alpha lambda)) V^T) is one matrix for every chain and every
 * point, so the flows of the explicit (:427-458) and implicit (:363-386) integrators are (chains x D).(D x D)
 * contractions: dH/dp = G^-1 p (the metric solve of cholesky_inverse, :130-149) and dH/dtheta = P (theta - mu).
 * The caller evaluates, once, with the reference's own torch ops:
 *   metric_inv  [D,D]  G^-1        metric_chol [D,D]  lower Cholesky factor of G (gibbs :183-184)
 *   log_det            log det G   (sum log lambda~ for SOFTABS :726, slogdet for HESSIAN :728)
 */
typedef struct hmcx_const_metric {
    const float* metric_inv;
    const float* metric_chol;
    float log_det;
} hmcx_const_metric_t;

size_t hmcx_rmhmc_dense_workspace_bytes(int32_t C, int32_t D);

/*
 * hmcx_rmhmc_dense_run == the sample() loop for sampler=RMHMC (as hmcx_rmhmc_run) for targets GAUSS_ISO / GAUSS_DIAG /
 * GAUSS_FULL of ANY dimension with cfg->jitter < 0 (None): every flow is a tcgen05 GEMM over all chains (3xTF32,
 * operands packed for 1-D bulk TMA), 8 per explicit leapfrog step.  workspace: hmcx_rmhmc_dense_workspace_bytes().
 * D <= 128 (and ld <= D rounded up to 32): the whole run is ONE persistent launch instead (hmcx_flow.cu: the matrices in
 * shared memory, a warp owns 1-4 chains, exact fp32 FMAs; the workspace is then unused; environment HMCX_FLOW_SMALL=0
 * keeps the GEMM path).  The same holds for hmcx_hmc_run with a dense precision or a 2-D inv_mass at 16 < D <= 128.
 */
int hmcx_rmhmc_dense_run(const hmcx_target_t* target, const hmcx_rmhmc_t* cfg, const hmcx_const_metric_t* metric,
                         const hmcx_rng_t* rng, const float* q_init, float* q_cur, const float* eps,
                         int32_t C, int32_t ld, int32_t L, int32_t num_samples, int32_t burn,
                         int32_t iter_begin, int32_t iter_end,
                         float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                         int32_t* num_rejected, float* workspace, void* stream);

/*
 * Sample sink -- what consumes the retained samples (the store_on_GPU=False contract of samplers.py:1008-1012 and the
 * step after the path), for runs whose C*S*D exceeds what the caller wants to keep:
 *   thin   keep every thin-th post-burn iteration: samples_out is [C, 1 + (num_samples-burn-1)/thin, ld], slot 0 =
 *          params_init, slot j = the chain state after iteration burn + j*thin  (thin = 1: the reference's list)
 *   sum, sumsq   optional [C, ld] in/out accumulators: running sum / sum of squares of the chain state over EVERY
 *          iteration n > burn (= elements 1.. of the reference's returned list), so posterior means and variances
 *          need no sample storage at all (samples_out may then be NULL).  Accumulated in registers with Neumaier
 *          compensation (the rounding of x*x included): relative error ~ n*eps^2 after n iterations (eps = 2^-24)
 *          instead of the ~ n*eps of a plain fp32 running sum:
 *   sum_lo, sumsq_lo   optional [C, ld] in/out: the compensation terms; the sums are hi + lo (combine in fp64: var =
 *          E[x^2] - mean^2 then holds for |mean| >> std; measured 1.6e-7 on a variance of 1 at mean 100, n = 2e4, where
 *          the plain sum is off by percents).  Without them the fp32 rounding of hi + lo is stored in sum / sumsq.
 * samples_out may point to device-mapped pinned HOST memory: the kernel's retained-row stores are coalesced 16-byte
 * streaming stores (st.global.cs), which is how samples leave the GPU while the chains keep running.
 */
typedef struct hmcx_sink {
    int32_t thin;
    float*  sum;
    float*  sumsq;
    float*  sum_lo;
    float*  sumsq_lo;
} hmcx_sink_t;

/* hmcx_hmc_run with a sample sink (sink == NULL or {1, NULL, NULL, NULL, NULL}: identical to hmcx_hmc_run).  Element-wise
 * targets (GAUSS_ISO / GAUSS_DIAG, mass none / diagonal, ld <= 4096); other combinations return HMCX_ERR_UNSUPPORTED
 * when the sink asks for thinning or moments. */
int hmcx_hmc_run_sink(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng,
                      const hmcx_nuts_t* nuts,
                      const float* q_init, float* q_cur, float* eps,
                      int32_t C, int32_t ld, int32_t L, int32_t num_samples, int32_t burn,
                      int32_t iter_begin, int32_t iter_end,
                      float* samples_out, uint8_t* accept_out, uint8_t* diverged_out, float* ham_out,
                      int32_t* num_rejected, int32_t tuning, float* workspace, const hmcx_sink_t* sink, void* stream);

/*
 * hmcx_copy_rows_async: `height` rows of `width` bytes from `src` (row pitch `spitch`) to `dst` (row pitch `dpitch`), either
 * side device or PINNED host memory, enqueued on `stream` (cudaMemcpy2DAsync, cudaMemcpyDefault).  The delivery half of a
 * windowed run: hmcx_hmc_run over iterations [a, b) on one stream, then the window's sample slots -- the same columns of
 * every chain's [num_samples-burn, ld] block -- leave for the host on a second stream through the copy engine while the next
 * window computes (the reference's store_on_GPU=False, samplers.py:1008-1012, without stalling the chains on PCIe).
 */
int hmcx_copy_rows_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMCX_H */
