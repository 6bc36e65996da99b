// hmcx_coupled.cu -- stand-alone samplers.leapfrog (plain HMC branch, samplers.py:269-304) and samplers.hamiltonian
// (sampler=HMC, :779-815) for everything that is NOT element-wise: GaussianFull and Funnel targets (gradients couple the
// coordinates) and the full (2-D) inv_mass of :294 / :812, at any dimension.
//
// These are the reference's utility entry points (its own reversibility test calls leapfrog directly); the sample() loop
// itself runs these cases in hmc_small_kernel (D <= 16) or on the tensor cores (hmcx_tc.cu).  One CTA owns one chain:
// q, p and two work vectors live in shared memory, a D x D matrix (precision or inv_mass) is applied by warps that each
// take whole rows -- lanes stride the row, so global reads are coalesced -- and reduce with shuffles.  Element-wise
// arithmetic keeps the reference's operation order (mul / add rounded separately); the matrix-vector sums accumulate in
// fp32 like torch.mv up to summation order.
#include "hmcx_common.cuh"

namespace hmcx {

struct CoupledArgs {
    int kind, mk, D, ld, C;
    const float* mean;
    const float* ivar;
    const float* prec;
    const float* im;          // inv_mass: [D] (diag) or [D, D] (full)
    float log_norm, inv_var_v;
};

constexpr int CP_THREADS = 256;

// out[i] = sum_j M[i, j] * (x[j] - shift[j])   (shift may be null); all threads call, result visible after the barrier
__device__ __forceinline__ void cp_matvec(const float* __restrict__ M, const float* x, const float* __restrict__ shift,
                                          float* out, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int i = warp; i < D; i += nwarp) {
        const float* row = M + (size_t)i * D;
        float s = 0.0f;
        for (int j = lane; j < D; j += 32) s += row[j] * (shift ? sub(x[j], shift[j]) : x[j]);
        s = warp_sum(s);
        if (lane == 0) out[i] = s;
    }
    __syncthreads();
}

// block-wide sum of f(i) over i in [lo, D); every thread returns the same bits
template <typename F>
__device__ __forceinline__ float cp_sum(int lo, int D, float* sred, F f) {
    float v[1] = {0.0f};
    for (int i = lo + (int)threadIdx.x; i < D; i += blockDim.x) v[0] = add(v[0], f(i));
    __syncthreads();                       // sred may still be read by a previous call
    block_sum<1>(v, sred);
    return v[0];
}

// g = grad log p(q) (all D entries, visible after return); `w` is a D-float work vector
__device__ __forceinline__ void cp_grad(const CoupledArgs& a, const float* q, float* g, float* sred) {
    const int D = a.D;
    if (a.kind == HMCX_TARGET_GAUSS_FULL) {
        cp_matvec(a.prec, q, a.mean, g, D);
        for (int i = threadIdx.x; i < D; i += blockDim.x) g[i] = -g[i];
    } else if (a.kind == HMCX_TARGET_FUNNEL) {                               // targets.Funnel.grad
        const float s = cp_sum(1, D, sred, [&](int i) { return mul(q[i], q[i]); });
        const float v = q[0], E = expf(v);
        for (int i = 1 + (int)threadIdx.x; i < D; i += blockDim.x) g[i] = -mul(E, q[i]);
        if (threadIdx.x == 0)
            g[0] = sub(add(-mul(a.inv_var_v, v), mul(0.5f, (float)(D - 1))), mul(mul(0.5f, E), s));
    } else {
        for (int i = threadIdx.x; i < D; i += blockDim.x)
            g[i] = (a.kind == HMCX_TARGET_GAUSS_ISO) ? -q[i] : -mul(a.ivar[i], sub(q[i], a.mean ? a.mean[i] : 0.0f));
    }
    __syncthreads();
}

__device__ __forceinline__ float cp_log_prob(const CoupledArgs& a, const float* q, float* w, float* sred) {
    const int D = a.D;
    if (a.kind == HMCX_TARGET_GAUSS_FULL) {
        cp_matvec(a.prec, q, a.mean, w, D);
        const float s = cp_sum(0, D, sred, [&](int i) { return mul(sub(q[i], a.mean ? a.mean[i] : 0.0f), w[i]); });
        return add(mul(-0.5f, s), a.log_norm);
    }
    if (a.kind == HMCX_TARGET_FUNNEL) {                                      // targets.Funnel.__call__
        const float s = cp_sum(1, D, sred, [&](int i) { return mul(q[i], q[i]); });
        const float v = q[0];
        const float t1 = mul((float)(-0.5 * (double)a.inv_var_v), mul(v, v));
        const float t2 = mul(0.5f * (float)(D - 1), v);
        const float t3 = mul(mul(0.5f, expf(v)), s);
        return add(sub(add(t1, t2), t3), a.log_norm);
    }
    const float s = cp_sum(0, D, sred, [&](int i) {
        if (a.kind == HMCX_TARGET_GAUSS_ISO) return mul(q[i], q[i]);
        const float y = sub(q[i], a.mean ? a.mean[i] : 0.0f);
        return mul(mul(y, y), a.ivar[i]);
    });
    return add(mul(-0.5f, s), a.log_norm);
}

// w = M^-1 p
__device__ __forceinline__ void cp_minv(const CoupledArgs& a, const float* p, float* w) {
    if (a.mk == HMCX_MASS_FULL) { cp_matvec(a.im, p, nullptr, w, a.D); return; }
    for (int i = threadIdx.x; i < a.D; i += blockDim.x) w[i] = (a.mk == HMCX_MASS_DIAG) ? mul(a.im[i], p[i]) : p[i];
    __syncthreads();
}

__global__ void __launch_bounds__(CP_THREADS)
coupled_leapfrog_kernel(const CoupledArgs a, const float* __restrict__ q_in, const float* __restrict__ p_in,
                        const float* __restrict__ eps_in, int L, float* __restrict__ q_out, float* __restrict__ p_out,
                        float* __restrict__ q_traj, float* __restrict__ p_traj) {
    extern __shared__ float sm[];
    __shared__ float sred[32];
    const int D = a.D, ld = a.ld, c = blockIdx.x;
    float* q = sm; float* p = sm + D; float* g = sm + 2 * D; float* w = sm + 3 * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { q[i] = q_in[(size_t)c * ld + i]; p[i] = p_in[(size_t)c * ld + i]; }
    __syncthreads();
    const float eps = eps_in[c], half = mul(0.5f, eps);
    const size_t traj_stride = (size_t)a.C * ld;
    cp_grad(a, q, g, sred);
    for (int i = threadIdx.x; i < D; i += blockDim.x) p[i] = add(p[i], mul(half, g[i]));                 // :281
    __syncthreads();
    for (int l = 0; l < L; ++l) {
        cp_minv(a, p, w);
        for (int i = threadIdx.x; i < D; i += blockDim.x)                                                  // :284/:294/:296
            q[i] = add(q[i], (a.mk == HMCX_MASS_DIAG) ? mul(mul(eps, a.im[i]), p[i]) : mul(eps, w[i]));
        __syncthreads();
        cp_grad(a, q, g, sred);
        const bool last = l + 1 == L;
        for (int i = threadIdx.x; i < ld; i += blockDim.x) {
            float pn = 0.0f, qn = 0.0f;
            if (i < D) {
                pn = add(p[i], mul(eps, g[i]));                                                            // :298
                if (last) pn = sub(pn, mul(half, g[i]));                                                   // :302
                p[i] = pn; qn = q[i];
            }
            if (q_traj) {                                                                                  // :299-300
                q_traj[(size_t)l * traj_stride + (size_t)c * ld + i] = qn;
                p_traj[(size_t)l * traj_stride + (size_t)c * ld + i] = pn;
            }
            if (last) { q_out[(size_t)c * ld + i] = qn; p_out[(size_t)c * ld + i] = pn; }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(CP_THREADS)
coupled_hamiltonian_kernel(const CoupledArgs a, const float* __restrict__ q_in, const float* __restrict__ p_in,
                           float* __restrict__ H, uint8_t* __restrict__ flags) {
    extern __shared__ float sm[];
    __shared__ float sred[32];
    const int D = a.D, ld = a.ld, c = blockIdx.x;
    float* q = sm; float* p = sm + D; float* w = sm + 2 * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { q[i] = q_in[(size_t)c * ld + i]; p[i] = p_in[(size_t)c * ld + i]; }
    __syncthreads();
    const float lp = cp_log_prob(a, q, w, sred);
    __syncthreads();
    cp_minv(a, p, w);
    const float kin = cp_sum(0, D, sred, [&](int i) { return mul(p[i], w[i]); });
    if (threadIdx.x == 0) {
        H[c] = add(-lp, mul(0.5f, kin));                                                                   // :815
        if (flags) flags[c] = finite_f(lp) ? 0 : 1;                                                        // :783-785
    }
}

static int fill_coupled(const hmcx_target_t* target, const hmcx_mass_t* mass, int C, int ld, CoupledArgs& a) {
    if (!target || C < 1 || target->dim < 1 || ld < target->dim || (ld & 3)) return HMCX_ERR_INVALID_ARG;
    a = {};
    a.kind = target->kind; a.D = target->dim; a.ld = ld; a.C = C;
    a.mean = target->mean; a.ivar = target->inv_var; a.prec = target->prec;
    a.log_norm = target->log_norm; a.inv_var_v = target->funnel_inv_var_v;
    switch (a.kind) {
        case HMCX_TARGET_GAUSS_ISO: break;
        case HMCX_TARGET_GAUSS_DIAG: if (!a.ivar) return HMCX_ERR_INVALID_ARG; break;
        case HMCX_TARGET_GAUSS_FULL: if (!a.prec) return HMCX_ERR_INVALID_ARG; break;
        case HMCX_TARGET_FUNNEL: if (a.D < 2) return HMCX_ERR_INVALID_ARG; break;
        default: return HMCX_ERR_UNSUPPORTED;
    }
    a.mk = mass ? mass->kind : HMCX_MASS_NONE;
    if (a.mk != HMCX_MASS_NONE && a.mk != HMCX_MASS_DIAG && a.mk != HMCX_MASS_FULL) return HMCX_ERR_INVALID_ARG;
    if (a.mk != HMCX_MASS_NONE && !mass->inv_mass) return HMCX_ERR_INVALID_ARG;
    a.im = mass ? mass->inv_mass : nullptr;
    return HMCX_OK;
}

int coupled_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass, const float* q_in, const float* p_in,
                     const float* eps, int C, int ld, int L, float* q_out, float* p_out, float* q_traj, float* p_traj,
                     cudaStream_t st) {
    CoupledArgs a;
    const int rc = fill_coupled(target, mass, C, ld, a);
    if (rc != HMCX_OK) return rc;
    if (!q_in || !p_in || !eps || !q_out || !p_out || L < 1 || ((q_traj == nullptr) != (p_traj == nullptr)))
        return HMCX_ERR_INVALID_ARG;
    const size_t smem = (size_t)4 * a.D * sizeof(float);
    if (smem > 200 * 1024) return HMCX_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(coupled_leapfrog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return HMCX_ERR_CUDA;
    }
    coupled_leapfrog_kernel<<<C, CP_THREADS, smem, st>>>(a, q_in, p_in, eps, L, q_out, p_out, q_traj, p_traj);
    return cudaGetLastError() == cudaSuccess ? HMCX_OK : HMCX_ERR_CUDA;
}

int coupled_hamiltonian(const hmcx_target_t* target, const hmcx_mass_t* mass, const float* q, const float* p, int C,
                        int ld, float* H, uint8_t* flags, cudaStream_t st) {
    CoupledArgs a;
    const int rc = fill_coupled(target, mass, C, ld, a);
    if (rc != HMCX_OK) return rc;
    if (!q || !p || !H) return HMCX_ERR_INVALID_ARG;
    const size_t smem = (size_t)3 * a.D * sizeof(float);
    if (smem > 200 * 1024) return HMCX_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(coupled_hamiltonian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return HMCX_ERR_CUDA;
    }
    coupled_hamiltonian_kernel<<<C, CP_THREADS, smem, st>>>(a, q, p, H, flags);
    return cudaGetLastError() == cudaSuccess ? HMCX_OK : HMCX_ERR_CUDA;
}

}  // namespace hmcx
