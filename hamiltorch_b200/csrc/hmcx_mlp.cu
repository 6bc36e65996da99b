// hmcx_mlp.cu -- Bayesian dense-stack (MLP) HMC on sm_100a: the BNN rows of the hot path.
//
//   define_model_log_prob / define_split_model_log_prob   samplers.py:1093-1258  -> mlp_log_prob(), mlp_grad_split()
//   leapfrog SPLITTING / SPLITTING_RAND / SPLITTING_KMID  samplers.py:465-603    -> trajectory in mlp_run_kernel
//   sample() loop around them (sample_model / sample_split_model, :1261-1466)    -> mlp_run_kernel (persistent)
//   predict_model                                          samplers.py:1468-1562  -> mlp_predict_kernel
//   collect_gradients (autograd) is replaced by a hand-written backward pass      -> mlp_grad_split()
//
// Design (round 1, fp32 SIMT; the tensor-core batched-over-chains form is the next step):
// one CTA of 256 threads owns one chain for the whole run.  The chain's flat parameter vector q, its momentum p and
// the split gradient g live in SHARED MEMORY in the reference's flat layout (util.py:121-136), so every weight is
// read from smem by the GEMM loops and the leapfrog kick/drift are conflict-free element-wise passes.  A gradient
// evaluation streams the split's data rows through in tiles of 32 rows: forward (Linear+activation per layer,
// register micro-tiles 4x4 with interleaved columns and a per-lane k-rotation that makes the strided weight reads
// bank-conflict-free), loss gradient, backward (dW += dZ^T A, db, dA = dZ W) accumulating straight into g.
// Multiply-adds inside the GEMM loops use FMA: these sums have no bit-parity counterpart in the reference (its
// sgemm order is unknowable); everything element-wise (kicks, drifts, prior, Hamiltonian assembly, MH) keeps the
// reference's separately-rounded fp32 operation order.
#include "hmcx_common.cuh"
#include "hmcx_umma.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace hmcx {

constexpr int MLP_THREADS = 512;
constexpr int MLP_T_MAX = 64;             // data rows per tile: 64 when the tile buffers fit next to q,p,g, else 32 / 16
// (rows of a 4x4 micro-tile are rg, rg+T/4, rg+2T/4, rg+3T/4)

struct MlpDev {
    int L, D, Dp, N, M, has_data, loss;
    int n[HMCX_MLP_MAX_LAYERS + 1], act[HMCX_MLP_MAX_LAYERS];
    int woff[HMCX_MLP_MAX_LAYERS], boff[HMCX_MLP_MAX_LAYERS];
    int aoff[HMCX_MLP_MAX_LAYERS + 1];    // smem offsets (floats, relative to the tile area) of A[l] (T x n[l])
    int dzoff[2];                         // two delta buffers (T x maxw)
    int tile_floats;
    int tile_base;                        // float offset of the tile area in dynamic shared memory (after the state vectors)
    int T;                                // rows per tile (multiple of 4)
    int tc;                               // 1: the first layer's GEMMs run on tcgen05 (layout below), 0: SIMT tiles
    int tc_f0, tc_f1, tc_b, tc_part;      // tile-area offsets (floats): two forward X operand buffers, the backward one, partials
    int tc_yraw;                          // cp.async landing buffer of the tile's targets
    const float* xp;                      // packed X operands (hmcx_mlp_pack_x): per tile [fwd hi|lo (128 n0) | bwd hi|lo (128 n0)]
    int tb[HMCX_MLP_MAX_SPLITS + 1];      // first packed tile of split s (tiles of a split start at its first row)
    int flat_base;                        // first packed tile of the all-rows tiling (rows 0, 64, ...)
    float tau_out, prior_scale, c_ll;     // c_ll = fp32(-0.5*tau_out) (regression, :1184) or fp32(-tau_out) (:1172-1180)
    float two_var[2 * HMCX_MLP_MAX_LAYERS], log_scale[2 * HMCX_MLP_MAX_LAYERS], gcoef[2 * HMCX_MLP_MAX_LAYERS];
    const float* x;
    const float* y;
    int sb[HMCX_MLP_MAX_SPLITS + 1];
};

__device__ __forceinline__ float act_fwd(float z, int a) {
    if (a == HMCX_ACT_RELU) return z > 0.0f ? z : 0.0f;
    if (a == HMCX_ACT_TANH) return tanhf(z);
    if (a == HMCX_ACT_SIGMOID) return 1.0f / (1.0f + expf(-z));
    return z;
}
// derivative expressed through the activation's OUTPUT (what the backward pass has at hand)
__device__ __forceinline__ float act_bwd(float aout, int a) {
    if (a == HMCX_ACT_RELU) return aout > 0.0f ? 1.0f : 0.0f;
    if (a == HMCX_ACT_TANH) return 1.0f - aout * aout;
    if (a == HMCX_ACT_SIGMOID) return aout * (1.0f - aout);
    return 1.0f;
}

// ---- tile primitives (all 256 threads; callers place the __syncthreads) ---------------------------------------
__device__ __forceinline__ void mlp_load_x(const MlpDev& m, float* A0, int r0, int cnt) {
    const int n0 = m.n[0];
    for (int i = threadIdx.x; i < m.T * n0; i += MLP_THREADS)
        A0[i] = (i < cnt * n0) ? __ldg(m.x + (size_t)r0 * n0 + i) : 0.0f;
}

// ---- "thin" layers (few outputs, e.g. the scalar regression head): split the reduction over a lane group ---------
// S = lanes per output (power of two, S | 32); each group's lanes stride the reduction index, then shuffle-reduce.
__device__ __forceinline__ int thin_group(int outputs) {
    int s = 32;
    while (s > 1 && outputs * s > MLP_THREADS) s >>= 1;
    return s;
}
__device__ __forceinline__ float group_sum(float v, int S) {
    for (int o = S >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void mlp_linear_fwd_thin(const float* Ain, const float* W, const float* b, float* Aout,
                                                    int n_in, int n_out, int act, int T) {
    const int outputs = T * n_out, S = thin_group(outputs);
    for (int base = 0; base < outputs; base += MLP_THREADS / S) {
        const int o = base + threadIdx.x / S, s = threadIdx.x % S;
        const bool live = o < outputs;
        const int r = live ? o / n_out : 0, j = live ? o % n_out : 0;
        float acc = 0.0f;
        for (int k = s; k < n_in; k += S) acc = fmaf(Ain[r * n_in + k], W[j * n_in + k], acc);
        acc = group_sum(acc, S);
        if (live && s == 0) Aout[r * n_out + j] = act_fwd(acc + b[j], act);
    }
}

// Aout[T x n_out] = act(Ain[T x n_in] . W^T + b),  W row-major (n_out, n_in) in shared memory.
// 4x4 register micro-tiles (rows rg+8i, columns cg+ncg*jj); the reduction index is rotated per lane so that the
// strided weight reads are bank-conflict-free; with n_in % 4 == 0 and 16-byte aligned rows the loads are float4.
__device__ __forceinline__ void mlp_linear_fwd(const float* Ain, const float* W, const float* b, float* Aout,
                                               int n_in, int n_out, int act, int T) {
    const int ncg = (n_out + 3) >> 2, RG = T >> 2;
    if (RG * ncg * 4 <= MLP_THREADS) { mlp_linear_fwd_thin(Ain, W, b, Aout, n_in, n_out, act, T); return; }
    const bool vec = ((n_in & 3) == 0) && ((((size_t)W) & 15) == 0) && ((((size_t)Ain) & 15) == 0);
    for (int tile = threadIdx.x; tile < RG * ncg; tile += MLP_THREADS) {
        const int cg = tile % ncg, rg = tile / ncg;
        int jc[4];
        float acc[4][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = cg + jj * ncg;
            jc[jj] = j < n_out ? j : n_out - 1;
            const float bj = b[jc[jj]];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][jj] = bj;
        }
        if (vec) {
            const int n4 = n_in >> 2, rot = cg % n4;          // rotation in units of float4
            const float4* ap[4];
            const float4* wp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ap[i] = reinterpret_cast<const float4*>(Ain) + (rg + i * RG) * n4;
                wp[i] = reinterpret_cast<const float4*>(W) + jc[i] * n4;
            }
            auto body = [&](int kk) {
                float4 a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = ap[i][kk]; w[i] = wp[i][kk]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i].x, w[jj].x, acc[i][jj]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i].y, w[jj].y, acc[i][jj]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i].z, w[jj].z, acc[i][jj]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i].w, w[jj].w, acc[i][jj]);
            };
            int kk = rot;                                     // uniform trip count: no divergence across lanes
#pragma unroll 2
            for (int k = 0; k < n4; ++k) {
                body(kk);
                kk = (kk + 1 == n4) ? 0 : kk + 1;
            }
        } else {
            int kk = cg % n_in;                               // per-lane rotation of the reduction index
            for (int k = 0; k < n_in; ++k) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Ain[(rg + i * RG) * n_in + kk];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) w[jj] = W[jc[jj] * n_in + kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i], w[jj], acc[i][jj]);
                if (++kk == n_in) kk = 0;
            }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = cg + jj * ncg;
            if (j < n_out) {
#pragma unroll
                for (int i = 0; i < 4; ++i) Aout[(rg + i * RG) * n_out + j] = act_fwd(acc[i][jj], act);
            }
        }
    }
}

// gW[n_out x n_in] += dz^T[n_out x T] . Ain[T x n_in];   gb[n_out] += column sums of dz
__device__ __forceinline__ void mlp_weight_grad(const float* Ain, const float* dz, float* gW, float* gb, int n_in,
                                                int n_out, int T) {
    const int njg = (n_out + 3) >> 2;
    if (n_out * n_in * 2 <= MLP_THREADS * 4 && n_out * n_in <= MLP_THREADS) {
        // thin: one output (j,k) per lane group, the T rows split over the group's lanes
        const int outputs = n_out * n_in, S = thin_group(outputs);
        for (int base = 0; base < outputs; base += MLP_THREADS / S) {
            const int o = base + threadIdx.x / S, s = threadIdx.x % S;
            const bool live = o < outputs;
            const int j = live ? o / n_in : 0, k = live ? o % n_in : 0;
            float acc = 0.0f;
            for (int r = s; r < T; r += S) acc = fmaf(dz[r * n_out + j], Ain[r * n_in + k], acc);
            acc = group_sum(acc, S);
            if (live && s == 0) gW[j * n_in + k] += acc;
        }
    } else if (((n_in & 3) == 0) && ((((size_t)gW) & 15) == 0) && ((((size_t)Ain) & 15) == 0)) {
        // 4 (rows of gW: j = jg + jj*njg) x 4 (contiguous k) micro-tiles, float4 activations and float4 RMW of gW
        const int nk4 = n_in >> 2;
        const float4* A4 = reinterpret_cast<const float4*>(Ain);
        for (int tile = threadIdx.x; tile < njg * nk4; tile += MLP_THREADS) {
            const int k4 = tile % nk4, jg = tile / nk4;
            int jc[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int j = jg + t * njg; jc[t] = j < n_out ? j : n_out - 1; }
            float4 acc[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int r = 0; r < T; ++r) {
                const float4 a = A4[r * nk4 + k4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float d = dz[r * n_out + jc[jj]];
                    acc[jj].x = fmaf(d, a.x, acc[jj].x); acc[jj].y = fmaf(d, a.y, acc[jj].y);
                    acc[jj].z = fmaf(d, a.z, acc[jj].z); acc[jj].w = fmaf(d, a.w, acc[jj].w);
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = jg + jj * njg;
                if (j >= n_out) continue;
                float4* dst = reinterpret_cast<float4*>(gW + j * n_in) + k4;
                float4 v = *dst;
                v.x += acc[jj].x; v.y += acc[jj].y; v.z += acc[jj].z; v.w += acc[jj].w;
                *dst = v;
            }
        }
    } else {
        const int nkg = (n_in + 3) >> 2;
        for (int tile = threadIdx.x; tile < njg * nkg; tile += MLP_THREADS) {
            const int kg = tile % nkg, jg = tile / nkg;
            int jc[4], kc[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = jg + t * njg, k = kg + t * nkg;
                jc[t] = j < n_out ? j : n_out - 1;
                kc[t] = k < n_in ? k : n_in - 1;
            }
            float acc[4][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[jj][t] = 0.0f;
            for (int r = 0; r < T; ++r) {
                float d[4], a[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) { d[t] = dz[r * n_out + jc[t]]; a[t] = Ain[r * n_in + kc[t]]; }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                    for (int t = 0; t < 4; ++t) acc[jj][t] = fmaf(d[jj], a[t], acc[jj][t]);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = jg + jj * njg;
                if (j >= n_out) continue;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int k = kg + t * nkg;
                    if (k < n_in) gW[j * n_in + k] += acc[jj][t];
                }
            }
        }
    }
    for (int j = threadIdx.x; j < n_out; j += MLP_THREADS) {
        float s = 0.0f;
        for (int r = 0; r < T; ++r) s += dz[r * n_out + j];
        gb[j] += s;
    }
}

// dz_prev[T x n_in] = (dz[T x n_out] . W[n_out x n_in]) * act'(A[T x n_in])
__device__ __forceinline__ void mlp_input_grad(const float* dz, const float* W, const float* A, float* dz_prev,
                                               int n_in, int n_out, int act_prev, int T) {
    const int nkg = (n_in + 3) >> 2, MLP_RG = T >> 2;
    for (int tile = threadIdx.x; tile < MLP_RG * nkg; tile += MLP_THREADS) {
        const int kg = tile % nkg, rg = tile / nkg;
        int kc[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { const int k = kg + t * nkg; kc[t] = k < n_in ? k : n_in - 1; }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[i][t] = 0.0f;
        for (int j = 0; j < n_out; ++j) {
            float d[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = dz[(rg + i * MLP_RG) * n_out + j];
#pragma unroll
            for (int t = 0; t < 4; ++t) w[t] = W[j * n_in + kc[t]];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(d[i], w[t], acc[i][t]);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = kg + t * nkg;
            if (k >= n_in) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int o = (rg + i * MLP_RG) * n_in + k;
                dz_prev[o] = acc[i][t] * act_bwd(A[o], act_prev);
            }
        }
    }
}

// forward pass of one tile; returns nothing, leaves A[0..L] in the tile area
__device__ __forceinline__ void mlp_forward_tile(const MlpDev& m, const float* q, float* tile, int r0, int cnt) {
    mlp_load_x(m, tile + m.aoff[0], r0, cnt);
    __syncthreads();
    for (int l = 0; l < m.L; ++l) {
        mlp_linear_fwd(tile + m.aoff[l], q + m.woff[l], q + m.boff[l], tile + m.aoff[l + 1], m.n[l], m.n[l + 1],
                       m.act[l], m.T);
        __syncthreads();
    }
}

// Loss stage of a forwarded tile: returns the thread's partial of the loss sum (ll = c_ll * sum, or c_ll * sum / rows
// for the mean-reduced nll_loss) and, when dz != nullptr, writes dz_L = d ll / d out.  `rows` = data rows of the closure
// being evaluated (the split), only used by the mean reduction.  With log_softmax_out the tile's outputs are replaced
// by their log-softmax (what predict_model returns for such a model).
__device__ __forceinline__ float mlp_loss_tile(const MlpDev& m, float* out, float* dz, int r0, int cnt, int rows,
                                               bool log_softmax_out = false, const float* ytile = nullptr) {
    const int nL = m.n[m.L];
    float sum = 0.0f;
    if (m.loss == HMCX_LOSS_REGRESSION || m.loss == HMCX_LOSS_BINARY) {
        const float c2 = mul(m.c_ll, 2.0f);                               // regression, autograd: ds * (2*diff)
        for (int i = threadIdx.x; i < m.T * nL; i += MLP_THREADS) {
            float d = 0.0f, gz = 0.0f;
            if (i < cnt * nL) {
                const float yv = ytile ? ytile[i] : __ldg(m.y + (size_t)r0 * nL + i), z = out[i];
                if (m.loss == HMCX_LOSS_REGRESSION) {
                    d = sub(z, yv);
                    sum = add(sum, mul(d, d));
                    gz = mul(c2, d);
                } else {                                                  // BCE with logits, sum reduction
                    sum += (1.0f - yv) * z + fmaxf(-z, 0.0f) + log1pf(expf(-fabsf(z)));
                    gz = m.c_ll * (1.0f / (1.0f + expf(-z)) - yv);
                }
            }
            if (dz) dz[i] = gz;
        }
        return sum;
    }
    // multi-class: one thread per data row
    const float cg = (m.loss == HMCX_LOSS_MULTICLASS_LOGSOFTMAX) ? m.c_ll / (float)rows : m.c_ll;
    for (int r = threadIdx.x; r < m.T; r += MLP_THREADS) {
        float* z = out + r * nL;
        if (r < cnt) {
            float mx = z[0];
            for (int k = 1; k < nL; ++k) mx = fmaxf(mx, z[k]);
            float se = 0.0f;
            for (int k = 0; k < nL; ++k) se += expf(z[k] - mx);
            const float lse = mx + logf(se);
            int label = (int)(ytile ? ytile[r] : __ldg(m.y + r0 + r));
            label = label < 0 ? 0 : (label >= nL ? nL - 1 : label);
            sum += lse - z[label];
            for (int k = 0; k < nL; ++k) {
                if (dz) dz[r * nL + k] = cg * (expf(z[k] - lse) - (k == label ? 1.0f : 0.0f));
                if (log_softmax_out) z[k] = z[k] - lse;
            }
        } else if (dz) {
            for (int k = 0; k < nL; ++k) dz[r * nL + k] = 0.0f;
        }
    }
    return sum;
}

// ll of one closure from its reduced loss sum
__device__ __forceinline__ float mlp_ll_from_sum(const MlpDev& m, float sum, int rows) {
    if (m.loss == HMCX_LOSS_MULTICLASS_LOGSOFTMAX) return mul(m.c_ll, __fdiv_rn(sum, (float)rows));
    return mul(m.c_ll, sum);
}

// ---- thread-block clusters: a chain may be owned by CS cooperating CTAs (CS SMs) --------------------------------
// Every CTA of the cluster keeps the full q, p, g in its own shared memory and does the (cheap) element-wise work
// redundantly and identically; the expensive part -- the data rows of a gradient / log-prob evaluation -- is divided
// by tile index (tile i -> rank i % CS), and the partial results are combined through DISTRIBUTED SHARED MEMORY in
// the fixed order rank 0, 1, ... so that all CTAs hold bit-identical sums.  CS is a function of the data layout only
// (never of the number of chains), so results do not depend on how chains are sharded over GPUs.
struct ClusterCtx { int rank, size; };

template <int CS>
__device__ __forceinline__ void cluster_sum_vector(float* v, int n) {
    if (CS == 1) return;
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                            // every partial is complete
    constexpr int MAXPT = 36;                                  // n <= 512*36 (the three state vectors must fit smem anyway)
    float acc[MAXPT];
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) {
        const int i = threadIdx.x + t * MLP_THREADS;
        float sum = 0.0f;
        if (i < n) {
#pragma unroll
            for (int r = 0; r < CS; ++r) {
                const float x = cluster.map_shared_rank(v, r)[i];
                sum = (r == 0) ? x : add(sum, x);
            }
        }
        acc[t] = sum;
    }
    cluster.sync();                                            // everybody has read everybody's partial
#pragma unroll
    for (int t = 0; t < MAXPT; ++t) {
        const int i = threadIdx.x + t * MLP_THREADS;
        if (i < n) v[i] = acc[t];
    }
    __syncthreads();
}

// scalar version: `slot` is a shared-memory word of this CTA; returns sum over ranks in rank order
template <int CS>
__device__ __forceinline__ float cluster_sum_scalar(float x, float* slot) {
    if (CS == 1) return x;
    cg::cluster_group cluster = cg::this_cluster();
    if (threadIdx.x == 0) *slot = x;
    cluster.sync();
    float sum = 0.0f;
#pragma unroll
    for (int r = 0; r < CS; ++r) {
        const float y = *cluster.map_shared_rank(slot, r);
        sum = (r == 0) ? y : add(sum, y);
    }
    cluster.sync();                                            // slot may be rewritten after this
    return sum;
}


// =========================================================================================================
// First-layer GEMMs on the 5th-generation tensor cores (tcgen05 / TMEM), for one-hidden-layer stacks
//   n0 -> 128 -> nL   (n0 in {16,32,48,64}, nL <= 4; BASELINE config 4 is 64-128-1)
// where  H = X W1^T  (forward) and  dW1 = dH^T X  (backward) carry ~all of the flops.  Per 64-row tile of the split:
//   forward, TRANSPOSED:  H^T[128 units x 64 rows] = W1[128 x n0] . X_tile^T   -- A = W1 (shared memory, K-major,
//       packed from the flat q at the start of every evaluation), B = X tile (staged from global memory), fp32
//       accumulators in TENSOR MEMORY; 3xTF32 split operands (hi*hi + hi*lo + lo*hi) keep fp32-level accuracy;
//   epilogue 1: each thread owns ONE hidden unit (TMEM lane) and 16 rows (columns): bias + activation in registers,
//       the thin output layer as an in-warp transpose-reduction -> z2 in shared memory -> the usual loss stage;
//   epilogue 2: dH^T = (W2^T dz2) * act'(H) computed in the same registers; db1 and dW2 are THREAD-LOCAL sums in this
//       orientation; dH^T is written back to TENSOR MEMORY (tcgen05.st) as tf32 hi / lo;
//   backward:  dW1[128 x n0] += dH^T[128 x 64 rows] . X_tile  with the A operand read FROM TENSOR MEMORY and
//       B = X tile re-staged rows-contiguous; the accumulator stays in TMEM across all tiles of the split.
// The big activations never touch shared memory.  Everything around (prior, schedules, kicks, drifts, Hamiltonians,
// MH, clusters) is the code of the SIMT path.
// =========================================================================================================
constexpr int TC_TR = 64;                 // data rows per tile (MMA N forward, MMA K backward)
constexpr int TC_H = 128;                 // hidden units = MMA M = TMEM lanes
constexpr int TC_NLMAX = 4;               // outputs handled by the register head
// TMEM columns: H hh / dH hi | H hl / dH lo | dW1 hh | dW1 hl | W1 hi | W1 lo (the forward A operand, written once per evaluation)
constexpr int TC_COL_H = 0, TC_COL_LO = 64, TC_COL_W = 128, TC_COL_W1HI = 256, TC_COL_W1LO = 320, TC_COLS = 512;

static_assert(TC_TR / 4 == MLP_THREADS / 32 && TC_TR == 64, "staging / epilogue thread maps assume 16 warps and 64-row tiles");

struct TcCtx { uint32_t tmem, barH, barW, barF[2], barB, parH, parW, parF[2], parB; int pre_id; };   // pre_id: tile whose operands an earlier evaluation already prefetched (-1: none)

#ifdef HMCX_TC_PROF
__device__ long long g_tc_prof[512];
__device__ int g_tc_prof_n;
#define TC_MARK(id) do { if (threadIdx.x == 0 && blockIdx.x == 0) { int k_ = g_tc_prof_n; if (k_ < 510) { g_tc_prof[k_] = ((long long)(id) << 48) | (clock64() & 0xFFFFFFFFFFFFll); g_tc_prof_n = k_ + 1; } } } while (0)
#else
#define TC_MARK(id) do {} while (0)
#endif

__device__ __forceinline__ void tc_init(TcCtx& tc, uint64_t* bars, uint32_t* slot) {
    if (threadIdx.x == 0) {
        for (int b = 0; b < 5; ++b) mbar_init(smem_u32(&bars[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(TC_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc.tmem = *slot;
    tc.barH = smem_u32(&bars[0]);
    tc.barW = smem_u32(&bars[1]);
    tc.barF[0] = smem_u32(&bars[2]); tc.barF[1] = smem_u32(&bars[3]); tc.barB = smem_u32(&bars[4]);
    tc.parH = tc.parW = tc.parF[0] = tc.parF[1] = tc.parB = 0;
    tc.pre_id = -1;
}
__device__ __forceinline__ void tc_fini(const TcCtx& tc) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tc.tmem), "r"(TC_COLS));
}

// round-to-nearest (ties away) tf32 in an fp32 container == cvt.rna.tf32.f32 for finite inputs, as two full-rate
// integer instructions instead of a conversion-pipe instruction (the epilogues split ~50 values per thread per tile)
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void tc_split4(const float4 v, float4& h, float4& l) {
    h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
    l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
}

// q's W1 (flat, row-major 128 x n0) -> the forward A operand IN TENSOR MEMORY (lane = hidden unit, one 32-bit column per
// input k), tf32 hi and lo: thread (unit u, column group cq) owns 16 consecutive k of row u.  The four float4 chunks are
// read in a per-lane rotated order (2-way instead of 8-way bank conflicts on the stride-n0 rows) and rotated back in
// registers.  Keeping W1 out of shared memory frees 64 KB for the X operand pipeline.
__device__ __forceinline__ void tc_pack_w1(const MlpDev& m, const float* q, const TcCtx& tc) {
    const int n0 = m.n[0], warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lq = warp & 3, cq = warp >> 2, u = 32 * lq + lane;
    if (16 * cq < n0) {
        const float* W = q + m.woff[0] + u * n0 + 16 * cq;
        float4 t[4], a[4], o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = *reinterpret_cast<const float4*>(W + 4 * ((j + lane) & 3));
        // t[j] holds chunk (j + lane) & 3, i.e. chunk c sits in t[(c - lane) & 3]: rotate right by lane & 3
        const bool r1 = lane & 1, r2 = lane & 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[c].x = r1 ? t[(c + 3) & 3].x : t[c].x; a[c].y = r1 ? t[(c + 3) & 3].y : t[c].y;
            a[c].z = r1 ? t[(c + 3) & 3].z : t[c].z; a[c].w = r1 ? t[(c + 3) & 3].w : t[c].w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o[c].x = r2 ? a[(c + 2) & 3].x : a[c].x; o[c].y = r2 ? a[(c + 2) & 3].y : a[c].y;
            o[c].z = r2 ? a[(c + 2) & 3].z : a[c].z; o[c].w = r2 ? a[(c + 2) & 3].w : a[c].w;
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float4 h, l;
            tc_split4(o[c], h, l);
            hi[4 * c] = __float_as_uint(h.x); hi[4 * c + 1] = __float_as_uint(h.y);
            hi[4 * c + 2] = __float_as_uint(h.z); hi[4 * c + 3] = __float_as_uint(h.w);
            lo[4 * c] = __float_as_uint(l.x); lo[4 * c + 1] = __float_as_uint(l.y);
            lo[4 * c + 2] = __float_as_uint(l.z); lo[4 * c + 3] = __float_as_uint(l.w);
        }
        const uint32_t tl = tc.tmem + ((uint32_t)(32 * lq) << 16) + 16 * cq;
        tmem_st16(tl + TC_COL_W1HI, hi);
        tmem_st16(tl + TC_COL_W1LO, lo);
    }
    tmem_st_wait();
    tc_fence_before();                                        // the caller's __syncthreads orders it before the UMMAs
}

// X never changes during a run, so its tf32 hi / lo split and both operand layouts are built ONCE (hmcx_mlp_pack_x ->
// mlp_pack_x_kernel) and a tile's operands arrive ready-made by one bulk TMA copy each -- no staging pass, no landing
// buffer (the first form re-split and re-laid-out every tile twice per evaluation: ~2.4k of its ~9.9k cycles):
//   forward B operand  [64 rows hi | 64 rows lo (N = 128)] x [n0 (K)], K-major core matrices (8 rows x 16 B); stacking
//       hi|lo along N lets ONE UMMA produce W1_hi X_hi^T and W1_hi X_lo^T side by side;
//   backward B operand [n0 hi | n0 lo (N = 2 n0)] x [64 rows (K)], K-major (4 consecutive ROWS of one input column per 16 B).
// Rows past the end of a ragged last tile are zero in the packed copy.
__device__ __forceinline__ int tc_pack_off_fwd(int r, int c) { return (c * (2 * TC_TR >> 3) + (r >> 3)) * 32 + (r & 7) * 4; }
__device__ __forceinline__ int tc_pack_off_bwd(int a, int n, int n0) { return (a * (2 * n0 >> 3) + (n >> 3)) * 32 + (n & 7) * 4; }

__global__ void __launch_bounds__(256) mlp_pack_x_kernel(const MlpDev m, float* __restrict__ out) {
    const int n0 = m.n[0], nch = n0 >> 2, tile_id = blockIdx.x;
    int r0, r_end;
    if (tile_id >= m.flat_base && m.flat_base >= m.tb[m.M]) {        // the all-rows tiling (only packed when it differs)
        r0 = TC_TR * (tile_id - m.flat_base); r_end = m.N;
    } else {
        int sp = 0;
        while (sp + 1 < m.M && tile_id >= m.tb[sp + 1]) ++sp;
        r0 = m.sb[sp] + TC_TR * (tile_id - m.tb[sp]); r_end = m.sb[sp + 1];
    }
    const int cnt = min(TC_TR, r_end - r0);
    float* fwd = out + (size_t)tile_id * (2 * 2 * TC_TR * n0);
    float* bwd = fwd + 2 * TC_TR * n0;
    for (int i = threadIdx.x; i < TC_TR * nch; i += blockDim.x) {
        const int r = i / nch, c = i - r * nch;
        const float4 v = r < cnt ? *reinterpret_cast<const float4*>(m.x + (size_t)(r0 + r) * n0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 h, l;
        tc_split4(v, h, l);
        *reinterpret_cast<float4*>(fwd + tc_pack_off_fwd(r, c)) = h;
        *reinterpret_cast<float4*>(fwd + tc_pack_off_fwd(r, c) + (TC_TR >> 3) * 32) = l;
    }
    for (int i = threadIdx.x; i < (TC_TR / 4) * n0; i += blockDim.x) {
        const int a = i / n0, n = i - a * n0;
        float e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] = (4 * a + k) < cnt ? m.x[(size_t)(r0 + 4 * a + k) * n0 + n] : 0.0f;
        float4 h, l;
        tc_split4(make_float4(e[0], e[1], e[2], e[3]), h, l);
        *reinterpret_cast<float4*>(bwd + tc_pack_off_bwd(a, n, n0)) = h;
        *reinterpret_cast<float4*>(bwd + tc_pack_off_bwd(a, n, n0) + (n0 >> 3) * 32) = l;
    }
}

// one bulk TMA copy per operand (thread 0; completion on the buffer's mbarrier)
__device__ __forceinline__ void tc_prefetch_fwd(const MlpDev& m, float* tile, const TcCtx& tc, int tile_id, int buf) {
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(2 * TC_TR * m.n[0]) * 4u;
        mbar_expect_tx(tc.barF[buf], bytes);
        bulk_g2s(smem_u32(tile + (buf ? m.tc_f1 : m.tc_f0)), m.xp + (size_t)tile_id * (4 * TC_TR * m.n[0]), bytes, tc.barF[buf]);
    }
}
__device__ __forceinline__ void tc_prefetch_bwd(const MlpDev& m, float* tile, const TcCtx& tc, int tile_id) {
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(2 * TC_TR * m.n[0]) * 4u;
        mbar_expect_tx(tc.barB, bytes);
        bulk_g2s(smem_u32(tile + m.tc_b), m.xp + (size_t)tile_id * (4 * TC_TR * m.n[0]) + 2 * TC_TR * m.n[0], bytes, tc.barB);
    }
}
// the tile's targets: 4-byte cp.async by warp 1 (waited for by the same warp at the top of tc_forward_tile)
__device__ __forceinline__ void tc_prefetch_y(const MlpDev& m, float* tile, int r0, int cnt) {
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        const int ycols = (m.loss == HMCX_LOSS_REGRESSION || m.loss == HMCX_LOSS_BINARY) ? m.n[2] : 1;
        const uint32_t yraw = smem_u32(tile + m.tc_yraw);
        for (int t = threadIdx.x - 32; t < cnt * ycols; t += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(yraw + 4u * (uint32_t)t), "l"(m.y + (size_t)r0 * ycols + t) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
}

// H^T = W1 . X^T  (one elected thread; completion -> barH).  A = W1 from TENSOR MEMORY (8 columns per k-step).  Two UMMAs per
// k-step: W1_hi . [X_hi | X_lo]^T (N = 128, the two products land in columns [0,64) and [64,128)) and W1_lo . X_hi^T
// (N = 64, accumulated onto [0,64)); the epilogue adds the two column blocks.  The B descriptors differ only in the
// start-address field (bits 0-13, 16-byte units), so a k-step is an integer add.
__device__ __forceinline__ void tc_issue_fwd(const MlpDev& m, const TcCtx& tc, const float* xop, uint32_t leader) {
    const uint32_t idesc2 = make_idesc_tf32(TC_H, 2 * TC_TR), idesc1 = make_idesc_tf32(TC_H, TC_TR);
    constexpr uint32_t B_LBO = (2 * TC_TR / 8) * 128;
    uint64_t b = make_kmajor_desc(smem_u32(xop), B_LBO, 128);
    const int ksteps = m.n[0] >> 3;
    const uint32_t d = tc.tmem + TC_COL_H;
    uint32_t a_hi = tc.tmem + TC_COL_W1HI, a_lo = tc.tmem + TC_COL_W1LO;
    tc_fence_after();
    for (int k = 0; k < ksteps; ++k) {
        umma_tf32_ta_p(d, a_hi, b, idesc2, k != 0, leader);
        umma_tf32_ta_p(d, a_lo, b, idesc1, true, leader);
        a_hi += 8; a_lo += 8; b += (2 * B_LBO) >> 4;
    }
    umma_commit_p(tc.barH, leader);
}

// dW1 (+)= dH^T . X   (A from tensor memory; completion -> barW): dH_hi . [X_hi | X_lo] (N = 2 n0) and dH_lo . X_hi (N = n0)
__device__ __forceinline__ void tc_issue_bwd(const MlpDev& m, const TcCtx& tc, const float* xop, bool accumulate, uint32_t leader) {
    const int n0 = m.n[0];
    const uint32_t idesc2 = make_idesc_tf32(TC_H, 2 * n0), idesc1 = make_idesc_tf32(TC_H, n0);
    const uint32_t B_LBO = (uint32_t)(2 * n0 / 8) * 128;
    uint64_t b = make_kmajor_desc(smem_u32(xop), B_LBO, 128);
    const uint32_t d = tc.tmem + TC_COL_W;
    uint32_t a_hi = tc.tmem + TC_COL_H, a_lo = tc.tmem + TC_COL_LO;
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < TC_TR / 8; ++k) {
        umma_tf32_ta_p(d, a_hi, b, idesc2, accumulate || k != 0, leader);
        umma_tf32_ta_p(d, a_lo, b, idesc1, true, leader);
        a_hi += 8; a_lo += 8; b += (2 * B_LBO) >> 4;
    }
    umma_commit_p(tc.barW, leader);
}

// v[i] = this lane's value for row i; returns the sum over the warp's 32 lanes for row (lane >> 1)
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16]) {
    const int lane = threadIdx.x & 31;
#define HMCX_TS_STEP(HALF, OFF)                                                          \
    {                                                                                    \
        const bool upper = (lane & OFF) != 0;                                            \
        _Pragma("unroll") for (int i = 0; i < HALF; ++i) {                               \
            const float send = upper ? v[i] : v[i + HALF];                               \
            const float keep = upper ? v[i + HALF] : v[i];                               \
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);                       \
        }                                                                                \
    }
    HMCX_TS_STEP(8, 16) HMCX_TS_STEP(4, 8) HMCX_TS_STEP(2, 4) HMCX_TS_STEP(1, 2)
#undef HMCX_TS_STEP
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// bias + activation / activation derivative over a thread's 16 values, the activation kind resolved once
template <int A>
__device__ __forceinline__ void tc_act16_t(const uint32_t (&v)[16], float b, float (&act)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) act[i] = act_fwd(__uint_as_float(v[i]) + b, A);
}
__device__ __forceinline__ void tc_act16(const uint32_t (&v)[16], float b, float (&act)[16], int a) {
    if (a == HMCX_ACT_RELU) tc_act16_t<HMCX_ACT_RELU>(v, b, act);
    else if (a == HMCX_ACT_TANH) tc_act16_t<HMCX_ACT_TANH>(v, b, act);
    else if (a == HMCX_ACT_SIGMOID) tc_act16_t<HMCX_ACT_SIGMOID>(v, b, act);
    else tc_act16_t<HMCX_ACT_NONE>(v, b, act);
}
template <int A>
__device__ __forceinline__ void tc_dact16_t(const float (&act)[16], float (&d)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = act_bwd(act[i], A);
}
__device__ __forceinline__ void tc_dact16(const float (&act)[16], float (&d)[16], int a) {
    if (a == HMCX_ACT_RELU) tc_dact16_t<HMCX_ACT_RELU>(act, d);
    else if (a == HMCX_ACT_TANH) tc_dact16_t<HMCX_ACT_TANH>(act, d);
    else if (a == HMCX_ACT_SIGMOID) tc_dact16_t<HMCX_ACT_SIGMOID>(act, d);
    else tc_dact16_t<HMCX_ACT_NONE>(act, d);
}

// per-thread constants / accumulators of the register epilogues: thread <-> hidden unit u, 16 rows of the tile
struct TcEpi {
    int u, cq, lq;
    float b1u, w2[TC_NLMAX];
    float db1, dw2[TC_NLMAX], db2;
};
__device__ __forceinline__ void tc_epi_begin(const MlpDev& m, const float* q, TcEpi& e) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    e.lq = warp & 3; e.cq = warp >> 2; e.u = 32 * e.lq + lane;
    e.b1u = q[m.boff[0] + e.u];
#pragma unroll
    for (int j = 0; j < TC_NLMAX; ++j) {
        e.w2[j] = j < m.n[2] ? q[m.woff[1] + j * TC_H + e.u] : 0.0f;
        e.dw2[j] = 0.0f;
    }
    e.db1 = 0.0f; e.db2 = 0.0f;
}

// forward of one (prefetched) tile up to the network outputs: out[r * nL + j] (the loss stage's layout); the hidden
// activations of this thread's (unit, 16 rows) block stay in `act`.  Ends with every thread past a __syncthreads.
__device__ __forceinline__ void tc_forward_tile(const MlpDev& m, const float* q, float* tile, TcCtx& tc, const TcEpi& e,
                                                float (&act)[16], int buf) {
    const int nL = m.n[2], lane = threadIdx.x & 31;
    if (threadIdx.x >= 32 && threadIdx.x < 64) asm volatile("cp.async.wait_group 0;" ::: "memory");    // the targets
    if (threadIdx.x < 32) {                                       // warp 0 issues, warp-uniformly (hmcx_umma.cuh): ~65 instead of
        mbar_wait(tc.barF[buf], tc.parF[buf]);                    // ~115 cycles per MMA.  This tile's forward operand has landed
        TC_MARK(3);
        tc_issue_fwd(m, tc, tile + (buf ? m.tc_f1 : m.tc_f0), elect_one());
    }
    tc.parF[buf] ^= 1;
    TC_MARK(4);
    mbar_wait(tc.barH, tc.parH);
    tc.parH ^= 1;
    tc_fence_after();
    TC_MARK(5);
    uint32_t v[16], w[16];
    tmem_ld16(tc.tmem + ((uint32_t)(32 * e.lq) << 16) + TC_COL_H + 16 * e.cq, v);
    tmem_ld16(tc.tmem + ((uint32_t)(32 * e.lq) << 16) + TC_COL_LO + 16 * e.cq, w);       // the W1_hi X_lo block
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
    tc_act16(v, e.b1u, act, m.act[0]);
    float* part = tile + m.tc_part;
#pragma unroll
    for (int j = 0; j < TC_NLMAX; ++j) {
        if (j < nL) {
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = e.w2[j] * act[i];
            const float sum = warp_transpose_sum16(t);
            if ((lane & 1) == 0) part[(e.lq * TC_TR + 16 * e.cq + (lane >> 1)) * TC_NLMAX + j] = sum;
        }
    }
    __syncthreads();
    TC_MARK(6);
    float* out = tile + m.aoff[2];
    for (int i = threadIdx.x; i < TC_TR * nL; i += MLP_THREADS) {
        const int r = i / nL, j = i - r * nL;
        const float s = ((part[(0 * TC_TR + r) * TC_NLMAX + j] + part[(1 * TC_TR + r) * TC_NLMAX + j]) +
                         part[(2 * TC_TR + r) * TC_NLMAX + j]) + part[(3 * TC_TR + r) * TC_NLMAX + j];
        out[i] = s + q[m.boff[1] + j];
    }
    __syncthreads();
    TC_MARK(7);
}

// g += d ll_split / dq over this rank's 64-row tiles of [r_begin, r_end)
__device__ __forceinline__ void tc_backprop_rows(const MlpDev& m, const float* q, float* g, float* tile, TcCtx& tc,
                                                 int r_begin, int r_end, int tile0, ClusterCtx cc, int next_s = -2) {
    const int nL = m.n[2], n0 = m.n[0];
    TcEpi e;
    tc_epi_begin(m, q, e);
    float* out = tile + m.aoff[2];
    float* dz = tile + m.dzoff[0];
    const float* ytile = tile + m.tc_yraw;
    const int stride = TC_TR * cc.size;
    int done = 0;
    int r0 = r_begin + TC_TR * cc.rank;                           // this rank's tiles: rank, rank + size, ...
    int id = tile0 + cc.rank;                                     // ... and their packed operands
    if (r0 < r_end && tc.pre_id != id) {                          // (else: the previous evaluation already fetched them)
        tc_prefetch_fwd(m, tile, tc, id, 0);
        tc_prefetch_bwd(m, tile, tc, id);
        tc_prefetch_y(m, tile, r0, min(TC_TR, r_end - r0));
    }
    tc.pre_id = -1;
    for (; r0 < r_end; r0 += stride, id += cc.size) {
        const int cnt = min(TC_TR, r_end - r0), buf = done & 1;
        const bool has_next = r0 + stride < r_end;
        if (has_next) tc_prefetch_fwd(m, tile, tc, id + cc.size, buf ^ 1);   // last read by the forward MMA of tile t-1: complete
        float act[16];
        tc_forward_tile(m, q, tile, tc, e, act, buf);
        mlp_loss_tile(m, out, dz, r0, cnt, r_end - r_begin, false, ytile);
        __syncthreads();
        TC_MARK(8);
        if (has_next) tc_prefetch_y(m, tile, r0 + stride, min(TC_TR, r_end - r0 - stride));   // the target buffer is free
        if (threadIdx.x < TC_TR * nL) e.db2 += dz[threadIdx.x];   // element (r, j) of every tile; reduced over r at the end
        float dact[16];
        tc_dact16(act, dact, m.act[0]);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = 16 * e.cq + i;
            float da = 0.0f;
#pragma unroll
            for (int j = 0; j < TC_NLMAX; ++j) {
                if (j < nL) {
                    const float d = dz[r * nL + j];
                    da = fmaf(d, e.w2[j], da);
                    e.dw2[j] = fmaf(d, act[i], e.dw2[j]);
                }
            }
            const float dh = da * dact[i];
            e.db1 += dh;
            const float h = tf32_rn(dh);
            hi[i] = __float_as_uint(h);
            lo[i] = __float_as_uint(tf32_rn(dh - h));
        }
        const uint32_t tl = tc.tmem + ((uint32_t)(32 * e.lq) << 16) + 16 * e.cq;
        tmem_st16(tl + TC_COL_H, hi);
        tmem_st16(tl + TC_COL_LO, lo);
        tmem_st_wait();
        tc_fence_before();
        TC_MARK(9);
        __syncthreads();
        TC_MARK(10);
        if (threadIdx.x < 32) {
            mbar_wait(tc.barB, tc.parB);                          // the backward operand has landed (long ago)
            tc_issue_bwd(m, tc, tile + m.tc_b, done != 0, elect_one());
        }
        tc.parB ^= 1;
        ++done;
        TC_MARK(11);
        mbar_wait(tc.barW, tc.parW);                              // the backward operand buffer and the dH columns are free again
        tc.parW ^= 1;
        if (has_next) tc_prefetch_bwd(m, tile, tc, id + cc.size);
        TC_MARK(12);
    }
    if (done == 0) return;                                        // (uniform over the CTA)
    // The first tile of the NEXT evaluation (the schedule knows its split): its operands do not depend on q, so they are
    // requested now and land while this evaluation finishes (dW1 read-out, reductions, cluster sum, kick, drift) -- an
    // evaluation used to start with ~2.3k cycles of exposed TMA latency.
    int nid = -1;
    if (next_s > -2) {
        const int nb = next_s < 0 ? 0 : m.sb[next_s], ne = next_s < 0 ? m.N : m.sb[next_s + 1];
        const int nr0 = nb + TC_TR * cc.rank;
        if (nr0 < ne) {
            nid = (next_s < 0 ? m.flat_base : m.tb[next_s]) + cc.rank;
            tc_prefetch_fwd(m, tile, tc, nid, 0);                 // every forward MMA of this evaluation has completed
            tc_prefetch_y(m, tile, nr0, min(TC_TR, ne - nr0));    // the last loss stage is over
        }
    }
    tc_fence_after();
    // ---- dW1: TMEM -> padded staging -> g, conflict-free both ways
    float* stg = tile + m.tc_f1;                 // forward buffer 1 | backward buffer (adjacent, both idle now)
    const int pitch = n0 + 4;
    if (16 * e.cq < n0) {
        uint32_t v[16], w[16];
        tmem_ld16(tc.tmem + ((uint32_t)(32 * e.lq) << 16) + TC_COL_W + 16 * e.cq, v);
        tmem_ld16(tc.tmem + ((uint32_t)(32 * e.lq) << 16) + TC_COL_W + n0 + 16 * e.cq, w);   // the dH_hi X_lo block
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
#pragma unroll
        for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(stg + e.u * pitch + 16 * e.cq + i) =
                make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    tc_fence_before();
    __syncthreads();
    float* gW = g + m.woff[0];
    for (int i4 = threadIdx.x; i4 < TC_H * n0 / 4; i4 += MLP_THREADS) {
        const int u = (4 * i4) / n0, k = 4 * i4 - u * n0;
        const float4 d = *reinterpret_cast<const float4*>(stg + u * pitch + k);
        float4 a = *reinterpret_cast<float4*>(gW + 4 * i4);
        a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
        *reinterpret_cast<float4*>(gW + 4 * i4) = a;
    }
    // ---- db1, dW2: four per-unit partials (one per 16-row column group) summed in a fixed order; db2
    float* red = tile + m.tc_part;
    red[(e.cq * TC_H + e.u) * (1 + TC_NLMAX)] = e.db1;
#pragma unroll
    for (int j = 0; j < TC_NLMAX; ++j) red[(e.cq * TC_H + e.u) * (1 + TC_NLMAX) + 1 + j] = e.dw2[j];
    fence_async_smem();                                           // the staging reads above precede the TMA write below
    __syncthreads();
    if (nid >= 0) { tc_prefetch_bwd(m, tile, tc, nid); tc.pre_id = nid; }
    if (threadIdx.x < TC_H) {
        const int u = threadIdx.x;
        auto tot = [&](int f) {
            return ((red[(0 * TC_H + u) * (1 + TC_NLMAX) + f] + red[(1 * TC_H + u) * (1 + TC_NLMAX) + f]) +
                    red[(2 * TC_H + u) * (1 + TC_NLMAX) + f]) + red[(3 * TC_H + u) * (1 + TC_NLMAX) + f];
        };
        g[m.boff[0] + u] += tot(0);
        for (int j = 0; j < nL; ++j) g[m.woff[1] + j * TC_H + u] += tot(1 + j);
    }
    __syncthreads();
    if (threadIdx.x < TC_TR * nL) red[threadIdx.x] = e.db2;       // db2[j] = sum over rows r of the per-(r, j) sums
    __syncthreads();
    if (threadIdx.x < 32) {
        for (int j = 0; j < nL; ++j) {
            const float sj = warp_sum(red[threadIdx.x * nL + j] + red[(threadIdx.x + 32) * nL + j]);
            if (threadIdx.x == 0) g[m.boff[1] + j] += sj;
        }
    }
    __syncthreads();
    TC_MARK(13);
}

// g += d ll_split / dq over this rank's tiles of the rows [r_begin, r_end)  (g must already hold the prior part / zeros)
__device__ __forceinline__ void mlp_backprop_rows(const MlpDev& m, const float* q, float* g, float* tile, int r_begin,
                                                  int r_end, ClusterCtx cc) {
    int ti = 0;
    for (int r0 = r_begin; r0 < r_end; r0 += m.T, ++ti) {
        if (ti % cc.size != cc.rank) continue;
        const int cnt = min(m.T, r_end - r0);
        mlp_forward_tile(m, q, tile, r0, cnt);
        float* dz = tile + m.dzoff[m.L & 1];
        mlp_loss_tile(m, tile + m.aoff[m.L], dz, r0, cnt, r_end - r_begin);
        __syncthreads();
        for (int l = m.L - 1; l >= 0; --l) {
            float* dz_prev = tile + m.dzoff[l & 1];
            mlp_weight_grad(tile + m.aoff[l], dz, g + m.woff[l], g + m.boff[l], m.n[l], m.n[l + 1], m.T);
            if (l > 0)
                mlp_input_grad(dz, q + m.woff[l], tile + m.aoff[l], dz_prev, m.n[l], m.n[l + 1], m.act[l - 1], m.T);
            __syncthreads();
            dz = dz_prev;
        }
    }
}

// prior part of the gradient: d(prior/prior_scale)/dw = -(coef_i * (2w))   (pow/div backward order)
__device__ __forceinline__ void mlp_prior_grad(const MlpDev& m, const float* q, float* g) {
    for (int l = 0; l < m.L; ++l) {
        const int nw = m.n[l] * m.n[l + 1];
        const float cw = m.gcoef[2 * l], cb = m.gcoef[2 * l + 1];
        if (((m.woff[l] | nw) & 3) == 0) {                         // 16-byte aligned tensor: four weights per access
            const float4* q4 = reinterpret_cast<const float4*>(q + m.woff[l]);
            float4* g4 = reinterpret_cast<float4*>(g + m.woff[l]);
            for (int i = threadIdx.x; i < (nw >> 2); i += MLP_THREADS) {
                const float4 w = q4[i];
                g4[i] = make_float4(-mul(cw, mul(2.0f, w.x)), -mul(cw, mul(2.0f, w.y)), -mul(cw, mul(2.0f, w.z)),
                                    -mul(cw, mul(2.0f, w.w)));
            }
        } else {
            for (int i = threadIdx.x; i < nw; i += MLP_THREADS) g[m.woff[l] + i] = -mul(cw, mul(2.0f, q[m.woff[l] + i]));
        }
        for (int i = threadIdx.x; i < m.n[l + 1]; i += MLP_THREADS)
            g[m.boff[l] + i] = -mul(cb, mul(2.0f, q[m.boff[l] + i]));
    }
}

// g = d log p_split / dq for split s (s < 0: all rows as one potential)
template <int CS>
__device__ __forceinline__ void mlp_grad_split(const MlpDev& m, const float* q, float* g, float* tile, int s,
                                               ClusterCtx cc, TcCtx& tc, bool leave_partials = false, int next_s = -2) {
    TC_MARK(1);
    if (cc.rank == 0) mlp_prior_grad(m, q, g);                 // the prior part enters the rank-ordered sum once
    else for (int i = threadIdx.x; i < m.Dp; i += MLP_THREADS) g[i] = 0.0f;
    if (m.tc) { tc_pack_w1(m, q, tc); fence_async_smem(); }    // (generic accesses of the operand buffers precede the TMA writes)
    __syncthreads();
    TC_MARK(2);
    if (m.has_data) {
        const int rb = s < 0 ? 0 : m.sb[s], re = s < 0 ? m.N : m.sb[s + 1];
        if (m.tc) tc_backprop_rows(m, q, g, tile, tc, rb, re, s < 0 ? m.flat_base : m.tb[s], cc, next_s);
        else mlp_backprop_rows(m, q, g, tile, rb, re, cc);
        TC_MARK(14);
        // leave_partials: the caller's kick adds the ranks' partial gradients itself (one DSMEM pass instead of reduce +
        // write back + kick); all it needs here is that every rank's partial is complete
        if (!leave_partials) cluster_sum_vector<CS>(g, m.D);
        else if (CS > 1) cg::this_cluster().sync();
        TC_MARK(15);
    }
}

// l_prior = sum_i Normal(0, scale_i).log_prob(w_i).sum(), accumulated tensor by tensor (samplers.py:1153-1157)
__device__ __forceinline__ float mlp_log_prior(const MlpDev& m, const float* q, float* sred) {
    float l_prior = 0.0f;
    for (int t = 0; t < 2 * m.L; ++t) {
        const int l = t >> 1;
        const int off = (t & 1) ? m.boff[l] : m.woff[l];
        const int cnt = (t & 1) ? m.n[l + 1] : m.n[l] * m.n[l + 1];
        float s[1] = {0.0f};
        for (int i = threadIdx.x; i < cnt; i += MLP_THREADS) {
            const float w = q[off + i];
            // -((w - 0)**2) / (2*var) - log_scale - log(sqrt(2*pi))
            const float v = sub(sub(__fdiv_rn(-mul(w, w), m.two_var[t]), m.log_scale[t]), 0.9189385332046727f);
            s[0] = add(s[0], v);
        }
        block_sum<1>(s, sred);
        __syncthreads();
        l_prior = add(s[0], l_prior);
    }
    return l_prior;
}

// log p(q) = sum over splits of (ll_m + l_prior/prior_scale)  (hamiltonian's split loop, samplers.py:787-796);
// with s >= 0 only that split.  Optionally writes the network outputs (predict_model).
template <int CS>
__device__ __forceinline__ float mlp_log_prob(const MlpDev& m, const float* q, float* tile, float* sred, int s,
                                              float* pred_out, ClusterCtx cc, float* xslot, TcCtx& tc) {
    const float prior_term = __fdiv_rn(mlp_log_prior(m, q, sred), m.prior_scale);
    if (!m.has_data) return prior_term;
    TcEpi te;
    if (m.tc) {
        tc_pack_w1(m, q, tc);
        tc_epi_begin(m, q, te);
        fence_async_smem();
        __syncthreads();
    }
    float lp = 0.0f;
    const int s0 = s < 0 ? 0 : s, s1 = s < 0 ? m.M : s + 1;
    for (int sp = s0; sp < s1; ++sp) {
        float sse[1] = {0.0f};
        int ti = 0;
        for (int r0 = m.sb[sp]; r0 < m.sb[sp + 1]; r0 += m.T, ++ti) {
            if (ti % cc.size != cc.rank) continue;
            const int cnt = min(m.T, m.sb[sp + 1] - r0);
            if (m.tc) {
                float act[16];
                tc_prefetch_fwd(m, tile, tc, m.tb[sp] + ti, 0);
                tc_prefetch_y(m, tile, r0, cnt);
                tc_forward_tile(m, q, tile, tc, te, act, 0);
            } else {
                mlp_forward_tile(m, q, tile, r0, cnt);
            }
            sse[0] = add(sse[0], mlp_loss_tile(m, tile + m.aoff[m.L], nullptr, r0, cnt, m.sb[sp + 1] - m.sb[sp],
                                               pred_out != nullptr && m.loss == HMCX_LOSS_MULTICLASS_LOGSOFTMAX,
                                               m.tc ? tile + m.tc_yraw : nullptr));
            if (pred_out) {
                __syncthreads();
                const int nL = m.n[m.L];
                for (int i = threadIdx.x; i < cnt * nL; i += MLP_THREADS)
                    pred_out[(size_t)r0 * nL + i] = tile[m.aoff[m.L] + i];
            }
            __syncthreads();
        }
        block_sum<1>(sse, sred);
        __syncthreads();
        sse[0] = cluster_sum_scalar<CS>(sse[0], xslot);
        const float ll = mlp_ll_from_sum(m, sse[0], m.sb[sp + 1] - m.sb[sp]);
        lp = (sp == s0) ? add(ll, prior_term) : add(lp, add(ll, prior_term));
    }
    return lp;
}

// ---------------------------------------------------------------------------------------------------------
// persistent sample() kernel for the BNN path
// ---------------------------------------------------------------------------------------------------------
struct MlpRunArgs {
    MlpDev m;
    int scheme, mk, C, ld;
    const float* im;
    const float* sd;
    int rng_mode;
    uint64_t seed, chain_offset;
    const float* normals;
    const float* logu;
    const int32_t* perms;
    int nuts;
    double delta, mu;
    const double* table;
    double* h_bar;
    double* eps_bar;
    const float* eps_schedule;
    double eps0;                  // hmcx_nuts_t.step_size_init (0 = not given)
    float* eps_trace;
    const float* q_init;
    float* q_cur;
    float* eps;
    int L, S, burn, it0, it1;
    float* samples;
    uint8_t* accept;
    uint8_t* diverged;
    float* ham;
    int32_t* num_rejected;
    // stand-alone samplers.leapfrog with a SPLITTING integrator (:494-603): the momentum is given, the trajectory recorded
    const float* p_given;         // [C, ld] or NULL
    float* q_traj;                // [L, C, ld]: params after every step (ret_params)
    float* p_traj;                // [L, C, ld]: momentum after every step (ret_momenta)
};

template <int CS>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_run_kernel(const MlpRunArgs a) {
    extern __shared__ __align__(128) float sm[];
    __shared__ float sred[64];
    __shared__ float s_bcast[4];
    __shared__ float s_xchg;
    __shared__ int s_perm[HMCX_MLP_MAX_SPLITS];
    __shared__ __align__(8) uint64_t s_bars[5];
    __shared__ uint32_t s_tmem;

    const MlpDev& m = a.m;
    ClusterCtx cc = {0, 1};
    if (CS > 1) { cc.rank = (int)cg::this_cluster().block_rank(); cc.size = CS; }
    const bool lead = cc.rank == 0;                            // rank 0 owns every global-memory output
    const int c = blockIdx.x / CS, tid = threadIdx.x, D = m.D, M = m.M;
    float* q = sm;
    float* p = q + m.Dp;
    float* g = p + m.Dp;
    float* tile = sm + m.tile_base;
    const size_t row = (size_t)c * a.ld;
    const uint64_t chain_id = a.chain_offset + (uint64_t)c;
    TcCtx tc = {};
    if (m.tc) tc_init(tc, s_bars, &s_tmem);

    for (int i = tid; i < m.Dp; i += MLP_THREADS) { q[i] = i < D ? a.q_cur[row + i] : 0.0f; p[i] = 0.0f; g[i] = 0.0f; }
    __syncthreads();
    float lp_cur = a.p_given ? 0.0f : mlp_log_prob<CS>(m, q, tile, sred, -1, nullptr, cc, &s_xchg, tc);

    float eps = a.eps[c];
    double h_bar = 0.0, eps_bar = 1.0;
    if (a.nuts && tid == 0) { h_bar = a.h_bar[c]; eps_bar = a.eps_bar[c]; }
    int rejected = 0;
    const int keep = a.S - a.burn;
    float* const my_samples = (a.samples && lead) ? a.samples + (size_t)c * keep * a.ld : nullptr;
    if (a.it0 == 0 && my_samples)
        for (int i = tid; i < a.ld; i += MLP_THREADS) my_samples[i] = i < D ? q[i] : 0.0f;

    auto kinetic = [&]() {                                    // 2*K: p.p or p.(im*p)   (samplers.py:801, :814)
        float s[1] = {0.0f};
        for (int i = tid; i < D; i += MLP_THREADS)
            s[0] = add(s[0], a.mk == HMCX_MASS_DIAG ? mul(p[i], mul(a.im[i], p[i])) : mul(p[i], p[i]));
        block_sum<1>(s, sred);
        __syncthreads();
        return s[0];
    };
    // element-wise passes over the state: 16-byte accesses, VPT independent vectors per thread in flight (the inverse-mass
    // read of a drift comes from L2 and the partial gradients of a fused kick from a peer SM: latency-bound otherwise).
    // The padding lanes [D, Dp) of q / p / g are zero and stay zero (inv_mass lanes past D are read as 0).
    constexpr int VPT = 5;                                     // Dp/4 <= 512 * 5 vectors covers D <= 10240 in one round
    const int nvec = m.Dp >> 2;
    const bool im_aligned = (reinterpret_cast<size_t>(a.im) & 15) == 0;
    auto ld_im4 = [&](int v) {
        const int i = 4 * v;
        if (i + 3 < D && im_aligned) return __ldg(reinterpret_cast<const float4*>(a.im) + v);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < D) r.x = a.im[i];
        if (i + 1 < D) r.y = a.im[i + 1];
        if (i + 2 < D) r.z = a.im[i + 2];
        if (i + 3 < D) r.w = a.im[i + 3];
        return r;
    };
    // momentum += coef * grad; `twice`: the NEXT schedule position re-uses this gradient with no drift in between, its
    // kick (a second, separately rounded `+= coef2 * grad`) is applied in the same pass
    auto kick = [&](float coef, bool twice = false, float coef2 = 0.0f) {
        float4* p4 = reinterpret_cast<float4*>(p);
        const float4* g4 = reinterpret_cast<const float4*>(g);
        for (int v = tid; v < nvec; v += MLP_THREADS) {
            float4 pv = p4[v];
            const float4 gv = g4[v];
            pv.x = add(pv.x, mul(coef, gv.x)); pv.y = add(pv.y, mul(coef, gv.y));
            pv.z = add(pv.z, mul(coef, gv.z)); pv.w = add(pv.w, mul(coef, gv.w));
            if (twice) {
                pv.x = add(pv.x, mul(coef2, gv.x)); pv.y = add(pv.y, mul(coef2, gv.y));
                pv.z = add(pv.z, mul(coef2, gv.z)); pv.w = add(pv.w, mul(coef2, gv.w));
            }
            p4[v] = pv;
        }
        __syncthreads();
    };
    // momentum += coef * (g_rank0 + g_rank1 + ...): the cluster reduction of the split gradient FUSED into the kick -- every
    // rank reads all partials through distributed shared memory (16-byte loads), adds them in rank order (the same bits as
    // reduce-then-kick) and updates its replica of p.  The peers may overwrite their g only after everybody has read it:
    // barrier.cluster arrive here, wait after the drift that follows (its latency hides behind the drift).
    // `drift_cd` != 0: the drift that follows this kick (params += drift_cd * M^-1 momentum, the same operations as drift()
    // below on the just-updated momentum) in the same pass -- one sweep over the state and one barrier less per position
    auto kick_partials = [&](float coef, bool twice = false, float coef2 = 0.0f, float drift_cd = 0.0f) {
        cg::cluster_group cluster = cg::this_cluster();
        const float4* gr[CS];
#pragma unroll
        for (int r = 0; r < CS; ++r) gr[r] = reinterpret_cast<const float4*>(r == cc.rank ? g : cluster.map_shared_rank(g, r));
        float4* p4 = reinterpret_cast<float4*>(p);
        float4* q4 = reinterpret_cast<float4*>(q);
        const bool with_drift = drift_cd != 0.0f, with_im = with_drift && a.mk == HMCX_MASS_DIAG;
        constexpr int KV = CS >= 4 ? 2 : VPT;                  // (register budget: KV * CS vectors live)
        for (int v0 = tid; v0 < nvec; v0 += MLP_THREADS * KV) {
            float4 part[KV][CS], imv[KV];
#pragma unroll
            for (int k = 0; k < KV; ++k) {                     // all remote (and L2) loads of the round in flight together
                const int v = v0 + k * MLP_THREADS;
#pragma unroll
                for (int r = 0; r < CS; ++r) part[k][r] = v < nvec ? gr[r][v] : make_float4(0.f, 0.f, 0.f, 0.f);
                imv[k] = (with_im && v < nvec) ? ld_im4(v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                const int v = v0 + k * MLP_THREADS;
                if (v < nvec) {
                    float4 sg = part[k][0];
#pragma unroll
                    for (int r = 1; r < CS; ++r) {
                        sg.x = add(sg.x, part[k][r].x); sg.y = add(sg.y, part[k][r].y);
                        sg.z = add(sg.z, part[k][r].z); sg.w = add(sg.w, part[k][r].w);
                    }
                    float4 pv = p4[v];
                    pv.x = add(pv.x, mul(coef, sg.x)); pv.y = add(pv.y, mul(coef, sg.y));
                    pv.z = add(pv.z, mul(coef, sg.z)); pv.w = add(pv.w, mul(coef, sg.w));
                    if (twice) {
                        pv.x = add(pv.x, mul(coef2, sg.x)); pv.y = add(pv.y, mul(coef2, sg.y));
                        pv.z = add(pv.z, mul(coef2, sg.z)); pv.w = add(pv.w, mul(coef2, sg.w));
                    }
                    p4[v] = pv;
                    if (with_drift) {
                        float4 qv = q4[v];
                        if (with_im) {
                            qv.x = add(qv.x, mul(mul(drift_cd, imv[k].x), pv.x)); qv.y = add(qv.y, mul(mul(drift_cd, imv[k].y), pv.y));
                            qv.z = add(qv.z, mul(mul(drift_cd, imv[k].z), pv.z)); qv.w = add(qv.w, mul(mul(drift_cd, imv[k].w), pv.w));
                        } else {
                            qv.x = add(qv.x, mul(drift_cd, pv.x)); qv.y = add(qv.y, mul(drift_cd, pv.y));
                            qv.z = add(qv.z, mul(drift_cd, pv.z)); qv.w = add(qv.w, mul(drift_cd, pv.w));
                        }
                        q4[v] = qv;
                    }
                }
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        __syncthreads();
    };
    const bool fused_kick = CS > 1 && m.has_data;
    auto drift = [&](float coef) {                             // params += coef * M^-1 momentum
        float4* q4 = reinterpret_cast<float4*>(q);
        const float4* p4 = reinterpret_cast<const float4*>(p);
        if (a.mk == HMCX_MASS_DIAG) {
            for (int v0 = tid; v0 < nvec; v0 += MLP_THREADS * VPT) {
                float4 im[VPT];
#pragma unroll
                for (int k = 0; k < VPT; ++k) {                // the L2 reads of the round in flight together
                    const int v = v0 + k * MLP_THREADS;
                    im[k] = v < nvec ? ld_im4(v) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < VPT; ++k) {
                    const int v = v0 + k * MLP_THREADS;
                    if (v < nvec) {
                        float4 qv = q4[v];
                        const float4 pv = p4[v];
                        qv.x = add(qv.x, mul(mul(coef, im[k].x), pv.x)); qv.y = add(qv.y, mul(mul(coef, im[k].y), pv.y));
                        qv.z = add(qv.z, mul(mul(coef, im[k].z), pv.z)); qv.w = add(qv.w, mul(mul(coef, im[k].w), pv.w));
                        q4[v] = qv;
                    }
                }
            }
        } else {
            for (int v = tid; v < nvec; v += MLP_THREADS) {
                float4 qv = q4[v];
                const float4 pv = p4[v];
                qv.x = add(qv.x, mul(coef, pv.x)); qv.y = add(qv.y, mul(coef, pv.y));
                qv.z = add(qv.z, mul(coef, pv.z)); qv.w = add(qv.w, mul(coef, pv.w));
                q4[v] = qv;
            }
        }
        __syncthreads();
    };

    // split drifts divide the step size as a DOUBLE (step_size/K_div, samplers.py:513, :558): the run's initial Python
    // double while the chain still uses it, the fp32 value once dual averaging produced one (:668)
    auto eps_double = [&](float e) { return (a.eps0 != 0.0 && e == (float)a.eps0) ? a.eps0 : (double)e; };

    int g_last_sp = -3;
    bool g_fresh = false;
    for (int n = a.it0; n < a.it1; ++n) {
        if (a.eps_schedule) eps = a.eps_schedule[(size_t)n * a.C + c];
        const float half = mul(0.5f, eps);
        // ---- gibbs (or the caller's momentum) ----
        if (a.p_given) for (int i = tid; i < D; i += MLP_THREADS) p[i] = a.p_given[row + i];
        else for (int v = tid; 4 * v < a.ld; v += MLP_THREADS) {
            float z[4];
            if (a.rng_mode == HMCX_RNG_INJECTED) ld4_stream(a.normals + ((size_t)(n - a.it0) * a.C + c) * a.ld + 4 * v, z);
            else philox_normal4(a.seed, chain_id, (uint64_t)n, (uint32_t)v, z);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * v + j;
                if (i < D) p[i] = a.mk == HMCX_MASS_DIAG ? mul(z[j], a.sd[i]) : z[j];
            }
        }
        if (a.scheme == HMCX_SCHEME_SPLIT_RAND && tid == 0) {            // idx = randperm(M), once per trajectory (:550)
            if (a.rng_mode == HMCX_RNG_INJECTED) {
                for (int s = 0; s < M; ++s) s_perm[s] = a.perms[((size_t)(n - a.it0) * a.C + c) * M + s];
            } else {
                for (int s = 0; s < M; ++s) s_perm[s] = s;
                for (int s = M - 1; s > 0; --s) {                        // Fisher-Yates on the PERM stream
                    const uint4 r = philox_draw(a.seed, chain_id, (uint64_t)n, (uint32_t)s, STREAM_PERM);
                    const int j = (int)(r.x % (uint32_t)(s + 1));
                    const int t = s_perm[s]; s_perm[s] = s_perm[j]; s_perm[j] = t;
                }
            }
        }
        __syncthreads();
        const float kin0 = a.p_given ? 0.0f : kinetic();
        // ---- trajectory: every integrator schedule is a sequence of steps [drift] grad(split) kick [drift] ----
        // ONE loop around ONE inlined copy of the gradient evaluation (nine call sites, one per schedule position, made
        // each instantiation of this kernel ~100k instructions = 1.6 MB of code and minutes of ptxas; an out-of-line copy
        // was measured 1.6x slower: the model descriptor then lives behind a pointer instead of in the constant bank).
        //   PLAIN (:281-302)      t = 0: grad, kick(eps/2); t = 1..L: drift(eps), grad, kick(eps); finally kick(-eps/2)
        //   SPLITTING (:499-540)  per step j = 0..2M-1: s = j < M ? j : 2M-1-j; grad(s), kick(eps/2), drift(eps/(2(M-1))) unless
        //                         the sweep's last split
        //   SPLITTING_RAND (:551-568)  per step, split s = perm[j/2]: grad, kick(eps/2), drift(eps/M) | grad, kick(eps/2)
        //   SPLITTING_KMID (:579-598)  M kicks up, drift(eps), M kicks down
        {
            const int twoM = 2 * M;
            const bool plain = a.scheme == HMCX_SCHEME_PLAIN;
            const int T = plain ? a.L + 1 : a.L * twoM;
            float cd = 0.0f;
            if (a.scheme == HMCX_SCHEME_SPLIT_SYM) cd = (float)(eps_double(eps) / (double)((M - 1) * 2));
            else if (a.scheme == HMCX_SCHEME_SPLIT_RAND) cd = (float)(eps_double(eps) / (double)M);
            else if (a.scheme == HMCX_SCHEME_SPLIT_KMID) cd = eps;
            int jj = 0;
            auto split_at = [&](int j) {                                  // the data split of schedule position j
                const int up = j < M ? j : twoM - 1 - j;
                return a.scheme == HMCX_SCHEME_SPLIT_RAND ? s_perm[j >> 1] : up;
            };
            auto post_at = [&](int j) {                                   // a drift follows the kick of schedule position j
                const int up = j < M ? j : twoM - 1 - j;
                if (a.scheme == HMCX_SCHEME_SPLIT_SYM) return j < M ? (up < M - 1) : (up > 0);
                if (a.scheme == HMCX_SCHEME_SPLIT_RAND) return (j & 1) == 0;
                return j == M - 1;
            };
            // Two consecutive schedule positions with the SAME split and NO drift between them differentiate the same function
            // at the same parameters: the turning point of the symmetric sweep (m = M-1 up, then M-1 down, :501-:523) and the
            // step boundary (m = 0 down, then m = 0 up of the next step); SPLITTING_KMID has the latter.  The reference calls
            // autograd twice and gets the same tensor twice; here the second evaluation is skipped and g (per-rank partials
            // included) is kicked again -- the same bits, 2 of the 2M evaluations of a symmetric step saved.
            // The same across iterations: a trajectory ends with an evaluation at its final parameters and no drift after it;
            // when the proposal is accepted the next trajectory starts by differentiating the same split (the whole potential
            // for PLAIN) at those parameters -- g_last_sp / g_fresh carry the gradient over (1 of L+1 evaluations of
            // sample_model's plain leapfrog at high acceptance).
            int prev_sp = g_fresh ? g_last_sp : -3;
            bool prev_post = !g_fresh, kicked_ahead = false;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                int sp = -1, sp_next = -1;
                float kc = half;
                bool post = false, reuse = false, kick_twice = false, drift_done = false;
                if (plain) {
                    if (t > 0) { drift(eps); kc = eps; }
                    else reuse = g_fresh && g_last_sp == -1;
                } else {
                    sp = split_at(jj);
                    post = post_at(jj);
                    reuse = sp == prev_sp && !prev_post;
                    if (++jj == twoM) jj = 0;
                    sp_next = split_at(jj);
                    int ahead = 1;
                    if (sp_next == sp && !post) {                         // the next position re-uses this gradient: prefetch for
                        sp_next = split_at(jj + 1 == twoM ? 0 : jj + 1);  // the one after it
                        ahead = 2;
                        // ... and its kick joins this one -- unless a trajectory is being recorded and a leapfrog step
                        // ends between the two (the recorded momentum is the one after the first kick only)
                        kick_twice = !reuse && t + 1 < T && !(a.q_traj && jj == 0);
                    }
                    if (t + ahead >= T) sp_next = -2;
                    prev_sp = sp; prev_post = post;
                }
                if (t + 1 == T) sp_next = -2;                             // the Hamiltonian evaluation follows: nothing to prefetch
                if (!reuse) mlp_grad_split<CS>(m, q, g, tile, sp, cc, tc, fused_kick, sp_next);
                if (kicked_ahead) {                                       // this position's kick was applied with the previous one
                    kicked_ahead = false;
                    if (fused_kick) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                } else if (fused_kick) {
                    // (plain: the drift belongs to the NEXT position and takes its own pass; a recorded trajectory keeps the
                    // separate pass too -- nothing to gain there)
                    drift_done = post && !plain && cd != 0.0f;
                    kick_partials(kc, kick_twice, half, drift_done ? cd : 0.0f);
                } else {
                    kick(kc, kick_twice, half);
                }
                kicked_ahead = kick_twice;
                TC_MARK(16);
                if (post && !drift_done) drift(cd);
                if (fused_kick) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
                TC_MARK(17);
                if (a.q_traj && !plain && jj == 0 && lead) {              // a leapfrog step just ended (:546-547, :570-571, :600-601)
                    const size_t o = ((size_t)(t / twoM) * a.C + c) * a.ld;
                    for (int i = tid; i < a.ld; i += MLP_THREADS) {
                        a.q_traj[o + i] = i < D ? q[i] : 0.0f;
                        a.p_traj[o + i] = i < D ? p[i] : 0.0f;
                    }
                }
            }
            if (plain) {                                                  // p - half*g == p + (-half)*g exactly
                if (fused_kick) { kick_partials(-half); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
                else kick(-half);
            }
            // g now holds the gradient of the last schedule position's split at the proposal, unless a drift followed it
            g_last_sp = plain ? -1 : prev_sp;
            g_fresh = plain || !prev_post;
        }
        if (a.p_given) break;                                             // stand-alone leapfrog: no Hamiltonian, no MH
        // ---- Hamiltonians + MH ----
        const float lp_new = mlp_log_prob<CS>(m, q, tile, sred, -1, nullptr, cc, &s_xchg, tc);
        const float kin1 = kinetic();
        const float h_old = add(-lp_cur, mul(0.5f, kin0));
        const float h_new = add(-lp_new, mul(0.5f, kin1));
        const bool bad = !finite_f(lp_cur) || !finite_f(lp_new);
        const float x = add(-h_new, h_old);
        const float rho = (x < 0.0f) ? x : 0.0f;
        if (tid == 0)
            s_bcast[0] = (a.rng_mode == HMCX_RNG_INJECTED) ? a.logu[(size_t)(n - a.it0) * a.C + c]
                                                           : philox_log_uniform(a.seed, chain_id, (uint64_t)n);
        __syncthreads();
        const float logu = s_bcast[0];
        const bool acc = !bad && (rho >= logu);
        if (acc) {
            lp_cur = lp_new;
            if (lead) for (int i = tid; i < D; i += MLP_THREADS) a.q_cur[row + i] = q[i];
        } else {
            ++rejected;
            g_fresh = false;                                              // q goes back: g belongs to the rejected proposal
            const float* src = (n == a.burn + 1) ? a.q_init : a.q_cur;    // the first-stored-iteration quirk (:1018)
            for (int i = tid; i < D; i += MLP_THREADS) q[i] = src[row + i];
            __syncthreads();
            if (n == a.burn + 1) {
                lp_cur = mlp_log_prob<CS>(m, q, tile, sred, -1, nullptr, cc, &s_xchg, tc);
                if (lead) for (int i = tid; i < D; i += MLP_THREADS) a.q_cur[row + i] = q[i];
            }
        }
        if (CS > 1) cg::this_cluster().sync();                 // q_cur is stable before any rank re-reads it
        if (n > a.burn && my_samples) {
            float* dst = my_samples + (size_t)(n - a.burn) * a.ld;
            for (int i = tid; i < a.ld; i += MLP_THREADS) dst[i] = i < D ? q[i] : 0.0f;
        }
        if (tid == 0 && lead) {
            const size_t o = (size_t)c * a.S + n;
            if (a.accept) a.accept[o] = acc ? 1 : 0;
            if (a.diverged) a.diverged[o] = bad ? 1 : 0;
            if (a.ham) { a.ham[2 * o] = h_old; a.ham[2 * o + 1] = h_new; }
        }
        if (a.nuts && n <= a.burn) {                                       // dual averaging, as hmc_run_kernel
            if (tid == 0) {
                float e = eps;
                if (n < a.burn || bad) {
                    const double* T = a.table + 5 * (size_t)n;
                    const double alpha = bad ? 0.0 : (double)expf(rho);
                    h_bar = __dadd_rn(__dmul_rn(T[0], h_bar), __dmul_rn(T[1], a.delta - alpha));
                    const double x_new = a.mu - __dmul_rn(T[2], h_bar);
                    e = expf((float)x_new);
                    const float xb = add((float)__dmul_rn(T[3], x_new), mul((float)T[4], logf((float)eps_bar)));
                    eps_bar = (double)expf(xb);
                }
                if (n == a.burn) e = (float)eps_bar;
                s_bcast[1] = e;
                if (a.eps_trace && lead) a.eps_trace[(size_t)c * a.S + n] = e;
            }
            __syncthreads();
            eps = s_bcast[1];
        } else if (a.eps_trace && tid == 0 && lead) {
            a.eps_trace[(size_t)c * a.S + n] = eps;
        }
        __syncthreads();
    }
    if (tid == 0 && lead && !a.p_given) {
        a.eps[c] = eps;
        if (a.nuts) { a.h_bar[c] = h_bar; a.eps_bar[c] = eps_bar; }
        if (a.num_rejected) a.num_rejected[c] += rejected;
    }
    if (m.tc) tc_fini(tc);
}

// gradient / log-prob of C parameter vectors (collect_gradients mirror, also the unit-test hook of the backward pass)
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_grad_kernel(const MlpDev m, const float* __restrict__ qin, int ld, int split, float* __restrict__ gout,
                float* __restrict__ lpout) {
    extern __shared__ __align__(128) float sm[];
    __shared__ float sred[64];
    __shared__ __align__(8) uint64_t s_bars[5];
    __shared__ uint32_t s_tmem;
    float* q = sm;
    float* g = q + m.Dp;
    float* tile = sm + m.tile_base;
    const size_t row = (size_t)blockIdx.x * ld;
    TcCtx tc = {};
    if (m.tc) tc_init(tc, s_bars, &s_tmem);
    for (int i = threadIdx.x; i < m.Dp; i += MLP_THREADS) { q[i] = i < m.D ? qin[row + i] : 0.0f; g[i] = 0.0f; }
    __syncthreads();
    if (gout) {
        mlp_grad_split<1>(m, q, g, tile, split, ClusterCtx{0, 1}, tc);
        __syncthreads();
        for (int i = threadIdx.x; i < ld; i += MLP_THREADS) gout[row + i] = i < m.D ? g[i] : 0.0f;
    }
    if (lpout) {
        const float lp = mlp_log_prob<1>(m, q, tile, sred, split, nullptr, ClusterCtx{0, 1}, nullptr, tc);
        if (threadIdx.x == 0) lpout[blockIdx.x] = lp;
    }
    if (m.tc) tc_fini(tc);
}

// predict_model: one CTA per posterior sample
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_predict_kernel(const MlpDev m, const float* __restrict__ samples, int ld, float* __restrict__ pred,
                   float* __restrict__ lpout) {
    extern __shared__ __align__(128) float sm[];
    __shared__ float sred[64];
    __shared__ __align__(8) uint64_t s_bars[5];
    __shared__ uint32_t s_tmem;
    float* q = sm;
    float* tile = sm + m.tile_base;
    const size_t row = (size_t)blockIdx.x * ld;
    for (int i = threadIdx.x; i < m.Dp; i += MLP_THREADS) q[i] = i < m.D ? samples[row + i] : 0.0f;
    __syncthreads();
    float* my_pred = pred + (size_t)blockIdx.x * m.N * m.n[m.L];
    TcCtx tc = {};
    if (m.tc) tc_init(tc, s_bars, &s_tmem);
    const float lp = mlp_log_prob<1>(m, q, tile, sred, -1, my_pred, ClusterCtx{0, 1}, nullptr, tc);
    if (threadIdx.x == 0 && lpout) lpout[blockIdx.x] = lp;
    if (m.tc) tc_fini(tc);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static void mlp_layout_tiles(MlpDev& m, int T) {
    int aoff = 0, maxw = 0;
    for (int l = 0; l <= m.L; ++l) {
        m.aoff[l] = aoff;
        aoff += T * m.n[l];
        if (m.n[l] > maxw) maxw = m.n[l];
    }
    m.dzoff[0] = aoff;
    m.dzoff[1] = aoff + T * maxw;
    m.tile_floats = aoff + 2 * T * maxw;
    m.T = T;
}

static bool mlp_tc_shape(const MlpDev& m) {
    return m.L == 2 && m.n[1] == TC_H && m.n[0] >= 16 && m.n[0] <= 64 && (m.n[0] & 15) == 0 && m.n[2] <= TC_NLMAX && m.has_data;
}

// tensor-core layout of the tile area (one-hidden-layer stacks n0 -> 128 -> nL, see the tcgen05 section above)
static bool mlp_layout_tc(MlpDev& m, int state_vectors) {
    if (!mlp_tc_shape(m) || !m.xp) return false;
    const int n0 = m.n[0];
    int off = 0;
    m.tc_f0 = off; off += 2 * TC_TR * n0;       // forward X operand (hi|lo), double-buffered; f0 | f1 also stage dW1
    m.tc_f1 = off; off += 2 * TC_TR * n0;       //   (128 rows, pitch n0 + 4) at the end of an evaluation
    m.tc_b = off; off += 2 * TC_TR * n0;        // backward X operand (hi|lo)
    m.tc_part = off; off += 4 * TC_H * (1 + TC_NLMAX);
    m.tc_yraw = off; off += TC_TR * TC_NLMAX;
    m.aoff[0] = m.aoff[1] = 0;
    m.aoff[2] = off; off += TC_TR * TC_NLMAX;
    m.dzoff[0] = m.dzoff[1] = off; off += TC_TR * TC_NLMAX;
    m.tile_floats = off;
    m.tile_base = (state_vectors * m.Dp + 31) / 32 * 32;      // 128-byte aligned operand buffers
    m.T = TC_TR;
    m.tc = 1;
    return (size_t)(m.tile_base + m.tile_floats) * sizeof(float) <= 227 * 1024 - 2048;
}

// pick the largest tile height whose buffers fit next to `state_vectors` copies of the parameter vector
static bool mlp_pick_tile(MlpDev& m, int state_vectors, bool allow_tc = false) {
    if (allow_tc && mlp_layout_tc(m, state_vectors)) return true;
    m.tc = 0;
    m.tile_base = state_vectors * m.Dp;
    for (int T = MLP_T_MAX; T >= 8; T >>= 1) {
        mlp_layout_tiles(m, T);
        if ((size_t)(state_vectors * m.Dp + m.tile_floats) * sizeof(float) <= 227 * 1024 - 4096) return true;
    }
    return false;
}

// packed-operand tile numbering (see mlp_pack_x_kernel): the tiles of every split, then -- only when some interior split
// boundary is not a multiple of the tile height -- the tiles of the all-rows tiling; returns the number of tiles
static int mlp_packed_tiles(MlpDev& m) {
    int t = 0;
    bool aligned = true;
    for (int s = 0; s < m.M; ++s) {
        m.tb[s] = t;
        t += (m.sb[s + 1] - m.sb[s] + TC_TR - 1) / TC_TR;
        if (s > 0 && (m.sb[s] % TC_TR) != 0) aligned = false;
    }
    m.tb[m.M] = t;
    if (aligned) { m.flat_base = 0; return t; }
    m.flat_base = t;
    return t + (m.N + TC_TR - 1) / TC_TR;
}

static int fill_mlp(const hmcx_target_t* target, MlpDev& m) {
    if (!target || target->kind != HMCX_TARGET_MLP || !target->mlp) return HMCX_ERR_INVALID_ARG;
    const hmcx_mlp_t& h = *target->mlp;
    if (h.num_layers < 1 || h.num_layers > HMCX_MLP_MAX_LAYERS) return HMCX_ERR_INVALID_ARG;
    if (h.loss < HMCX_LOSS_REGRESSION || h.loss > HMCX_LOSS_MULTICLASS_LOGSOFTMAX) return HMCX_ERR_UNSUPPORTED;
    m.loss = h.loss;
    if (h.activation[h.num_layers - 1] != HMCX_ACT_NONE) return HMCX_ERR_INVALID_ARG;
    m.L = h.num_layers;
    int off = 0;
    for (int l = 0; l <= m.L; ++l) {
        if (h.widths[l] < 1) return HMCX_ERR_INVALID_ARG;
        m.n[l] = h.widths[l];
    }
    for (int l = 0; l < m.L; ++l) {
        m.act[l] = h.activation[l];
        if (m.act[l] < 0 || m.act[l] > HMCX_ACT_SIGMOID) return HMCX_ERR_INVALID_ARG;
        m.woff[l] = off; off += m.n[l] * m.n[l + 1];
        m.boff[l] = off; off += m.n[l + 1];
    }
    m.D = off;
    if (m.D != target->dim) return HMCX_ERR_INVALID_ARG;
    m.Dp = (m.D + 3) / 4 * 4;
    mlp_layout_tiles(m, 8);
    m.tau_out = h.tau_out;
    m.prior_scale = h.prior_scale;
    m.c_ll = (h.loss == HMCX_LOSS_REGRESSION) ? (float)(-0.5 * (double)h.tau_out) : (float)(-(double)h.tau_out);
    for (int t = 0; t < 2 * m.L; ++t) {
        m.two_var[t] = h.prior_two_var[t]; m.log_scale[t] = h.prior_log_scale[t]; m.gcoef[t] = h.prior_grad_coef[t];
    }
    m.x = h.x; m.y = h.y; m.N = h.num_rows;
    m.has_data = (h.x != nullptr) ? 1 : 0;
    m.M = h.num_splits;
    if (m.M < 1 || m.M > HMCX_MLP_MAX_SPLITS) return HMCX_ERR_INVALID_ARG;
    if (m.has_data) {
        if (!h.y || h.num_rows < 1) return HMCX_ERR_INVALID_ARG;
        if (h.split_begin[0] != 0 || h.split_begin[m.M] != h.num_rows) return HMCX_ERR_INVALID_ARG;
        for (int s = 0; s <= m.M; ++s) {
            m.sb[s] = h.split_begin[s];
            if (s && m.sb[s] <= m.sb[s - 1]) return HMCX_ERR_INVALID_ARG;
        }
    } else {
        for (int s = 0; s <= m.M; ++s) m.sb[s] = 0;
    }
    mlp_packed_tiles(m);
    m.xp = h.x_packed;
    return HMCX_OK;
}

static inline int cuda_status() { return cudaGetLastError() == cudaSuccess ? HMCX_OK : HMCX_ERR_CUDA; }

template <typename Kern>
static int prepare_smem(Kern kern, size_t bytes) {
    if (bytes > 227 * 1024) return HMCX_ERR_UNSUPPORTED;       // the chain state does not fit one SM's shared memory
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return HMCX_ERR_UNSUPPORTED;
    }
    return HMCX_OK;
}

int mlp_split_run(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng, const hmcx_nuts_t* nuts,
                  int scheme, const float* q_init, float* q_cur, float* eps, int C, int ld, int L, int S, int burn,
                  int it0, int it1, float* samples, uint8_t* accept, uint8_t* diverged, float* ham,
                  int32_t* num_rejected, cudaStream_t st, const float* p_given, float* q_traj, float* p_traj) {
    MlpRunArgs a = {};
    a.p_given = p_given; a.q_traj = q_traj; a.p_traj = p_traj;
    int rc = fill_mlp(target, a.m);
    if (rc != HMCX_OK) return rc;
    const int mk = mass ? mass->kind : HMCX_MASS_NONE;
    if (mk != HMCX_MASS_NONE && mk != HMCX_MASS_DIAG) return HMCX_ERR_UNSUPPORTED;
    if (mk == HMCX_MASS_DIAG && (!mass->inv_mass || !mass->mass_factor)) return HMCX_ERR_INVALID_ARG;
    if (!rng || !q_init || !q_cur || !eps || C < 1 || ld < a.m.D || (ld & 3) || L < 1 || S < 1 || burn < 0 || burn >= S ||
        it0 < 0 || it1 > S || it0 > it1)
        return HMCX_ERR_INVALID_ARG;
    if (scheme < HMCX_SCHEME_PLAIN || scheme > HMCX_SCHEME_SPLIT_KMID) return HMCX_ERR_INVALID_ARG;
    if ((scheme == HMCX_SCHEME_SPLIT_SYM || scheme == HMCX_SCHEME_SPLIT_KMID) && a.m.M < 2) return HMCX_ERR_INVALID_ARG;  // :497-498
    if (rng->mode == HMCX_RNG_INJECTED) {
        if (!p_given && (!rng->normals || !rng->log_uniforms)) return HMCX_ERR_INVALID_ARG;
        if (scheme == HMCX_SCHEME_SPLIT_RAND && !rng->perms) return HMCX_ERR_INVALID_ARG;
    } else if (rng->mode != HMCX_RNG_PHILOX) {
        return HMCX_ERR_INVALID_ARG;
    }
    a.scheme = scheme; a.mk = mk; a.C = C; a.ld = ld;
    a.im = mass ? mass->inv_mass : nullptr; a.sd = mass ? mass->mass_factor : nullptr;
    a.rng_mode = rng->mode; a.seed = rng->seed; a.chain_offset = rng->chain_offset;
    a.normals = rng->normals; a.logu = rng->log_uniforms; a.perms = rng->perms;
    a.nuts = (nuts && nuts->enabled) ? 1 : 0;
    a.eps0 = nuts ? nuts->step_size_init : 0.0;
    if (a.nuts) {
        if (!nuts->table || !nuts->h_bar || !nuts->eps_bar || burn < 1) return HMCX_ERR_INVALID_ARG;
        a.delta = nuts->desired_accept_rate; a.mu = nuts->mu; a.table = nuts->table;
        a.h_bar = nuts->h_bar; a.eps_bar = nuts->eps_bar;
        a.eps_schedule = nuts->eps_schedule; a.eps_trace = nuts->eps_trace;
    }
    a.q_init = q_init; a.q_cur = q_cur; a.eps = eps; a.L = L; a.S = S; a.burn = burn; a.it0 = it0; a.it1 = it1;
    a.samples = samples; a.accept = accept; a.diverged = diverged; a.ham = ham; a.num_rejected = num_rejected;
    if (!mlp_pick_tile(a.m, 3, target->mlp->tensor_cores != HMCX_MLP_TC_OFF))
        return HMCX_ERR_UNSUPPORTED;                            // q, p, g do not fit one SM's shared memory
    const size_t smem = (size_t)(a.m.tile_base + a.m.tile_floats) * sizeof(float);
    // CTAs per chain (thread-block cluster size): at most the tiles of the smallest split, at most 4, and -- unless the
    // caller pins it (hmcx_mlp_t.cluster_size) -- no more than keeps every chain's cluster resident at once.  The
    // partial gradients are associated per cluster rank, so low-order bits depend on this number: pin it to make
    // chains bit-reproducible across launches with different chain counts (e.g. different multi-GPU shardings).
    int cs = 1;
    if (a.m.has_data && a.m.D <= MLP_THREADS * 36) {
        int min_tiles = 1 << 30;
        for (int s = 0; s < a.m.M; ++s) min_tiles = min(min_tiles, (a.m.sb[s + 1] - a.m.sb[s] + a.m.T - 1) / a.m.T);
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int pinned = target->mlp->cluster_size;
        while (cs * 2 <= 4 && cs * 2 <= min_tiles && (pinned ? cs * 2 <= pinned : C * cs * 2 <= sms)) cs *= 2;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * cs);
    cfg.blockDim = dim3(MLP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t err;
    if (cs == 4) { rc = prepare_smem(mlp_run_kernel<4>, smem); if (rc != HMCX_OK) return rc; err = cudaLaunchKernelEx(&cfg, mlp_run_kernel<4>, a); }
    else if (cs == 2) { rc = prepare_smem(mlp_run_kernel<2>, smem); if (rc != HMCX_OK) return rc; err = cudaLaunchKernelEx(&cfg, mlp_run_kernel<2>, a); }
    else { rc = prepare_smem(mlp_run_kernel<1>, smem); if (rc != HMCX_OK) return rc; err = cudaLaunchKernelEx(&cfg, mlp_run_kernel<1>, a); }
    if (err != cudaSuccess) { cudaGetLastError(); return HMCX_ERR_CUDA; }
    return cuda_status();
}

int mlp_grad_log_prob(const hmcx_target_t* target, const float* q, int C, int ld, int split, float* grad_out,
                      float* log_prob_out, cudaStream_t st) {
    MlpDev m = {};
    int rc = fill_mlp(target, m);
    if (rc != HMCX_OK) return rc;
    if (!q || C < 1 || ld < m.D || (ld & 3) || split < -1 || split >= m.M || (!grad_out && !log_prob_out))
        return HMCX_ERR_INVALID_ARG;
    if (!mlp_pick_tile(m, 2, target->mlp->tensor_cores != HMCX_MLP_TC_OFF)) return HMCX_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(m.tile_base + m.tile_floats) * sizeof(float);
    rc = prepare_smem(mlp_grad_kernel, smem);
    if (rc != HMCX_OK) return rc;
    mlp_grad_kernel<<<C, MLP_THREADS, smem, st>>>(m, q, ld, split, grad_out, log_prob_out);
    return cuda_status();
}

int mlp_predict(const hmcx_target_t* target, const float* samples, int S, int ld, float* pred_out, float* log_prob_out,
                cudaStream_t st) {
    MlpDev m = {};
    int rc = fill_mlp(target, m);
    if (rc != HMCX_OK) return rc;
    if (!samples || !pred_out || S < 1 || ld < m.D || (ld & 3) || !m.has_data) return HMCX_ERR_INVALID_ARG;
    if (!mlp_pick_tile(m, 1, target->mlp->tensor_cores != HMCX_MLP_TC_OFF)) return HMCX_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(m.tile_base + m.tile_floats) * sizeof(float);
    rc = prepare_smem(mlp_predict_kernel, smem);
    if (rc != HMCX_OK) return rc;
    mlp_predict_kernel<<<S, MLP_THREADS, smem, st>>>(m, samples, ld, pred_out, log_prob_out);
    return cuda_status();
}

// stand-alone samplers.leapfrog with Integrator.SPLITTING / _RAND / _KMID (:494-603) over C chains: one "iteration" of the run
// kernel with the momentum given, no Hamiltonian and no MH; q_traj / p_traj [L, C, ld] take the state after every step
int mlp_leapfrog(const hmcx_target_t* target, const hmcx_mass_t* mass, const hmcx_rng_t* rng, int scheme, double step_size,
                 const float* q_in, const float* p_in, float* eps, int C, int ld, int L, float* q_traj, float* p_traj,
                 cudaStream_t st) {
    if (!p_in || !q_traj || !p_traj || scheme == HMCX_SCHEME_PLAIN) return HMCX_ERR_INVALID_ARG;
    hmcx_nuts_t no_nuts = {};
    no_nuts.step_size_init = step_size;                        // the Python double the drifts divide (:513, :558)
    return mlp_split_run(target, mass, rng, &no_nuts, scheme, q_in, const_cast<float*>(q_in), eps, C, ld, L, 1, 0, 0, 1,
                         nullptr, nullptr, nullptr, nullptr, nullptr, st, p_in, q_traj, p_traj);
}

// packed X operands of the tensor-core path (hmcx_mlp_t.x_packed): size in floats (0: the stack does not use it) / build
size_t mlp_packed_x_floats(const hmcx_target_t* target) {
    MlpDev m = {};
    if (fill_mlp(target, m) != HMCX_OK || !mlp_tc_shape(m)) return 0;
    return (size_t)mlp_packed_tiles(m) * 4 * TC_TR * m.n[0];
}

int mlp_pack_x(const hmcx_target_t* target, float* out, cudaStream_t st) {
    MlpDev m = {};
    const int rc = fill_mlp(target, m);
    if (rc != HMCX_OK) return rc;
    if (!out) return HMCX_ERR_INVALID_ARG;
    if (!mlp_tc_shape(m)) return HMCX_ERR_UNSUPPORTED;
    mlp_pack_x_kernel<<<mlp_packed_tiles(m), 256, 0, st>>>(m, out);
    return cuda_status();
}

#ifdef HMCX_TC_PROF
extern "C" int hmcx_debug_tc_prof(long long* out) {           // developer build only (scripts/prof_tc_phases.py)
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_tc_prof_n, sizeof(int));
    cudaMemcpyFromSymbol(out, g_tc_prof, sizeof(long long) * 512);
    int zero = 0;
    cudaMemcpyToSymbol(g_tc_prof_n, &zero, sizeof(int));
    return n;
}
#endif

}  // namespace hmcx
