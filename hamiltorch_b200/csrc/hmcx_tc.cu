// hmcx_tc.cu -- 5th-generation tensor-core (tcgen05 / TMEM) building block for the dense contractions of the path:
//     D[M x N] = A[M x K] . B[N x K]^T        (A, B, D row-major fp32; fp32-accurate via 3xTF32 split operands)
// This is the GEMM behind full-covariance Gaussian targets / full mass matrices at large D
// (grad log p = -(q - mu) P for ALL chains at once is exactly this contraction with M = chains, N = K = D,
// samplers.py:294, :812 and targets.GaussianFull) -- SURVEY.md section 8f item 2.
//
// Blackwell mechanics (one CTA per 128 x 128 output tile, 128 threads):
//   * operands are staged by the CTA's threads from global memory into shared memory in the canonical UMMA
//     K-major / no-swizzle layout (8-row x 16-byte core matrices), each fp32 split into tf32 hi + tf32 lo;
//   * ONE elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8 per instruction), three per
//     k-step: hi*hi + hi*lo + lo*hi, accumulating in fp32 in TENSOR MEMORY (128 lanes x 128 columns);
//   * completion is tracked with tcgen05.commit -> mbarrier; the epilogue reads the accumulator back with
//     tcgen05.ld (32x32b.x32: warp w owns TMEM lanes 32w..32w+31) and stores rows to global memory.
// Descriptor encodings follow the CUTLASS sm100 definitions (cute/arch/mma_sm100_desc.hpp: SmemDescriptor,
// InstrDescriptor) -- re-derived here, no CUTLASS code is used.
#include "hmcx_common.cuh"

namespace hmcx {

constexpr int TC_M = 128, TC_N = 128, TC_KC = 32;        // CTA tile and K chunk (fp32 elements)
constexpr int TC_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {      // round-to-nearest tf32, returned in an fp32 container
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// 64-bit shared-memory matrix descriptor: K-major, SWIZZLE_NONE.  Fields (16-byte units): start address [0,14),
// leading byte offset = distance between the two core matrices along K [16,30), stride byte offset = distance
// between 8-row groups [32,46), descriptor version 1 (Blackwell) [46,48), layout type 0 [61,64).
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    while (!mbar_try_wait(mbar, parity)) {}
}

// Stage a [128 rows x KC] fp32 slab (row stride `ld`) into the canonical layout, split into tf32 hi and lo copies.
// Core matrix (rg = row/8, kc = k/4) lives at ((kc * 16 + rg) * 128) bytes: SBO = 128 B, LBO = 16*128 = 2048 B.
__device__ __forceinline__ void stage_operand(const float* __restrict__ g, int ld, float* s_hi, float* s_lo) {
    // 128 rows x 8 float4 per row = 1024 float4; 128 threads -> 8 each.  Thread t handles row r = t, all 8 chunks:
    // consecutive threads touch consecutive rows => conflict-free 16-byte shared stores (row stride 16 B in a core).
    const int r = threadIdx.x;
    const float4* src = reinterpret_cast<const float4*>(g + (size_t)r * ld);
#pragma unroll
    for (int kc = 0; kc < TC_KC / 4; ++kc) {
        const float4 v = __ldg(src + kc);
        float4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y); l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
        const int off = ((kc * 16 + (r >> 3)) * 128 + (r & 7) * 16) >> 2;     // in floats
        *reinterpret_cast<float4*>(s_hi + off) = h;
        *reinterpret_cast<float4*>(s_lo + off) = l;
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_nt_tf32x3_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int M, int N,
                      int K) {
    extern __shared__ __align__(1024) float smem[];
    float* a_hi = smem;                          // each 128 x 32 fp32 = 16 KB
    float* a_lo = a_hi + TC_M * TC_KC;
    float* b_hi = a_lo + TC_M * TC_KC;
    float* b_lo = b_hi + TC_N * TC_KC;
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;

    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mbar = smem_u32(&s_mbar);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                                           // TMEM: 128 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const float* Ag = A + (size_t)tile_m * TC_M * K;
    const float* Bg = B + (size_t)tile_n * TC_N * K;
    const uint32_t idesc = make_idesc_tf32(TC_M, TC_N);
    uint32_t parity = 0;
    for (int k0 = 0; k0 < K; k0 += TC_KC) {
        stage_operand(Ag + k0, K, a_hi, a_lo);
        stage_operand(Bg + k0, K, b_hi, b_lo);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy writes -> async proxy (UMMA)
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int s = 0; s < TC_KC / 8; ++s) {                              // UMMA K = 8 tf32 = two core matrices
                const uint32_t koff = (uint32_t)s * 2u * 2048u;
                const uint64_t ah = make_kmajor_desc(smem_u32(a_hi) + koff, 2048, 128);
                const uint64_t al = make_kmajor_desc(smem_u32(a_lo) + koff, 2048, 128);
                const uint64_t bh = make_kmajor_desc(smem_u32(b_hi) + koff, 2048, 128);
                const uint64_t bl = make_kmajor_desc(smem_u32(b_lo) + koff, 2048, 128);
                umma_tf32(tmem, ah, bh, idesc, (k0 | s) != 0);
                umma_tf32(tmem, ah, bl, idesc, true);
                umma_tf32(tmem, al, bh, idesc, true);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
        }
        mbar_wait(mbar, parity);                                               // MMAs done: smem may be overwritten
        parity ^= 1;
        __syncthreads();
    }
    // ---- epilogue: TMEM -> registers -> global ----
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tile_m * TC_M + warp * 32 + lane;
    float* drow = D + (size_t)row * N + (size_t)tile_n * TC_N;
#pragma unroll 1
    for (int c0 = 0; c0 < TC_N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(drow + c0 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int gemm_nt_tf32x3(const float* A, const float* B, float* D, int M, int N, int K, cudaStream_t st) {
    if (!A || !B || !D || M < 1 || N < 1 || K < 1) return HMCX_ERR_INVALID_ARG;
    if ((M % TC_M) || (N % TC_N) || (K % TC_KC)) return HMCX_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(2 * TC_M + 2 * TC_N) * TC_KC * sizeof(float);          // 64 KB
    if (cudaFuncSetAttribute(gemm_nt_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return HMCX_ERR_CUDA;
    }
    dim3 grid(N / TC_N, M / TC_M);
    gemm_nt_tf32x3_kernel<<<grid, TC_THREADS, smem, st>>>(A, B, D, M, N, K);
    return cudaGetLastError() == cudaSuccess ? HMCX_OK : HMCX_ERR_CUDA;
}

}  // namespace hmcx
