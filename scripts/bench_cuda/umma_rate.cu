// Micro-benchmark: cycles per tcgen05.mma for the operand forms the library uses (developer tool, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hamiltorch_b200/csrc scripts/bench_cuda/umma_rate.cu -o scripts/bench_cuda/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "hmcx_umma.cuh"
using namespace hmcx;

__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {     // D fp32, A/B bf16 (format 1), K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

// predicated form: the WHOLE warp runs the issue loop (warp-uniform control flow, descriptors stay in uniform registers), only
// the elected lane's instruction takes effect
__device__ __forceinline__ void umma_tf32_ta_pred(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, bool acc, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_tf32_pred(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc), "r"(leader) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t r;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.b32 %0, 1, 0, e;\n\t}\n" : "=r"(r));
    return r;
}

// mode 0: tf32, A smem; 1: tf32, A tmem; 2: bf16 A smem; 3: tf32 alternating N / N/2 (the library's hi|lo pattern), A tmem
__global__ void __launch_bounds__(128) k(int mode, int N, int reps, long long* out) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 16384; i += 128) sm[i] = 0.001f * (i & 63);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (mode < 4 && threadIdx.x == 0) {
        const uint32_t d = tmem, a_t = tmem + 256;
        // operands: A 128 x K (K-major core matrices), B N x K
        const uint32_t A_LBO = (128 / 8) * 128, B_LBO = (uint32_t)(N / 8) * 128;
        const uint64_t ad = make_kmajor_desc(smem_u32(sm), A_LBO, 128);
        const uint64_t bd = make_kmajor_desc(smem_u32(sm + 8192), B_LBO, 128);
        const uint32_t id_t = make_idesc_tf32(128, N), id_t2 = make_idesc_tf32(128, N / 2), id_b = make_idesc_bf16(128, N);
        uint32_t par = 0;
        for (int warm = 0; warm < 2; ++warm) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (mode == 0) umma_tf32(d, ad, bd, id_t, r != 0);
                else if (mode == 1) umma_tf32_ta(d, a_t, bd, id_t, r != 0);
                else if (mode == 2) umma_f16(d, ad, bd, id_b, r != 0);
                else { umma_tf32_ta(d, a_t, bd, id_t, r != 0); umma_tf32_ta(d, a_t + 64, bd, id_t2, true); }
            }
            const long long t1 = clock64();
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), par); par ^= 1;
            const long long t2 = clock64();
            if (warm == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    // modes 4-6: the same as 0 / 1 / 3 issued by the whole warp with a predicated instruction
    if (mode >= 4 && threadIdx.x < 32) {
        const uint32_t leader = elect_one();
        const uint32_t d = tmem, a_t = tmem + 256;
        const uint32_t A_LBO = (128 / 8) * 128, B_LBO = (uint32_t)(N / 8) * 128;
        const uint64_t ad = make_kmajor_desc(smem_u32(sm), A_LBO, 128);
        uint64_t bd = make_kmajor_desc(smem_u32(sm + 8192), B_LBO, 128);
        const uint32_t id_t = make_idesc_tf32(128, N), id_t2 = make_idesc_tf32(128, N / 2);
        uint32_t par = 0;
        for (int warm = 0; warm < 2; ++warm) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (mode == 4) umma_tf32_pred(d, ad, bd, id_t, r != 0, leader);
                else if (mode == 5) umma_tf32_ta_pred(d, a_t, bd, id_t, r != 0, leader);
                else { umma_tf32_ta_pred(d, a_t, bd, id_t, r != 0, leader); umma_tf32_ta_pred(d, a_t + 64, bd, id_t2, true, leader); }
                bd += (r & 1) ? -2 : 2;                      // a loop-variant descriptor, as in a real k-loop
            }
            const long long t1 = clock64();
            if (leader) umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), par); par ^= 1;
            const long long t2 = clock64();
            if (warm == 1 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4);
    const char* names[] = {"tf32 A=smem", "tf32 A=tmem", "bf16 A=smem", "tf32 A=tmem, N then N/2 (hi|lo pattern)",
                           "WARP-UNIFORM issue: tf32 A=smem", "WARP-UNIFORM issue: tf32 A=tmem", "WARP-UNIFORM issue: tf32 A=tmem, N then N/2"};
    for (int mode = 0; mode < 7; ++mode)
        for (int N : {64, 128, 256})
            for (int reps : {8, 64}) {
                k<<<1, 128, 16384 * 4>>>(mode, N, reps, d);
                long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                const int per = (mode == 3 || mode == 6) ? 2 * reps : reps;
                printf("%-42s M=128 N=%3d  %3d MMAs: issue %6lld cyc (%5.1f/MMA), issue+complete %6lld cyc (%5.1f/MMA)  floor %d  %s\n",
                       names[mode], N, per, h[0], (double)h[0] / per, h[1], (double)h[1] / per, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
