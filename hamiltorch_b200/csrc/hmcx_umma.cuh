// hmcx_umma.cuh -- sm_100a tensor-core plumbing shared by the dense-target GEMM (hmcx_tc.cu) and the BNN first-layer
// GEMMs (hmcx_mlp.cu): UMMA shared-memory / instruction descriptors, tcgen05.mma (A from shared memory or from TENSOR
// MEMORY), tcgen05.ld / st, mbarriers, 1-D bulk TMA.  Descriptor encodings follow the CUTLASS sm100 definitions
// (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor) -- re-derived here, no CUTLASS code is used.
#pragma once
#include "hmcx_common.cuh"

namespace hmcx {

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
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, bool a_mn_major = false, bool b_mn_major = false) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}


// D[tmem] (+)= A[tmem] . B[smem]^T : the A operand read from tensor memory (lane = row of A, one 32-bit column per
// tf32 element, 8 columns per instruction) -- how a TMEM-resident epilogue result feeds the next GEMM without smem.
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// ---- warp-uniform issue ------------------------------------------------------------------------------------------
// tcgen05.mma takes its descriptors from UNIFORM registers.  Issued from `if (threadIdx.x == 0)` -- divergent control flow --
// every operand of every instruction goes through an R2UR first, and the issue rate is ~112-120 cycles per MMA whatever
// its size (scripts/bench_cuda/umma_rate.cu: N = 64, 128 and 256 all cost the same).  Issued by a WHOLE warp in warp-uniform
// control flow with only the instruction predicated on the elected lane, the descriptor arithmetic stays in uniform
// registers: 65 cycles per M=128 N=128 K=8 tf32 MMA (its floor is 64), 64 per MMA for the hi|lo N / N/2 pattern.
__device__ __forceinline__ uint32_t elect_one() {            // 1 in exactly one lane of a converged warp
    uint32_t r;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.b32 %0, 1, 0, e;\n\t}\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void umma_tf32_p(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate,
                                            uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_tf32_ta_p(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               bool accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint32_t mbar, uint32_t leader) {     // by the lane that issued the MMAs
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %1, 0;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
                 :: "r"(mbar), "r"(leader) : "memory");
}

// ---- thread-block clusters: TMA multicast + multicast commit --------------------------------------------------------
// One bulk copy lands at the SAME shared-memory offset of every CTA in `cta_mask` and completes `bytes` on the mbarrier at
// the same offset in each of them.  A commit with a mask arrives on that barrier offset in every CTA of the mask.
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast_p(uint32_t mbar, uint16_t cta_mask, uint32_t leader) {
    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}\n"
                 :: "r"(mbar), "h"(cta_mask), "r"(leader) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctaid_y() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.y;" : "=r"(r)); return r; }

// 32 lanes x 16 columns of 32-bit cells: thread (lane) <-> TMEM lane, register i <-> column i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
           "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace hmcx
