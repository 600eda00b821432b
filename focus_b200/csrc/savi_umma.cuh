// sm_100a primitives of the tcgen05 path: UMMA shared-memory / instruction descriptors, TMEM
// allocation and loads, mbarriers, 1-D bulk copies (TMA engine, no tensor map), proxy fences.
//
// Operand images in shared memory are the canonical SWIZZLE_128B layouts of the 5th-gen tensor
// core ("UMMA"), for 16-bit elements:
//   block  = [rows][64 elements] : row r at byte r*128, its 16-byte chunk c stored at chunk c ^ (r & 7);
//            8-row groups are 1024 B apart (SBO), the block base is 1024-byte aligned.
//   K-major operand  (rows = M or N index, the 64 elements run along K): one MMA (K = 16) reads 32 B of
//            every row; k-step j of a block starts at byte j*32.
//   MN-major operand (rows = K index, the 64 elements run along M/N): one MMA (K = 16) reads 16 rows;
//            64-element column blocks are LBO bytes apart; k-step j starts at byte j*2048.
// The same [rows][64] image therefore serves as K-major A (rows = M) and as MN-major A (rows = K):
// the token tiles are used both ways (logits and update products) without a transposed copy.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

constexpr uint32_t BLK_BYTES_PER_ROW = 128;       // one swizzle-atom row: 64 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, c) inside a [rows][64] SWIZZLE_128B block (c < 64)
__host__ __device__ __forceinline__ uint32_t sw128_off(int r, int c) {
    return (uint32_t)r * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)c & 7u) << 1);
}

// ---- descriptors -----------------------------------------------------------------------------
// shared-memory matrix descriptor (PTX "tcgen05 matrix descriptor"): start address, leading / stride byte
// offsets (16-byte units), version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand block: 8-row groups 1024 B apart; LBO unused for swizzled K-major layouts
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return smem_desc(saddr, 16, 1024); }
// MN-major operand: column blocks (64 elements of M/N) `blk_stride` bytes apart, 8-row K groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t blk_stride) { return smem_desc(saddr, blk_stride, 1024); }

// instruction descriptor, kind::f16: bf16 x bf16 -> fp32, dense
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with fp16 x fp16 operands (format 0).  A and B must share the format: a mixed fp16 x bf16 kind::f16 MMA is an
// illegal instruction on sm_100a (measured, round 2)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]   (issued by ONE thread on behalf of the CTA)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Issue-efficient form for the hot loops: the single issuing thread is latency-bound on its own instruction
// stream, so the descriptors are kept as (lo, hi) register pairs; hi is a compile-time constant per layout and
// stepping along K is one 32-bit add on lo (start address field, 16-byte units).
constexpr uint32_t DHI = (1024u >> 4) | (1u << 14) | (2u << 29);       // SBO = 1024 B, version 1, SWIZZLE_128B
// (inside a cluster the shared::cta window address carries the CTA's rank in its upper bits: keep the 18-bit mask)
__device__ __forceinline__ constexpr uint32_t dlo_k(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((16u >> 4) << 16); }
__device__ __forceinline__ constexpr uint32_t dlo_mn(uint32_t saddr, uint32_t blk_stride) { return ((saddr & 0x3FFFFu) >> 4) | ((blk_stride >> 4) << 16); }
__device__ __forceinline__ void mma_lo(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"(DHI) : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}

// one lane of a fully active warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// ---- TMEM ------------------------------------------------------------------------------------
// one full warp; ncols a power of two >= 32; the base address lands in *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// 32 lanes x 32 bit, N consecutive columns: thread i of warp w reads lane 32*(w%4)+i (the warp's own quadrant)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// TMEM address of (lane quadrant of this warp, column)
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int warp, int col) { return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col; }

// ---- mbarriers -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// UMMA_BOUNDED_WAIT (development builds): trap instead of spinning forever on a barrier that never completes
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef UMMA_BOUNDED_WAIT
    for (uint32_t n = 0; !mbar_try_wait(bar, parity); ++n) { if (n > (1u << 24)) { printf("mbar_wait timeout: block %d thread %d bar %p parity %u\n", blockIdx.x, threadIdx.x, (void*)bar, parity); __trap(); } }
#else
    while (!mbar_try_wait(bar, parity)) { }
#endif
}

// ---- bulk copies and proxy fences -------------------------------------------------------------
// global -> shared, `bytes` % 16 == 0, both addresses 16-byte aligned; completes on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// same, delivered to the same offset of every CTA in `mask` of the cluster (each CTA's own barrier gets the bytes)
__device__ __forceinline__ void bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// bring `bytes` (multiple of 16, 16-byte aligned) of global memory into L2; fire and forget
__device__ __forceinline__ void prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" :: "l"(src), "r"(bytes) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- cluster ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
// shared::cluster address of `p` (a shared::cta address of this CTA) in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f1(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;\n" :: "r"(addr), "f"(v) : "memory");
}

}  // namespace umma
