// Shared pieces of the tcgen05 ("UMMA") clip kernels (forward: savi_fwd_umma.cu, backward: savi_bwd_umma.cu).
//
// Execution model.  One CTA (or a cluster of 2) owns one clip for the whole recurrence.
//   * 16 compute warps (512 threads): thread (wg, o), wg = warp / 4, o = (warp % 4) * 32 + lane.  For slot-side
//     tensors the thread owns FEATURE o of the 8 slots [8 wg, 8 wg + 8); the fp32 slot state lives in its
//     registers for the whole kernel.  In the token pass it owns TOKEN o of a tile with all K <= 24 slots of
//     that token (one whole tile per warpgroup, see "Token pass" below).
//   * warp 16, one lane: producer.  Streams the static schedule of 16 KB operand blocks (weight images and
//     token tiles, all pre-swizzled in global memory) into a shared-memory ring with 1-D bulk copies.
//   * warp 17, one lane: tcgen05.mma issuer.  Every product is computed TRANSPOSED, Y^T = W . X^T, so the
//     128 output features are the M dimension of the tensor core and the (<= 32) slots are N: accumulator
//     row = TMEM lane = feature = the compute thread that post-processes it (32x32b tcgen05.ld, no shuffles).
//
// fp32 activations are split x = hi + lo (two bf16).  An activation operand is ONE MN-major B matrix
// [128 rows = contraction index][64 = 32 slots hi | 32 slots lo] (SWIZZLE_128B, 128 B per row): the thread that
// owns feature / token `o` writes row `o` with two 16-byte stores.  Weights are single 16-bit images (savi_layout.h:
// fp16 in the forward, where the activation operands are fp16 hi | lo too; bf16 in the backward): per 16-wide k-step a
// product is ONE MMA,  W . [X_hi | X_lo]  (N = 64), and the epilogue adds accumulator columns s and 32 + s.
#pragma once
#ifndef SAVI_FWD_RING8
#define SAVI_FWD_RING8 1          // forward shared-memory plan with 8 ring stages (0: round-1 plan, 6 stages; A/B builds)
#endif
#include <cuda_fp16.h>
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"

namespace uc {
using namespace umma;
typedef __nv_bfloat16 bf16;

constexpr int NCW = 16, NCT = NCW * 32;       // compute warps / threads
constexpr int W_PROD = 16, W_MMA = 17;
constexpr int NTHREADS = 576;
constexpr int BLK = 16384;                    // ring block: [128 rows][64] bf16, SWIZZLE_128B
constexpr int OPB = 16384;                    // activation operand: [128 rows][32 hi | 32 lo]
constexpr int F = 128;                        // feature width handled by this path (D = Ds = M = 128)
constexpr int KH = 8;                         // slots per compute thread
constexpr int KTOK = 24;                      // slot columns a token thread handles in the token pass (K <= 24 on this path)
constexpr uint32_t IDESC_K_MN64 = idesc_bf16(128, 64, false, true);    // A K-major,  B MN-major, N = 64
constexpr uint32_t IDESC_K_MN32 = idesc_bf16(128, 32, false, true);    //                          N = 32 (hi half of B's rows)
constexpr uint32_t IDESC_MN_MN64 = idesc_bf16(128, 64, true, true);    // A MN-major, B MN-major, N = 64
// slot-side linears: ONE 16-bit weight block per panel (savi_layout.h).  The forward kernel defines SAVI_CLIP_F16_LINEAR:
// fp16 weights x fp16 hi | lo activations; the backward multiplies bf16 weights by bf16 hi | lo gradient operands.
#ifdef SAVI_CLIP_F16_LINEAR
constexpr uint32_t IDESC_LIN = idesc_f16(128, 64, false, true);
#else
constexpr uint32_t IDESC_LIN = IDESC_K_MN64;
#endif
constexpr int NBW = WIMG_NB;                  // ring blocks per [128 x 64] weight panel (savi_layout.h)

// predictor attention core (shared-memory, SIMT): row stride of the q / k / v / dO tiles and of the attention matrices
constexpr int MHA_LD = F + 4;
constexpr int MHA_JP = 6;                     // keys per thread in the logits loops: ceil(K / 4) with K <= 24 on this path
__host__ __device__ __forceinline__ constexpr int mha_ka(int K) { return (K + 3) & ~3; }
// bytes the backward core needs from opA on (dO, q, k, v tiles + A, dL, dL^T); the forward needs less
__host__ __device__ __forceinline__ constexpr int mha_bwd_bytes(int K, int H) { return (4 * K * MHA_LD + 3 * H * K * mha_ka(K)) * 4 + 64; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// Fast transcendental forms for the bf16-mode clip kernels (2e-2 parity class; relative error ~2^-22, one MUFU each):
// the 512 compute threads of a CTA share four issue ports, and the IEEE-exact expf / division / tanhf sequences were
// ~40 % of the instructions of the GRU phase, which sits on the serial chain of every step.
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.0f + ex2_fast(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.0f, rcp_fast(1.0f + ex2_fast(-2.8853900817779268f * x)), -1.0f); }

__device__ __forceinline__ void bar_sync_n(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_sync_compute() { bar_sync_n(1, NCT); }

__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#ifdef UMMA_BOUNDED_WAIT
    for (uint32_t n = 0; !mbar_try_wait_cluster(bar, parity); ++n) { if (n > (1u << 24)) { printf("mbar_wait_cluster timeout: block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); } }
#else
    while (!mbar_try_wait_cluster(bar, parity)) { }
#endif
}
// A buffer that other agents refill (peer stores, bulk copies) must only be released once the values read from it have ARRIVED:
// an arrive that merely follows the loads in program order can be scheduled ahead of their data.  Naming a loaded value as the
// input of an (empty) volatile asm pins the release behind the load's completion at no cost.
__device__ __forceinline__ void loaded_before_release(float v) { asm volatile("" :: "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" :: "r"(cluster_addr) : "memory");
}

// ---- shared-memory plan -------------------------------------------------------------------------
struct Smem {
    int ring, nst;                 // ring of nst blocks
    int opA, opB, opC;             // slot-side activation operands
    int aw0, aw1;                  // token pass: attention-weight tiles [128 tokens][32 hi | 32 lo]
    int scratch;                   // fp32 [KR][128] transposition scratch (LayerNorm statistics); forward: follows aw1 (predictor q/k/v alias aw0..scratch), backward: aliases aw1
    int stats;                     // float2 [32]
    int ones;                      // MN-major ones operand [16][128] bf16
    int inbox;                     // CN = 2: [KR][128] (+ [32]) fp32 written by the peer CTA; forward: two buffers, backward: one + handshake
    int inbox_stride;
    int aux;                       // backward: extra region (see savi_bwd_umma.cu)
    int bars;                      // mbarriers + tmem base
    int total;
};
constexpr int NBAR = 64;
constexpr int SMEM_MAX = 227 * 1024;             // the kernels use no static shared memory: the whole opt-in window is dynamic
__host__ __device__ inline Smem plan_smem(int K, int CN, bool bwd) {
    Smem s;
    const int KR = (K + 3) & ~3;
    int p = 0;
    s.opA = p; p += OPB; s.opB = p; p += OPB; s.opC = p; p += OPB;
    s.aw0 = p; p += OPB; s.aw1 = p; p += OPB;
    // backward: its token tiles take ~6 us each, so all four of a CTA must be in flight: that needs 8 ring stages (two per
    // tile).  The fp32 LayerNorm / c-vector scratch therefore lives in aw1 (not an operand buffer in the backward), the
    // float2 LayerNorm-backward scratch in opC | aw0 (idle at every call site: the products that read them have completed),
    // and the cluster inbox is single-buffered behind a "consumed" handshake.  (Measured in the forward: no gain, not done.)
    // Round 2: the forward uses the same plan (8 ring stages = the token blocks of four tiles in flight).  With many tiles
    // per CTA (N >= 4096) the forward token pass was starved by the ring: a tile's two blocks stay resident from their bulk copy
    // until its second product completes (~4.5 us), so 6 stages delivered one tile per ~1.5 us while the compute warpgroups
    // needed ~0.5 us (profiles/r02p_phase_n16k.txt: 70 % of the forward spent in "WAIT logits").
    const bool alias = bwd || SAVI_FWD_RING8;
    if (alias) s.scratch = s.aw1; else { s.scratch = p; p += KR * F * 4; }
    s.stats = p; p += 512;                         // float2 [32] LayerNorm statistics | float [64] c, 1/S (backward)
    p = (p + 1023) & ~1023;
    s.ones = p; p += bwd ? 0 : 4096;
    const int KR8 = (K + 7) & ~7;                 // inbox layout [128 features][KR8 slots] (+ [KR8] sums in the forward)
    s.inbox_stride = (KR8 * F * 4 + (bwd ? 0 : 32 * 4) + 127) & ~127;
    s.inbox = p; p += (CN > 1) ? (alias ? 1 : 2) * s.inbox_stride : 0;
    s.aux = p;
    p = (p + 15) & ~15;
    s.bars = p; p += NBAR * 8 + 16 + 512;          // mbarriers, TMEM base, development counters
    p = (p + 1023) & ~1023;
    s.ring = p;
    s.nst = (SMEM_MAX - p) / BLK;
    if (s.nst > 12) s.nst = 12;
    s.nst &= ~1;                                 // even: a token tile's two blocks never wrap around the ring
    s.total = p + s.nst * BLK;
    return s;
}

// barrier indices inside Smem::bars
enum { B_FULL = 0, B_EMPTY = 12, B_OPND = 24, B_ACC = 25, B_SFREE4 = 26 /* x4: logits accumulator of a warpgroup drained */,
       B_AREADY = 30 /* x2: weight / dL tile written */, B_TOK = 34,
       B_INBOX = 35, B_OPND2 = 37, B_ACC2 = 38,
       B_FACC = 40 /* x4: predictor FFN hidden tiles */, B_FOPND = 44 /* x4: their operand chunks */,
       B_SFULL4 = 48 /* x4: logits of a tile ready, one barrier per compute warpgroup */,
       B_AFREE4 = 52 /* x4: the weight / dL tile buffer this warpgroup writes next has been consumed */, B_END = 56 };
// An mbarrier only distinguishes the parity of its phase: a barrier must never complete two phases before its
// waiter has observed the first.  Strictly alternating producer / consumer pairs share B_OPND / B_ACC; anything
// that is signalled several times in a row (FFN tiles and chunks, token tiles) has its own barrier per item.

// Token pass.  Tiles carry a running sequence number n over the whole kernel.  Tile n is handled by compute warpgroup n & 3,
// which owns logits accumulator n & 3 (32 TMEM columns: the hi and lo halves of the slot-side operand accumulate into the
// same columns) and processes a whole tile (all slots of a token in one thread: the softmax needs no exchange); the
// weight / dL tile it writes for the second product is buffer n & 1, shared with warpgroup (n + 2) & 3.  Barriers that a
// warpgroup WAITS on are private to it (B_SFULL4 / B_AFREE4 + wg): an mbarrier only distinguishes the parity of its phase,
// and a warpgroup would otherwise wait two phases ahead of a barrier shared with its partner.
// first products issued ahead of the second ones: the ring must hold the token blocks (2 per tile) of LA + 1 tiles
__host__ __device__ __forceinline__ constexpr int tok_lookahead(int nst) { return nst >= 8 ? 3 : nst >= 6 ? 2 : 1; }

struct Ring {
    unsigned char* base; uint64_t* full; uint64_t* empty; int nst; int stage; uint32_t phase;
    long long* wait_cycles;      // development: cycles the issuer spent waiting for ring blocks (nullptr = off)
    __device__ __forceinline__ void advance() { if (++stage == nst) { stage = 0; phase ^= 1u; } }
};

// ---- producer ----------------------------------------------------------------------------------
__device__ __forceinline__ void prod_blocks(Ring& r, const unsigned char* src, int nblk) {
#pragma unroll 1                 // one thread, always ahead of its consumers: keep its code small (the SM's instruction cache is shared)
    for (int i = 0; i < nblk; ++i) {
        mbar_wait(&r.empty[r.stage], r.phase ^ 1u);
        mbar_expect_tx(&r.full[r.stage], BLK);
        bulk_g2s(r.base + (size_t)r.stage * BLK, src + (size_t)i * BLK, BLK, &r.full[r.stage]);
        r.advance();
    }
}

// Token blocks: the same, plus an L2 prefetch `ahead` blocks in front of the shared-memory copy.  A block stays in the ring from its
// bulk copy until the tile's second product completes, so with many tiles per CTA (N >= 4096) the ring turns over once per
// HBM round trip; pulling the blocks into L2 a few tiles early makes that round trip an L2 hit.
__device__ __forceinline__ void prod_token_blocks(Ring& r, const unsigned char* src, int nblk, int ahead) {
    for (int i = 0; i < ahead && i < nblk; ++i) prefetch_l2(src + (size_t)i * BLK, BLK);
#pragma unroll 1
    for (int i = 0; i < nblk; ++i) {
        if (i + ahead < nblk) prefetch_l2(src + (size_t)(i + ahead) * BLK, BLK);
        mbar_wait(&r.empty[r.stage], r.phase ^ 1u);
        mbar_expect_tx(&r.full[r.stage], BLK);
        bulk_g2s(r.base + (size_t)r.stage * BLK, src + (size_t)i * BLK, BLK, &r.full[r.stage]);
        r.advance();
    }
}

// ---- issuer ------------------------------------------------------------------------------------
// One linear layer, transposed: acc[rt] (+)= Wimg[rt] . X^T.  Weight blocks arrive through the ring in image
// order (rt, cb, hi, lo).  xop: shared address of the activation operand; its rows [cb*64, cb*64 + 64) are the
// contraction range of block cb.  Accumulators are 64 columns wide (32 slots x {hi, lo} of X).
// The WHOLE issuer warp runs this code with warp-uniform values and only the tcgen05 instructions are predicated
// on one elected lane `el`: the descriptors then live in uniform registers.  (Issued from single-lane divergent
// code, every tcgen05.mma is wrapped in a register-to-uniform "waterfall" loop of ~100 cycles.)
// The body is ONE real function (the kernels call it from 15-25 sites; inlined copies were ~15 % of the SASS, and the
// clip kernels are sensitive to instruction-cache misses).  All state travels by value in registers: ring position in,
// ring position out (a reference to the Ring would force it into local memory).
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mma_commit_a(uint32_t bar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar_addr) : "memory");
}
static __device__ __noinline__ uint32_t issue_linear_core(uint32_t rs /* stage | phase << 8 */, uint32_t el, uint32_t xop, uint32_t tacc,
                                                          int ntile, int ncb, uint32_t accumulate,
                                                          uint32_t ring_base, uint32_t full0, uint32_t empty0, int nst) {
    uint32_t stage = rs & 0xffu, phase = rs >> 8;
    for (int rt = 0; rt < ntile; ++rt) {
        const uint32_t d = tacc + rt * 64;
        for (int cb = 0; cb < ncb; ++cb) {
            const uint32_t xb = dlo_mn(xop + cb * 8192, BLK);              // 64 rows x 128 B of the activation operand
            while (!mbar_try_wait_a(full0 + stage * 8u, phase)) { }
            fence_after_sync();
            uint32_t a = dlo_k(ring_base + stage * BLK);
            if (el) {                                                     // W . [X_hi | X_lo]: one 16-bit weight block per panel
                mma_lo(d, a, xb, IDESC_LIN, (accumulate || cb > 0) ? 1u : 0u);
#pragma unroll
                for (int k4 = 1; k4 < 4; ++k4) mma_lo(d, a + k4 * 2, xb + k4 * 128, IDESC_LIN, 1u);
                mma_commit_a(empty0 + stage * 8u);
            }
            __syncwarp();
            if (++stage == (uint32_t)nst) { stage = 0; phase ^= 1u; }
        }
    }
    return stage | (phase << 8);
}
__device__ __forceinline__ void issue_linear(Ring& r, bool el, uint32_t xop, uint32_t tacc, int ntile, int ncb, bool accumulate) {
    const uint32_t rs = issue_linear_core((uint32_t)r.stage | (r.phase << 8), el ? 1u : 0u, xop, tacc, ntile, ncb, accumulate ? 1u : 0u,
                                          smem_u32(r.base), smem_u32(r.full), smem_u32(r.empty), r.nst);
    r.stage = (int)(rs & 0xffu); r.phase = rs >> 8;
}

// ---- compute-thread helpers ----------------------------------------------------------------------
struct Ctx {
    int tid, warp, lane, wg, o, K;
    int nk;                      // valid slots of this thread: clamp(K - 8 wg, 0, 8)   (warp-uniform)
    int k0;                      // first slot: 8 wg
    unsigned char* sm;
    Smem L;
    uint32_t tb;                 // TMEM base
    uint32_t tlane;              // TMEM lane field of this warp's quadrant
    uint32_t row_hi, row_lo;     // byte offsets of this thread's hi / lo 16-byte chunk inside an activation operand
    uint64_t* bars;
    uint32_t ph_acc;             // running phase of B_ACC (compute side)
};

__device__ __forceinline__ void ctx_init(Ctx& c, int tid, int K, unsigned char* sm, const Smem& L, uint32_t tb, uint64_t* bars) {
    c.tid = tid; c.warp = __shfl_sync(0xffffffffu, tid >> 5, 0); c.lane = tid & 31; c.wg = c.warp >> 2; c.o = (c.warp & 3) * 32 + c.lane; c.K = K;
    c.k0 = c.wg * KH;
    c.nk = min(max(K - c.k0, 0), KH);
    c.sm = sm; c.L = L; c.tb = tb; c.bars = bars; c.ph_acc = 0;
    c.tlane = (uint32_t)((c.warp & 3) * 32) << 16;
    c.row_hi = (uint32_t)c.o * 128u + ((uint32_t)(c.wg ^ (c.o & 7)) << 4);
    c.row_lo = (uint32_t)c.o * 128u + ((uint32_t)((4 + c.wg) ^ (c.o & 7)) << 4);
}

// X^T operand row of this thread: its 8 slot values as bf16 hi (chunk wg) and lo (chunk 4 + wg); slots >= K are zero
__device__ __forceinline__ void write_operand(const Ctx& c, int op_off, const float (&v)[KH]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = (2 * j < c.nk) ? v[2 * j] : 0.f, b = (2 * j + 1 < c.nk) ? v[2 * j + 1] : 0.f;
        const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
        const float2 hf = __bfloat1622float2(hh);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = pack_bf16x2(a - hf.x, b - hf.y);
    }
    *reinterpret_cast<uint4*>(c.sm + op_off + c.row_hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(c.sm + op_off + c.row_lo) = make_uint4(l[0], l[1], l[2], l[3]);
}
// Two [K][128] tensors a, b that multiply the SAME A operand, packed by precision half: op_hi row = [a_hi 32 slots | b_hi 32 slots],
// op_lo row = [a_lo | b_lo] (bf16).  One N = 64 MMA per half then yields [A . a | A . b] in one 64-column accumulator.
__device__ __forceinline__ void write_operand_halves(const Ctx& c, int op_hi, int op_lo, const float (&a)[KH], const float (&b)[KH]) {
    uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a0 = (2 * j < c.nk) ? a[2 * j] : 0.f, a1 = (2 * j + 1 < c.nk) ? a[2 * j + 1] : 0.f;
        const float b0 = (2 * j < c.nk) ? b[2 * j] : 0.f, b1 = (2 * j + 1 < c.nk) ? b[2 * j + 1] : 0.f;
        const __nv_bfloat162 ha = __floats2bfloat162_rn(a0, a1), hb = __floats2bfloat162_rn(b0, b1);
        const float2 fa = __bfloat1622float2(ha), fb = __bfloat1622float2(hb);
        ah[j] = *reinterpret_cast<const uint32_t*>(&ha); bh[j] = *reinterpret_cast<const uint32_t*>(&hb);
        al[j] = pack_bf16x2(a0 - fa.x, a1 - fa.y); bl[j] = pack_bf16x2(b0 - fb.x, b1 - fb.y);
    }
    // row_hi = this thread's chunk in the first 32 columns, row_lo = the same chunk in the second 32
    *reinterpret_cast<uint4*>(c.sm + op_hi + c.row_hi) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
    *reinterpret_cast<uint4*>(c.sm + op_hi + c.row_lo) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
    *reinterpret_cast<uint4*>(c.sm + op_lo + c.row_hi) = make_uint4(al[0], al[1], al[2], al[3]);
    *reinterpret_cast<uint4*>(c.sm + op_lo + c.row_lo) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
}
// same as fp16 hi | lo (operands of the forward's slot-side linears, whose weight images are fp16).  The conversion
// saturates: a slot state beyond fp16's range (65504) would otherwise turn into inf - inf = NaN in the lo half.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void write_operand_f16(const Ctx& c, int op_off, const float (&v)[KH]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = (2 * j < c.nk) ? v[2 * j] : 0.f, b = (2 * j + 1 < c.nk) ? v[2 * j + 1] : 0.f;
        h[j] = pack_f16x2_sat(a, b);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
        l[j] = pack_f16x2_sat(a - hf.x, b - hf.y);
    }
    *reinterpret_cast<uint4*>(c.sm + op_off + c.row_hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(c.sm + op_off + c.row_lo) = make_uint4(l[0], l[1], l[2], l[3]);
}
// operand complete: make the generic-proxy writes visible to the tensor core, one arrival per warp
__device__ __forceinline__ void signal_operand(const Ctx& c, int bar = B_OPND) {
    fence_async_smem();
    fence_before_sync();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.bars[bar]);
}
__device__ __forceinline__ void wait_acc(Ctx& c) {
    mbar_wait(&c.bars[B_ACC], c.ph_acc);
    c.ph_acc ^= 1u;
    fence_after_sync();
}
// this thread's 8 slot columns of a 64-column accumulator: (W_hi X_hi + W_lo X_hi) + W_hi X_lo
__device__ __forceinline__ void load_acc(const Ctx& c, int col, float (&v)[KH]) {
    float lo[KH];
    tmem_ld8(c.tb + c.tlane + col + c.k0, v);
    tmem_ld8(c.tb + c.tlane + col + 32 + c.k0, lo);
    tmem_wait_ld();
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) v[kk] += lo[kk];
}
// this thread's column `col` of a [K][ld] fp32 field-array row block
__device__ __forceinline__ void save_field(const Ctx& c, float* dst, int ld, int col, const float (&v)[KH]) {
    float* p = dst + (size_t)c.k0 * ld + col;
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) p[kk * ld] = v[kk];
}
__device__ __forceinline__ void load_field(const Ctx& c, const float* src, int ld, int col, float (&v)[KH]) {
    const float* p = src + (size_t)c.k0 * ld + col;
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) v[kk] = (kk < c.nk) ? p[kk * ld] : 0.f;
}

// The barrier-to-barrier middle of the LayerNorm helpers takes no register arrays: ONE real function each instead of an
// inlined copy per call site (instruction-cache diet, see issue_linear_core).
static __device__ __noinline__ void slot_stats_reduce(const float* scr, float2* st, int warp, int lane, int K, float eps) {
    bar_sync_compute();
#pragma unroll 1
    for (int k = warp; k < K; k += NCW) {
        const float4 x = ld4(scr + k * F + lane * 4);
        const float mean = warp_sum((x.x + x.y) + (x.z + x.w)) * (1.0f / F);
        const float a = x.x - mean, b = x.y - mean, e = x.z - mean, f = x.w - mean;
        const float var = warp_sum((a * a + b * b) + (e * e + f * f)) * (1.0f / F);
        if (lane == 0) st[k] = make_float2(mean, 1.0f / sqrtf(var + eps));
    }
    bar_sync_compute();
}
// per-slot mean / rstd over the 128 features (torch LayerNorm: biased variance, two-pass), via the scratch tile
__device__ __forceinline__ void slot_stats(const Ctx& c, const float (&v)[KH], float eps) {
    float* scr = reinterpret_cast<float*>(c.sm + c.L.scratch);
    float2* st = reinterpret_cast<float2*>(c.sm + c.L.stats);
    float* p = scr + c.k0 * F + c.o;
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) p[kk * F] = v[kk];
    slot_stats_reduce(scr, st, c.warp, c.lane, c.K, eps);
}
// stats_out (nullable): global [K][2] record of (mean, rstd), written by one thread per warpgroup
__device__ __forceinline__ void layer_norm(const Ctx& c, const float (&v)[KH], float (&y)[KH], float gamma, float beta, float eps,
                                           float2* stats_out = nullptr) {
    slot_stats(c, v, eps);
    const float2* st = reinterpret_cast<const float2*>(c.sm + c.L.stats) + c.k0;
    if (stats_out && c.o == 0) {
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) stats_out[c.k0 + kk] = st[kk];
    }
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        if (kk < c.nk) { const float2 s = st[kk]; y[kk] = (v[kk] - s.x) * s.y * gamma + beta; }
        else y[kk] = 0.f;
    }
}

}  // namespace uc
