// Shared pieces of the tcgen05 ("UMMA") clip kernels (forward: savi_fwd_umma.cu, backward: savi_bwd_umma.cu).
//
// Execution model.  One CTA (or a cluster of 2) owns one clip for the whole recurrence.
//   * 8 compute warps (256 threads): thread (g, o), g = warp / 4, o = (warp % 4) * 32 + lane.  For slot-side
//     tensors the thread owns FEATURE o of the slots [g*KH, g*KH + KH); the fp32 slot state lives in its
//     registers for the whole kernel.  For the token pass it owns TOKEN o of the tiles with tile % 2 == g.
//   * warp 8, one lane: producer.  Streams the static schedule of 16 KB operand blocks (weight images and
//     token tiles, all pre-swizzled in global memory) into a shared-memory ring with 1-D bulk copies.
//   * warp 9, one lane: tcgen05.mma issuer.  Every product is computed TRANSPOSED, Y^T = W . X^T, so the
//     128 output features are the M dimension of the tensor core and the (<= 32) slots are N: accumulator
//     row = TMEM lane = feature = the compute thread that post-processes it (32x32b tcgen05.ld, no shuffles).
// fp32 operands are split x = hi + lo (two bf16) and multiplied with three MMAs (hi.hi + hi.lo + lo.hi).
#pragma once
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"

namespace uc {
using namespace umma;
typedef __nv_bfloat16 bf16;

constexpr int NCW = 8, NCT = NCW * 32;        // compute warps / threads
constexpr int W_PROD = 8, W_MMA = 9;
constexpr int NTHREADS = 320;
constexpr int BLK = 16384;                    // ring block: [128 rows][64] bf16, SWIZZLE_128B
constexpr int OPB = 16384;                    // operand buffer: hi [2 blocks of 32 x 64] 8 KB | lo 8 KB
constexpr int OP_LO = 8192, OP_CB = 4096;     // byte offsets: lo half, second 64-column block
constexpr int NS = 32;                        // MMA N (slots, zero-padded)
constexpr int F = 128;                        // feature width handled by this path (D = Ds = M = 128)
constexpr int KHMAX = 16;                     // slots per thread (register arrays)
constexpr uint32_t IDESC_KK = idesc_bf16(128, NS, false, false);    // A K-major,  B K-major
constexpr uint32_t IDESC_MK = idesc_bf16(128, NS, true, false);     // A MN-major, B K-major

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void bar_sync_compute(int id) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(NCT) : "memory"); }

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
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" :: "r"(cluster_addr) : "memory");
}

// ---- shared-memory plan -------------------------------------------------------------------------
struct Smem {
    int ring, nst;                 // ring of nst blocks
    int opA, opB, opC;             // slot-side B operands
    int aw0, aw1;                  // token pass: attention-weight tiles [32 slots][128 tokens] hi | lo
    int scratch;                   // fp32 [KR][128] transposition scratch (LayerNorm statistics); follows aw1 (predictor q/k/v alias aw0..scratch)
    int stats;                     // float2 [32]
    int ones;                      // MN-major ones operand [16][128] bf16
    int inbox;                     // CN = 2: two buffers of [KR][128] + [32] fp32 written by the peer CTA
    int inbox_stride;
    int aux;                       // backward: extra region (see savi_bwd_umma.cu)
    int bars;                      // mbarriers + tmem base
    int total;
};
constexpr int NBAR = 64;
__host__ __device__ inline Smem plan_smem(int K, int CN, int aux_bytes, int max_bytes) {
    Smem s;
    const int KR = (K + 3) & ~3;
    int p = 0;
    s.opA = p; p += OPB; s.opB = p; p += OPB; s.opC = p; p += OPB;
    s.aw0 = p; p += OPB; s.aw1 = p; p += OPB;
    s.scratch = p; p += KR * F * 4;
    s.stats = p; p += 32 * 8;
    p = (p + 1023) & ~1023;
    s.ones = p; p += 4096;
    s.inbox_stride = (KR * F * 4 + 32 * 4 + 127) & ~127;
    s.inbox = p; p += (CN > 1) ? 2 * s.inbox_stride : 0;
    s.aux = p; p += aux_bytes;
    p = (p + 15) & ~15;
    s.bars = p; p += NBAR * 8 + 16;
    p = (p + 1023) & ~1023;
    s.ring = p;
    s.nst = (max_bytes - 1024 - p) / BLK;        // 1024: base alignment slack
    if (s.nst > 12) s.nst = 12;
    s.nst &= ~1;                                 // even: a token tile's two blocks never wrap around the ring
    s.total = p + s.nst * BLK + 1024;
    return s;
}

// barrier indices inside Smem::bars
enum { B_FULL = 0, B_EMPTY = 12, B_OPND = 24, B_ACC = 25, B_SFULL = 26, B_SFREE = 28, B_AREADY = 30, B_AFREE = 32, B_TOK = 34,
       B_INBOX = 35, B_OPND2 = 37, B_ACC2 = 38,
       B_FACC = 40 /* x4: predictor FFN hidden tiles */, B_FOPND = 44 /* x4: their operand chunks */, B_END = 48 };
// An mbarrier only distinguishes the parity of its phase: a barrier must never complete two phases before its
// waiter has observed the first.  Strictly alternating producer / consumer pairs share B_OPND / B_ACC; anything
// that is signalled several times in a row (FFN tiles and chunks, token tiles) has its own barrier per item.

struct Ring {
    unsigned char* base; uint64_t* full; uint64_t* empty; int nst; int stage; uint32_t phase;
    __device__ __forceinline__ void advance() { if (++stage == nst) { stage = 0; phase ^= 1u; } }
};

// ---- producer ----------------------------------------------------------------------------------
__device__ __forceinline__ void prod_blocks(Ring& r, const unsigned char* src, int nblk) {
    for (int i = 0; i < nblk; ++i) {
        mbar_wait(&r.empty[r.stage], r.phase ^ 1u);
        mbar_expect_tx(&r.full[r.stage], BLK);
        bulk_g2s(r.base + (size_t)r.stage * BLK, src + (size_t)i * BLK, BLK, &r.full[r.stage]);
        r.advance();
    }
}

// ---- issuer ------------------------------------------------------------------------------------
// One linear layer, transposed: acc[rt] (+)= Wimg[rt] . X^T.  Blocks arrive through the ring in image order
// (rt, cb, hi/lo).  xop: shared address of the operand buffer holding X hi | lo for columns [cb0*64, ...).
__device__ __forceinline__ void issue_linear(Ring& r, uint32_t xop, uint32_t tacc, int col_stride, int ntile, int ncb, bool accumulate) {
    for (int rt = 0; rt < ntile; ++rt) {
        const uint32_t d = tacc + rt * col_stride;
        for (int cb = 0; cb < ncb; ++cb) {
            const uint32_t xh = xop + cb * OP_CB, xl = xh + OP_LO;
            mbar_wait(&r.full[r.stage], r.phase);
            fence_after_sync();
            uint32_t a = smem_u32(r.base + (size_t)r.stage * BLK);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                mma_ss(d, desc_kmajor(a + k4 * 32), desc_kmajor(xh + k4 * 32), IDESC_KK, accumulate || cb > 0 || k4 > 0);
                mma_ss(d, desc_kmajor(a + k4 * 32), desc_kmajor(xl + k4 * 32), IDESC_KK, true);
            }
            mma_commit(&r.empty[r.stage]);
            r.advance();
            mbar_wait(&r.full[r.stage], r.phase);
            fence_after_sync();
            a = smem_u32(r.base + (size_t)r.stage * BLK);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) mma_ss(d, desc_kmajor(a + k4 * 32), desc_kmajor(xh + k4 * 32), IDESC_KK, true);
            mma_commit(&r.empty[r.stage]);
            r.advance();
        }
    }
}

// ---- compute-thread helpers ----------------------------------------------------------------------
struct Ctx {
    int tid, warp, lane, g, o, KH, K;
    unsigned char* sm;
    Smem L;
    uint32_t tb;                 // TMEM base
    uint64_t* bars;
    uint32_t ph_acc, ph_opnd;    // running phases of B_ACC (compute side) / B_OPND (issuer side)
};

// X^T operand: this thread's feature column o of its slots, split into bf16 hi / lo
__device__ __forceinline__ void write_operand(const Ctx& c, int op_off, const float (&v)[KHMAX]) {
    unsigned char* base = c.sm + op_off + (c.o >> 6) * OP_CB;
    const int oc = c.o & 63;
#pragma unroll
    for (int kk = 0; kk < KHMAX; ++kk) {
        const int k = c.g * c.KH + kk;
        if (kk < c.KH && k < c.K) {
            const bf16 hi = __float2bfloat16_rn(v[kk]);
            const bf16 lo = __float2bfloat16_rn(v[kk] - __bfloat162float(hi));
            const uint32_t off = sw128_off(k, oc);
            *reinterpret_cast<bf16*>(base + off) = hi;
            *reinterpret_cast<bf16*>(base + OP_LO + off) = lo;
        }
    }
}
// operand complete: make the generic-proxy writes visible to the tensor core, one arrival per warp
__device__ __forceinline__ void signal_operand(const Ctx& c, int bar = B_OPND) {
    fence_async_smem();
    fence_before_sync();
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.bars[bar]);
}
// wait for the next accumulator commit, then load this thread's KH slot columns of accumulator `col`
__device__ __forceinline__ void wait_acc(Ctx& c) {
    mbar_wait(&c.bars[B_ACC], c.ph_acc);
    c.ph_acc ^= 1u;
    fence_after_sync();
}
__device__ __forceinline__ void load_acc(const Ctx& c, int col, float (&v)[KHMAX]) {
    tmem_ld16(tmem_addr(c.tb, c.warp, col + c.g * c.KH), v);
}
// store this thread's column of a [K][ld] fp32 field array row block
__device__ __forceinline__ void save_field(const Ctx& c, float* dst, int ld, int col, const float (&v)[KHMAX]) {
#pragma unroll
    for (int kk = 0; kk < KHMAX; ++kk) {
        const int k = c.g * c.KH + kk;
        if (kk < c.KH && k < c.K) dst[(size_t)k * ld + col] = v[kk];
    }
}
__device__ __forceinline__ void load_field(const Ctx& c, const float* src, int ld, int col, float (&v)[KHMAX]) {
#pragma unroll
    for (int kk = 0; kk < KHMAX; ++kk) {
        const int k = c.g * c.KH + kk;
        v[kk] = (kk < c.KH && k < c.K) ? src[(size_t)k * ld + col] : 0.f;
    }
}

// per-slot mean / rstd over the 128 features (torch LayerNorm: biased variance, two-pass), via the scratch tile
__device__ __forceinline__ void slot_stats(const Ctx& c, const float (&v)[KHMAX], float eps) {
    float* scr = reinterpret_cast<float*>(c.sm + c.L.scratch);
    float2* st = reinterpret_cast<float2*>(c.sm + c.L.stats);
#pragma unroll
    for (int kk = 0; kk < KHMAX; ++kk) {
        const int k = c.g * c.KH + kk;
        if (kk < c.KH && k < c.K) scr[k * F + c.o] = v[kk];
    }
    bar_sync_compute(1);
    for (int k = c.warp; k < c.K; k += NCW) {
        const float4 x = ld4(scr + k * F + c.lane * 4);
        const float mean = warp_sum((x.x + x.y) + (x.z + x.w)) * (1.0f / F);
        const float a = x.x - mean, b = x.y - mean, e = x.z - mean, f = x.w - mean;
        const float var = warp_sum((a * a + b * b) + (e * e + f * f)) * (1.0f / F);
        if (c.lane == 0) st[k] = make_float2(mean, 1.0f / sqrtf(var + eps));
    }
    bar_sync_compute(1);
}
__device__ __forceinline__ void layer_norm(const Ctx& c, const float (&v)[KHMAX], float (&y)[KHMAX], float gamma, float beta, float eps) {
    slot_stats(c, v, eps);
    const float2* st = reinterpret_cast<const float2*>(c.sm + c.L.stats);
#pragma unroll
    for (int kk = 0; kk < KHMAX; ++kk) {
        const int k = c.g * c.KH + kk;
        if (kk < c.KH && k < c.K) { const float2 s = st[k]; y[kk] = (v[kk] - s.x) * s.y * gamma + beta; }
        else y[kk] = 0.f;
    }
}

}  // namespace uc
