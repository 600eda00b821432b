// tcgen05 forward clip kernel: the whole T x I recurrence (+ predictor) of one clip per CTA / CTA pair.
//
// Reference semantics: /root/reference/slowfast/models/STEVE/steve.py:52-105 (SlotAttentionVideo.forward),
// transformer.py:22-49, 70-86, 106-114 (predictor).  Execution model: savi_umma_clip.cuh.
// Writes exactly the saved-for-backward records of the mma.sync kernel (savi_layout.h: SavedLayout).
#define SAVI_CLIP_F16_LINEAR 1      // slot-side linears: fp16 weight images x fp16 hi | lo activations (savi_layout.h)
#include "savi_umma_clip.cuh"

using namespace uc;

// development aid: cycle counters of compute thread 0 of CTA 0 (tools/phase_times.py); no barriers added
#define UPH(id) do { if (dbg) { const long long t_ = clock64(); dbg[id] += t_ - ph_last; ph_last = t_; } } while (0)

// TMEM columns: fp32 accumulators, 64 columns each (32 slots x {X_hi, X_lo} partial products).  IN shares A's columns (A holds
// NUMX during the token pass and the mlp.0 product after the GRU gates have been read: never live together with IN), which
// leaves room for THREE 64-column logits buffers: one tcgen05.mma costs 45 / 51 cycles at N = 32 / 64 (tools/umma_rate.cu),
// so S = xhat . [qk_hi | qk_lo] as ONE N = 64 MMA per k-step (the thread adds the halves) takes a tile's first product from
// 720 to 408 cycles; the tiles' first products are issued back to back at the head of every step's chain.
enum { TC_A = 0, TC_B = 64, TC_R = 128, TC_Z = 192, TC_HN = 256, TC_IN = TC_A, TC_S0 = 320 /* 3 x 64 logits buffers: tile n -> n % 3 */,
       TC_NUMX = TC_A, TC_SSUM = TC_B,        // live only during the token pass, when A and B are dead
       TC_F0 = TC_R,                          // predictor FFN hidden tiles (4 x 64): the GRU accumulators are dead there
       TC_COLS = 512 };

struct FwdUArgs {
    FwdArgs a;
    const unsigned char* wimg;     // blocked weight images (savi_layout.h: WImg)
    WImg wi;
};

// ------------------------------------------------------------------------------------------------
// issuer: attention step products of this CTA's token tiles
//   P1(i): S[i%3]  = xhat_i . [qk_hi | qk_lo]       (A = token tile, K-major;  B = qk operand, N = 64)
//   P2(i): NUMX   += xhat_i^T . [A_hi | A_lo]       (A = token tile, MN-major; B = attention-weight tile)
//          SSUM   += 1 . [A_hi | A_lo]
// ------------------------------------------------------------------------------------------------
struct TokState { uint32_t nseq;                      // issuer: sequence number of the next tile; compute: of the step's first tile
                  uint32_t own_s, own_a; };           // compute warpgroup: its waits on B_SFULL4 / B_AFREE4 so far

__device__ __forceinline__ void issue_token_pass(Ring& r, bool el, unsigned char* sm, const Smem& L, uint64_t* bars, uint32_t tb, int ntile,
                                                 TokState& ts, uint32_t qk_op) {
    int ts0[4] = {0, 0, 0, 0};                              // ring stage of the tile's first token block, per in-flight tile
    const uint32_t ones = dlo_mn(smem_u32(sm + L.ones), 2048);
    const uint32_t rb = smem_u32(r.base);
    const uint32_t n0 = ts.nseq;
    const int la = tok_lookahead(r.nst);                    // tiles whose token blocks the ring can hold besides the one being issued
    auto p2 = [&](int j) {
        const uint32_t n = n0 + (uint32_t)j;
        const int g = (int)(n & 1u);
        mbar_wait(&bars[B_AREADY + g], (n >> 1) & 1u);
        fence_after_sync();
        const int st = ts0[n & 3u];
        const uint32_t x0 = dlo_mn(rb + st * BLK, BLK);
        const uint32_t aw = dlo_mn(smem_u32(sm + (g ? L.aw1 : L.aw0)), BLK);
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        if (el) {
            mma_lo(tb + TC_NUMX, x0, aw, IDESC_MN_MN64, acc0);
            mma_lo(tb + TC_SSUM, ones, aw, IDESC_MN_MN64, acc0);
#pragma unroll
            for (int kt = 1; kt < 8; ++kt) {                              // 16 tokens per k-step
                mma_lo(tb + TC_NUMX, x0 + kt * 128, aw + kt * 128, IDESC_MN_MN64, 1u);
                mma_lo(tb + TC_SSUM, ones, aw + kt * 128, IDESC_MN_MN64, 1u);
            }
            mma_commit(&r.empty[st]);
            mma_commit(&r.empty[st + 1]);
            mma_commit(&bars[B_AFREE4 + ((n + 2u) & 3u)]);                // the weight tile's next writer may proceed
        }
        __syncwarp();
    };
    const uint32_t qk0 = dlo_mn(qk_op, BLK);
    for (int i = 0; i < ntile; ++i) {
        const uint32_t n = n0 + (uint32_t)i, w = n & 3u;
        if (n >= 3u) {                                      // logits buffer n % 3: tile n - 3 (warpgroup (n - 3) & 3) has drained it
            mbar_wait(&bars[B_SFREE4 + ((n - 3u) & 3u)], ((n - 3u) >> 2) & 1u);
            fence_after_sync();
        }
        ts0[w] = r.stage;                                   // the ring has an even number of stages: the pair never wraps
        const uint32_t acc_s = tb + TC_S0 + 64u * (n % 3u);
#pragma unroll
        for (int db = 0; db < 2; ++db) {
            mbar_wait(&r.full[r.stage], r.phase);
            fence_after_sync();
            const uint32_t a = dlo_k(rb + r.stage * BLK);
            if (el) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)            // [S_hi | S_lo]: one N = 64 MMA per k-step
                    mma_lo(acc_s, a + k4 * 2, qk0 + (db * 4 + k4) * 128, IDESC_K_MN64, (db > 0 || k4 > 0) ? 1u : 0u);
            }
            __syncwarp();
            r.advance();
        }
        if (el) mma_commit(&bars[B_SFULL4 + w]);
        __syncwarp();
        if (i >= la) p2(i - la);
    }
    for (int j = ntile > la ? ntile - la : 0; j < ntile; ++j) p2(j);
    ts.nseq = n0 + (uint32_t)ntile;
    if (el) mma_commit(&bars[B_TOK]);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// compute threads: softmax over the slot axis, thread = token, all K <= 24 slots of the token in registers (steve.py:76-83).
// Warpgroup w drains the tiles with sequence number n & 3 == w (savi_umma_clip.cuh, "Token pass").
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax_tiles(const Ctx& c, const Dims& d, int ntile, int tile0, bf16* attn_frame, TokState& ts, long long* dbg) {
    constexpr float LOG2E = 1.4426950408889634f;
    long long ph_last = clock64();
    const int K = c.K;
    const uint32_t sw = (uint32_t)(c.o & 7);
    for (int i = 0; i < ntile; ++i) {
        const uint32_t nseq = ts.nseq + (uint32_t)i;
        if ((nseq & 3u) != (uint32_t)c.wg) continue;
        const int g = (int)(nseq & 1u);
        unsigned char* awrow = c.sm + (g ? c.L.aw1 : c.L.aw0) + c.o * 128;
        mbar_wait(&c.bars[B_SFULL4 + c.wg], ts.own_s & 1u);
        ++ts.own_s;
        fence_after_sync();
        UPH(50);
        const uint32_t scol = c.tb + c.tlane + TC_S0 + 64u * (nseq % 3u);
        float l[KTOK];
        {
            float lo[KTOK];
            tmem_ld16(scol, l); tmem_ld8(scol + 16, l + 16);
            tmem_ld16(scol + 32, lo); tmem_ld8(scol + 48, lo + 16);
            tmem_wait_ld();
#pragma unroll
            for (int s = 0; s < KTOK; ++s) l[s] += lo[s];
        }
        fence_before_sync();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.bars[B_SFREE4 + c.wg]);
        float mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < KTOK; ++s) if (s < K) mx = fmaxf(mx, l[s]);
        mx *= LOG2E;
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < KTOK; ++s) { l[s] = (s < K) ? ex2_fast(fmaf(l[s], LOG2E, -mx)) : 0.f; sum += l[s]; }
        const float scale = rcp_fast(sum);
        const int n = (tile0 + i) * 128 + c.o;                  // token index inside the frame
        const bool valid = n < d.N;
        UPH(51);
        if (nseq > 1u) { mbar_wait(&c.bars[B_AFREE4 + c.wg], ts.own_a & 1u); ++ts.own_a; }   // the second product of tile n - 2 has read this weight tile
        UPH(52);
#pragma unroll
        for (int s = 0; s < KTOK; s += 8) {
            uint32_t hv[4], lv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                l[s + 2 * e] *= scale; l[s + 2 * e + 1] *= scale;                        // P (steve.py:77)
                const float a0 = (valid && s + 2 * e < K) ? l[s + 2 * e] + d.eps : 0.f;  // A = P + eps (:81); padded tokens carry no weight
                const float a1 = (valid && s + 2 * e + 1 < K) ? l[s + 2 * e + 1] + d.eps : 0.f;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                const float2 f2 = __bfloat1622float2(h2);
                hv[e] = *reinterpret_cast<const uint32_t*>(&h2);
                lv[e] = pack_bf16x2(a0 - f2.x, a1 - f2.y);
            }
            const uint32_t ch = (uint32_t)(s >> 3);                                      // 16-byte chunk of the hi half; lo half: chunk 4 + ch
            *reinterpret_cast<uint4*>(awrow + ((ch ^ sw) << 4)) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<uint4*>(awrow + (((4u + ch) ^ sw) << 4)) = make_uint4(lv[0], lv[1], lv[2], lv[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.bars[B_AREADY + g]);
        UPH(53);
        if (attn_frame && valid) {                               // attns (:96): pre-epsilon softmax, [N][K]
            bf16* row = attn_frame + (size_t)n * K;
            if ((K & 7) == 0) {
#pragma unroll
                for (int s = 0; s < KTOK; s += 8)
                    if (s < K) *reinterpret_cast<uint4*>(row + s) = make_uint4(pack_bf16x2(l[s], l[s + 1]), pack_bf16x2(l[s + 2], l[s + 3]),
                                                                               pack_bf16x2(l[s + 4], l[s + 5]), pack_bf16x2(l[s + 6], l[s + 7]));
            } else {
#pragma unroll
                for (int s = 0; s < KTOK; ++s) if (s < K) row[s] = __float2bfloat16_rn(l[s]);
            }
        }
        UPH(54);
    }
    ts.nseq += (uint32_t)ntile;
}

// ------------------------------------------------------------------------------------------------
// predictor attention core on one clip, in shared memory (transformer.py:34-47).  Inputs: this thread's
// q (scaled), k, v columns.  Output: its column of O = softmax(q k^T) v per head.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mha_core(const Ctx& c, int H, const float (&q)[KH], const float (&kx)[KH], const float (&v)[KH],
                                         float (&out)[KH], float* att_g /* nullable: [H][K][K] saved */,
                                         const float* matt /* nullable: training-mode dropout mask of the probabilities [H][K][K] */) {
    // rows of sQ / sK / sV are 16-byte aligned (ld % 4 == 0) and 4 banks apart: float4 reads along a head's features are
    // conflict-free across slots, column reads [k][o] across o.  Attention rows are padded to ka = 4-aligned K.
    const int K = c.K, ld = MHA_LD, ka = mha_ka(K), dh = F / H;
    float* sQ = reinterpret_cast<float*>(c.sm + c.L.aw0);            // q, k tiles in aw0 | aw1 (the LayerNorm scratch aliases aw1: idle here)
    float* sK = sQ + K * ld;
    float* sA = reinterpret_cast<float*>(c.sm + c.L.opB);            // [H*K][ka] logits -> probabilities (opB | opC are free here)
    float* sAT = sA + H * K * ka;                                    // [H][K key][ka query]: the transposed probabilities
    float* sV = sAT + H * K * ka;                                    // v tile behind them, still inside opB | opC (savi_umma_mha_fits)
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        const int k = c.k0 + kk;
        if (kk < c.nk) { sQ[k * ld + c.o] = q[kk]; sK[k * ld + c.o] = kx[kk]; sV[k * ld + c.o] = v[kk]; }
    }
    bar_sync_compute();
    // logits: 4 threads per (head, query) row, each a contiguous share of <= MHA_JP keys with independent accumulators
    // (the query row chunk is loaded once and reused; 16-byte conflict-free reads)
#pragma unroll 1
    for (int row = c.tid >> 2; row < H * K; row += NCT / 4) {
        const int jq = c.tid & 3, JP = (K + 3) >> 2, j0 = jq * JP;
        const int h = row / K, i = row - h * K;
        const float* a = sQ + i * ld + h * dh;
        const float* b = sK + h * dh;
        float acc[MHA_JP];
#pragma unroll
        for (int jj = 0; jj < MHA_JP; ++jj) acc[jj] = 0.f;
#pragma unroll 1
        for (int e = 0; e < dh; e += 4) {
            const float4 x = ld4(a + e);
#pragma unroll
            for (int jj = 0; jj < MHA_JP; ++jj) {
                const float4 y = ld4(b + min(j0 + jj, K - 1) * ld + e);
                acc[jj] = fmaf(x.x, y.x, acc[jj]); acc[jj] = fmaf(x.y, y.y, acc[jj]); acc[jj] = fmaf(x.z, y.z, acc[jj]); acc[jj] = fmaf(x.w, y.w, acc[jj]);
            }
        }
#pragma unroll
        for (int jj = 0; jj < MHA_JP; ++jj) if (jj < JP && j0 + jj < K) sA[row * ka + j0 + jj] = acc[jj];
    }
    bar_sync_compute();
    // softmax over the keys: one warp per (head, query) row, lane = key; three rows in flight per warp
#pragma unroll 1
    for (int row0 = c.warp; row0 < H * K; row0 += 3 * NCW) {
        float x[3], e[3], mx[3], sm[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int row = row0 + r * NCW;
            x[r] = (row < H * K && c.lane < K) ? sA[row * ka + c.lane] : -INFINITY;
            mx[r] = x[r];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < 3; ++r) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) { e[r] = (c.lane < K && row0 + r * NCW < H * K) ? expf(x[r] - mx[r]) : 0.f; sm[r] = e[r]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < 3; ++r) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], o);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int row = row0 + r * NCW;
            if (row < H * K && c.lane < K) {
                const float pr = e[r] * (1.0f / sm[r]);
                // attn_dropout (transformer.py:44) acts on what multiplies V; the saved matrix stays the undropped softmax
                sAT[((row / K) * K + c.lane) * ka + row % K] = matt ? pr * matt[row * K + c.lane] : pr;
                if (att_g) att_g[row * K + c.lane] = pr;
            }
        }
    }
    bar_sync_compute();
    const float* at = sAT + (c.o / dh) * K * ka + c.k0;
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) out[kk] = 0.f;
    if (c.nk > 0) {
#pragma unroll 4
        for (int j = 0; j < K; ++j) {
            const float vj = sV[j * ld + c.o];
            const float4 p0 = ld4(at + j * ka), p1 = ld4(at + j * ka + 4);       // this warpgroup's 8 queries (warp-uniform address)
            out[0] = fmaf(p0.x, vj, out[0]); out[1] = fmaf(p0.y, vj, out[1]); out[2] = fmaf(p0.z, vj, out[2]); out[3] = fmaf(p0.w, vj, out[3]);
            out[4] = fmaf(p1.x, vj, out[4]); out[5] = fmaf(p1.y, vj, out[5]); out[6] = fmaf(p1.z, vj, out[6]); out[7] = fmaf(p1.w, vj, out[7]);
        }
    }
    bar_sync_compute();                                              // sQ/sK/sV/sA are free again
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// KFIX / CNFIX / HFIX: slots, CTAs per clip and predictor heads when the launch site knows them at compile time (0: read them from
// the shape).  The MOVi configurations get their own instances: with K fixed, the per-slot predicates of the token threads and
// the row / K index arithmetic of the predictor fold away; everything else runs the generic <0, 0, 0> instance.
template <int KFIX, int CNFIX, int HFIX>
__global__ void __launch_bounds__(NTHREADS, 1) savi_fwd_umma_kernel(const __grid_constant__ FwdUArgs ua) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const FwdArgs& a = ua.a;
    const Dims& d = a.d;
    const ParamOff& po = a.po;
    unsigned char* sm = smem_raw;
    if ((smem_u32(sm) & 1023u) != 0u) __trap();                // SWIZZLE_128B operands need a 1024-byte aligned base
    // (warp index through a lane-0 broadcast: ptxas then treats it, and every role branch on it, as warp-uniform)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int CN = CNFIX ? CNFIX : d.CN, b = blockIdx.x / CN, rank = blockIdx.x % CN;
    const int K = KFIX ? KFIX : d.K, B = d.B, KP = KFIX ? ((KFIX + 3) & ~3) : d.KP;
    const int HEADS = HFIX ? HFIX : d.heads;
    const Smem L = plan_smem(K, CN, false);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const bool lead = (rank == 0);
    // both CTAs of a pair hold bit-identical slot state: the saved-for-backward records are split between them
    const bool svA = lead, svB = (rank == CN - 1);
    // token tiles of this CTA
    const int per = (d.NTILE + CN - 1) / CN;
    const int tile0 = min(d.NTILE, rank * per), ntile = min(d.NTILE, tile0 + per) - tile0;
    const float* P = a.packed;
    float* fb = reinterpret_cast<float*>(a.saved + a.sl.fbase);
    const unsigned char* ximg = a.saved + a.sl.ximg;

    // ---- one-time setup ----
    for (int i = tid * 16; i < L.bars; i += NTHREADS * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    for (int i = tid; i < 2048; i += NTHREADS) reinterpret_cast<uint16_t*>(sm + L.ones)[i] = 0x3F80;                     // bf16 1.0
    if (tid == 0) {
        for (int s = 0; s < L.nst; ++s) { mbar_init(&bars[B_FULL + s], 1); mbar_init(&bars[B_EMPTY + s], 1); }
        mbar_init(&bars[B_OPND], NCW); mbar_init(&bars[B_ACC], 1); mbar_init(&bars[B_TOK], 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&bars[B_AREADY + g], 4);                                         // one warpgroup (4 warps) per tile
            mbar_init(&bars[B_INBOX + g], NCW);
        }
        for (int f = 0; f < 4; ++f) { mbar_init(&bars[B_FACC + f], 1); mbar_init(&bars[B_FOPND + f], NCW); mbar_init(&bars[B_SFULL4 + f], 1); mbar_init(&bars[B_AFREE4 + f], 1); mbar_init(&bars[B_SFREE4 + f], 4); }
        mbar_init_fence();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TC_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (CN > 1) { cluster_arrive(); cluster_wait(); }          // the peer's barriers exist before anyone signals them
    const uint32_t tb = *tmem_slot;

    Ring ring;
    ring.base = sm + L.ring; ring.full = &bars[B_FULL]; ring.empty = &bars[B_EMPTY]; ring.nst = L.nst; ring.stage = 0; ring.phase = 0;
    ring.wait_cycles = nullptr;

    if (warp == W_PROD) {
        // =====================================================================================
        // producer: the static block schedule (must mirror the issuer's consumption order)
        // =====================================================================================
        if (lane == 0) {
            const unsigned char* W = ua.wimg;
            const WImg& wi = ua.wi;
            for (int t = 0; t < d.T; ++t) {
                const unsigned char* xf = ximg + ((size_t)(b * d.T + t) * d.NTILE + tile0) * 2 * BLK;
                for (int it = 0; it < d.I; ++it) {
                    prod_blocks(ring, W + wi.wqk, 2 * NBW);
                    prod_token_blocks(ring, xf, 2 * ntile, 16);
                    prod_blocks(ring, W + wi.whh, 6 * NBW);
                    prod_blocks(ring, W + wi.wg, 6 * NBW);
                    if (it < d.I - 1) { prod_blocks(ring, W + wi.w1, 2 * NBW); prod_blocks(ring, W + wi.w2, 2 * NBW); }
                }
                if (t < d.T - 1) {
                    for (int j = 0; j < d.blocks; ++j) {
                        const WImgBlock& wb = wi.blk[j];
                        prod_blocks(ring, W + wb.pq, 2 * NBW); prod_blocks(ring, W + wb.pk, 2 * NBW); prod_blocks(ring, W + wb.pv, 2 * NBW);
                        prod_blocks(ring, W + wb.po, 2 * NBW);
                        prod_blocks(ring, W + wb.f1, 8 * NBW);
                        prod_blocks(ring, W + wb.f2, 8 * NBW);
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // =====================================================================================
        // tcgen05.mma issuer
        // =====================================================================================
        {
            const bool el = elect_one();
            uint32_t ph_opnd = 0, pcall = 0;
            TokState ts = {0, 0, 0};
            const uint32_t opA = smem_u32(sm + L.opA), opB = smem_u32(sm + L.opB), opC = smem_u32(sm + L.opC);
            const uint32_t aw0 = smem_u32(sm + L.aw0), aw1 = smem_u32(sm + L.aw1);
            // development counters of the issuer (CTA 0): [60] waiting for operands, [61] waiting for ring blocks, [62] token pass, [63] total
#ifdef SAVI_PHASE_PROFILE
            long long* idbg = (a.dbg && blockIdx.x == 0 && el) ? reinterpret_cast<long long*>(sm + L.bars + NBAR * 8 + 16) : nullptr;
#else
            long long* const idbg = nullptr;
#endif
            long long i_opnd = 0, i_ring = 0, i_tok = 0;
            const long long i_t0 = clock64();
#ifdef SAVI_PHASE_PROFILE
            if (a.dbg && blockIdx.x == 0) ring.wait_cycles = &i_ring;
#endif
            auto wait_opnd = [&]() {
                const long long t0 = clock64();
                mbar_wait(&bars[B_OPND], ph_opnd); ph_opnd ^= 1u; fence_after_sync();
                i_opnd += clock64() - t0;
            };
            for (int t = 0; t < d.T; ++t) {
                for (int it = 0; it < d.I; ++it) {
                    wait_opnd();                                                          // s~ in opA, h_prev in opC
                    issue_linear(ring, el, opA, tb + TC_B, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);          // qk = s~ . wqk^T  (steve.py:75 and the fold of :61,63,76 in one product)
                    wait_opnd();                                                          // qk in opA
                    { const long long t0 = clock64(); issue_token_pass(ring, el, sm, L, bars, tb, ntile, ts, opA); i_tok += clock64() - t0; }
                    // GRU hidden-side product, off the critical path: runs while the compute threads combine the partial sums.
                    // (It must not be streamed while token tiles are held in the ring: the ring is filled in order.)
                    issue_linear(ring, el, opC, tb + TC_R, 3, 2, false);                      // R, Z, HN = W_hh . h_prev
                    wait_opnd();                                                          // Ux in opB
                    issue_linear(ring, el, opB, tb + TC_R, 2, 2, true);                                      // R, Z += (W_i{r,z} W_v) . Ux   (updates :83 folded into :87)
                    issue_linear(ring, el, opB, tb + TC_IN, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);         // IN = (W_in W_v) . Ux
                    if (it < d.I - 1) {
                        wait_opnd();                                                      // LN_m(h') in opB
                        issue_linear(ring, el, opB, tb + TC_A, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);      // mlp.0 (:92)
                        wait_opnd();                                                      // a in opA
                        issue_linear(ring, el, opA, tb + TC_B, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);      // mlp.2
                    }
                }
                if (t < d.T - 1) {
                    for (int j = 0; j < d.blocks; ++j) {
                        wait_opnd();                                                      // y in opA
                        issue_linear(ring, el, opA, tb + TC_R, 3, 2, false); if (el) mma_commit(&bars[B_ACC]);      // q, k, v -> R, Z, HN columns
                        wait_opnd();                                                      // attention output in opB
                        issue_linear(ring, el, opB, tb + TC_A, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);      // proj_o
                        wait_opnd();                                                      // LN2 in opA
                        for (int f = 0; f < 4; ++f) { issue_linear(ring, el, opA, tb + TC_F0 + 64 * f, 1, 2, false); if (el) mma_commit(&bars[B_FACC + f]); }   // ffn.0
                        for (int f = 0; f < 4; ++f) {                                     // ffn.2, contraction split in 4 chunks of 128
                            mbar_wait(&bars[B_FOPND + f], pcall & 1u); fence_after_sync();
                            issue_linear(ring, el, f == 0 ? opB : f == 1 ? opC : f == 2 ? aw0 : aw1, tb + TC_B, 1, 2, f > 0);
                        }
                        if (el) mma_commit(&bars[B_ACC]);
                        ++pcall;
                    }
                }
            }
            if (idbg) { a.dbg[60] += i_opnd; a.dbg[61] += i_ring; a.dbg[62] += i_tok; a.dbg[63] += clock64() - i_t0; }
        }
    } else {
        // =====================================================================================
        // compute threads
        // =====================================================================================
        Ctx c;
        ctx_init(c, tid, K, sm, L, tb, bars);
        const int o = c.o;
        TokState ts = {0, 0, 0};
        // per-feature parameters of this thread
        const float g_s = P[po.ln_s_w + o], b_s = P[po.ln_s_b + o], g_m = P[po.ln_m_w + o], b_m = P[po.ln_m_b + o];
        const float b_r = P[po.bih + o] + P[po.bhh + o], b_z = P[po.bih + F + o] + P[po.bhh + F + o];
        const float bin = P[po.bih + 2 * F + o], bhn = P[po.bhh + 2 * F + o];
        const float b1 = P[po.b1 + o], b2 = P[po.b2 + o];
        float h[KH], y[KH];
        {   // slots0 = mu + exp(log_sigma) * noise   (steve.py:56-57)
            const float mu = P[po.slot_mu + o], sg = expf(P[po.slot_log_sigma + o]);
            load_field(c, a.noise + (size_t)b * K * F, F, o, y);
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) h[kk] = mu + sg * y[kk];
        }
        uint32_t step = 0, pcall = 0;
        long long* sdbg = reinterpret_cast<long long*>(sm + L.bars + NBAR * 8 + 16);
        if (a.dbg && blockIdx.x == 0 && tid == 0) for (int i = 0; i < 64; ++i) sdbg[i] = 0;
#ifdef SAVI_PHASE_PROFILE
        long long* dbg = (a.dbg && blockIdx.x == 0 && tid == 0) ? sdbg : nullptr;      // counters in shared memory: a global RMW per probe would stall the warp
#else
        long long* const dbg = nullptr;      // production build: the ~25 UPH probes (a per-thread branch + a live 64-bit timestamp each) fold away
#endif
        long long ph_last = clock64();
        for (int t = 0; t < d.T; ++t) {
            for (int it = 0; it < d.I; ++it, ++step) {
                const int64_t s = (int64_t)t * d.I + it;
                // ---- slots_prev, LayerNorm, q ----
                // (the saved-for-backward records are stored AFTER the operand hand-over of their phase: the stores then overlap
                // the tensor-core round trip instead of sitting in front of it on the serial chain)
                write_operand_f16(c, L.opC, h);
                layer_norm(c, h, y, g_s, b_s, d.ln_eps, lead ? reinterpret_cast<float2*>(fb + a.sl.lns) + (s * B + b) * K : nullptr);   // :72
                write_operand_f16(c, L.opA, y);
                signal_operand(c);
                if (svA) save_field(c, frow(fb, a.sl.hp, s, b, B, K, F), F, o, h);
                if (svB) save_field(c, frow(fb, a.sl.st, s, b, B, K, F), F, o, y);
                UPH(1);
                wait_acc(c); UPH(4); load_acc(c, TC_B, y);                                  // qk = Ds^-1/2 (s~ Wq^T) Wk, the scale folded into wqk
                write_operand(c, L.opA, y);                                                 // bf16 hi | lo: multiplied by the bf16 token tiles
                signal_operand(c);
                if (svB) save_field(c, frow(fb, a.sl.qk, s, b, B, K, F), F, o, y);
                // ---- attention step over the token tiles ----
                bf16* attn_frame = (it == d.I - 1) ? reinterpret_cast<bf16*>(a.attn_out) + ((size_t)b * d.T + t) * d.N * K : nullptr;
                UPH(5);
                softmax_tiles(c, d, ntile, tile0, attn_frame, ts, dbg);
                UPH(6);
                mbar_wait(&bars[B_TOK], step & 1u);
                fence_after_sync();
                UPH(7);
                float num[KH], den[KH];
                load_acc(c, TC_NUMX, num); load_acc(c, TC_SSUM, den);
                if (CN > 1) {                                                              // exchange the partial sums with the peer CTA
                    // inbox layout [feature o][KR8 slots] (+ [KR8] sums): a thread's 8 slot values are 32 contiguous bytes, two
                    // 16-byte remote stores instead of eight scalar ones
#if SAVI_FWD_RING8
                    const int buf = 0;                       // single inbox: the peer must have consumed what this CTA sent in the previous step
                    if (step > 0) mbar_wait_cluster(&bars[B_INBOX + 1], (step - 1) & 1u);     // (it says so on OUR barrier)
#else
                    const int buf = step & 1;
#endif
                    const int KR8 = (K + 7) & ~7;
                    float* ib = reinterpret_cast<float*>(sm + L.inbox + buf * L.inbox_stride);
                    const uint32_t peer = rank ^ 1u;
                    if (c.nk > 0) {
                        const uint32_t rb = map_to_rank(ib, peer) + (uint32_t)(o * KR8 + c.k0) * 4u;
                        st_cluster_f4(rb, make_float4(num[0], num[1], num[2], num[3]));
                        st_cluster_f4(rb + 16u, make_float4(num[4], num[5], num[6], num[7]));
                        if (o == 0) {
                            const uint32_t rd = map_to_rank(ib + F * KR8 + c.k0, peer);
                            st_cluster_f4(rd, make_float4(den[0], den[1], den[2], den[3]));
                            st_cluster_f4(rd + 16u, make_float4(den[4], den[5], den[6], den[7]));
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(map_to_rank(&bars[B_INBOX + buf], peer));
                    UPH(8);
#if SAVI_FWD_RING8
                    mbar_wait_cluster(&bars[B_INBOX], step & 1u);
#else
                    mbar_wait_cluster(&bars[B_INBOX + buf], (step >> 1) & 1u);
#endif
                    UPH(9);
                    if (c.nk > 0) {
                        float pn[KH], pd[KH];
                        *reinterpret_cast<float4*>(pn) = ld4(ib + o * KR8 + c.k0); *reinterpret_cast<float4*>(pn + 4) = ld4(ib + o * KR8 + c.k0 + 4);
                        *reinterpret_cast<float4*>(pd) = ld4(ib + F * KR8 + c.k0); *reinterpret_cast<float4*>(pd + 4) = ld4(ib + F * KR8 + c.k0 + 4);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) {
                            // fixed order rank 0 + rank 1 on both CTAs: their slot states stay bit-identical
                            num[kk] = lead ? num[kk] + pn[kk] : pn[kk] + num[kk];
                            den[kk] = lead ? den[kk] + pd[kk] : pd[kk] + den[kk];
                        }
                        loaded_before_release(pn[0]); loaded_before_release(pn[4]); loaded_before_release(pd[0]); loaded_before_release(pd[4]);
                    }
#if SAVI_FWD_RING8
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(map_to_rank(&bars[B_INBOX + 1], peer));     // consumed: the peer may send again
#endif
                }
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) y[kk] = (kk < c.nk) ? num[kk] * rcp_fast(den[kk]) : 0.f;      // Ux (:82-83)
                write_operand_f16(c, L.opB, y);
                signal_operand(c);
                if (lead) {
                    save_field(c, frow(fb, a.sl.ux, s, b, B, K, F), F, o, y);
                    if (o == 0) {
                        float* r_ss = fb + a.sl.ssum + (s * B + b) * KP + c.k0;
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) r_ss[kk] = den[kk];
                    }
                }
                UPH(10);
                // ---- GRUCell (:87-89): the input-side product takes Ux directly (updates = Ux Wv^T is never formed) ----
                UPH(12);
                wait_acc(c);
                UPH(13);
                {
                    float r_[KH], z_[KH], n_[KH], hn[KH];
                    load_acc(c, TC_R, r_); load_acc(c, TC_Z, z_); load_acc(c, TC_IN, n_); load_acc(c, TC_HN, hn);
                    const bool mlp = it < d.I - 1;
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) {
                        const float hnb = hn[kk] + bhn;
                        const float vr = sigmoid_fast(r_[kk] + b_r);
                        const float vz = sigmoid_fast(z_[kk] + b_z);
                        const float vn = tanh_fast(n_[kk] + bin + vr * hnb);
                        r_[kk] = vr; z_[kk] = vz; n_[kk] = vn; hn[kk] = hnb;
                        h[kk] = fmaf(vz, h[kk] - vn, vn);                                  // (1 - z) n + z h
                    }
                    if (mlp) {                                                             // residual MLP (:92-93)
                        const int64_t smi = (int64_t)t * (d.I - 1) + it;
                        UPH(14);
                        layer_norm(c, h, y, g_m, b_m, d.ln_eps, svB ? reinterpret_cast<float2*>(fb + a.sl.lnm) + (smi * B + b) * K : nullptr);
                        write_operand_f16(c, L.opB, y);
                        signal_operand(c);
                        if (svA) {
                            save_field(c, frow(fb, a.sl.r, s, b, B, K, F), F, o, r_);
                            save_field(c, frow(fb, a.sl.n, s, b, B, K, F), F, o, n_);
                            save_field(c, frow(fb, a.sl.hg, smi, b, B, K, F), F, o, h);
                        }
                        if (svB) {
                            save_field(c, frow(fb, a.sl.z, s, b, B, K, F), F, o, z_);
                            save_field(c, frow(fb, a.sl.ghn, s, b, B, K, F), F, o, hn);
                            save_field(c, frow(fb, a.sl.m, smi, b, B, K, F), F, o, y);
                        }
                        UPH(15);
                        wait_acc(c); UPH(16); load_acc(c, TC_A, y);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) y[kk] = fmaxf(y[kk] + b1, 0.f);
                        write_operand_f16(c, L.opA, y);
                        signal_operand(c);
                        if (svA) save_field(c, frow(fb, a.sl.a, smi, b, B, K, F), F, o, y);
                        UPH(17);
                        wait_acc(c); UPH(18); load_acc(c, TC_B, y);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) h[kk] += y[kk] + b2;
                    } else {
                        if (svA) {
                            save_field(c, frow(fb, a.sl.r, s, b, B, K, F), F, o, r_);
                            save_field(c, frow(fb, a.sl.n, s, b, B, K, F), F, o, n_);
                        }
                        if (svB) {
                            save_field(c, frow(fb, a.sl.z, s, b, B, K, F), F, o, z_);
                            save_field(c, frow(fb, a.sl.ghn, s, b, B, K, F), F, o, hn);
                        }
                    }
                    UPH(19);
                }
            }
            if (svB) save_field(c, a.slots_out + ((size_t)b * d.T + t) * K * F, F, o, h);      // collect (:96-97)
            if (t < d.T - 1) {
                // ---- predictor (:100; transformer.py:106-114) ----
                if (svA) save_field(c, fb + a.sl.px0 + ((size_t)t * B + b) * K * F, F, o, h);
                const float hscale = 1.0f / sqrtf((float)(F / HEADS));
                float x[KH];
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) x[kk] = h[kk];
                long long pp_last = clock64();
#define PPH(id) do { if (dbg) { const long long t_ = clock64(); dbg[id] += t_ - pp_last; pp_last = t_; } } while (0)
                for (int j = 0; j < d.blocks; ++j) {
                    const int64_t f = (int64_t)j * (d.T - 1) + t;
                    const BlockOff& bo = po.blk[j];
                    float yv[KH], q[KH], kx[KH], v[KH], x1[KH];
                    layer_norm(c, x, yv, P[bo.ln1_w + o], P[bo.ln1_b + o], d.ln_eps);
                    if (svB) save_field(c, frow(fb, a.sl.py, f, b, B, K, F), F, o, yv);
                    write_operand_f16(c, L.opA, yv);
                    signal_operand(c);
                    
                    wait_acc(c);
                    PPH(21);
                    load_acc(c, TC_R, q); load_acc(c, TC_Z, kx); load_acc(c, TC_HN, v);
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) q[kk] *= hscale;
                    if (svA) {
                        save_field(c, frow(fb, a.sl.pq, f, b, B, K, F), F, o, q);
                        save_field(c, frow(fb, a.sl.pv, f, b, B, K, F), F, o, v);
                    }
                    if (svB) save_field(c, frow(fb, a.sl.pk, f, b, B, K, F), F, o, kx);
                    float ov[KH];
                    // training-mode dropout masks of this block evaluation (nullptr otherwise): savi_args.h, DropLayout
                    const DropLayout dl = savi_dropout_layout(d);
                    const float* m_att = a.drop ? a.drop + dl.att + (f * B + b) * ((int64_t)HEADS * K * K) : nullptr;
                    const float* m_out = a.drop ? a.drop + dl.out + (f * B + b) * ((int64_t)K * F) : nullptr;
                    const float* m_ffn = a.drop ? a.drop + dl.ffn + (f * B + b) * ((int64_t)K * F) : nullptr;
                    mha_core(c, HEADS, q, kx, v, ov, svB ? fb + a.sl.patt + (f * B + b) * ((int64_t)HEADS * K * K) : nullptr, m_att);
                    if (svA) save_field(c, frow(fb, a.sl.po, f, b, B, K, F), F, o, ov);
                    write_operand_f16(c, L.opB, ov);
                    signal_operand(c);
                    PPH(22);
                    wait_acc(c);  load_acc(c, TC_A, x1);
                    if (m_out) {                                                           // output_dropout (transformer.py:48)
                        float mo[KH];
                        load_field(c, m_out, F, o, mo);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) x1[kk] *= mo[kk];
                    }
                    // the first block adds the residual to the NORMALISED input (transformer.py:75-78)
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) x1[kk] += (j == 0) ? yv[kk] : x[kk];
                    if (svB) save_field(c, frow(fb, a.sl.px1, f, b, B, K, F), F, o, x1);
                    layer_norm(c, x1, yv, P[bo.ln2_w + o], P[bo.ln2_b + o], d.ln_eps);
                    if (svA) save_field(c, frow(fb, a.sl.pl2, f, b, B, K, F), F, o, yv);
                    write_operand_f16(c, L.opA, yv);
                    signal_operand(c);
                    PPH(23);
#pragma unroll
                    for (int ff = 0; ff < 4; ++ff) {
                        mbar_wait(&bars[B_FACC + ff], pcall & 1u); fence_after_sync();
                        load_acc(c, TC_F0 + 64 * ff, yv);
                        const float bb = P[bo.f1b + ff * F + o];
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) yv[kk] = fmaxf(yv[kk] + bb, 0.f);
                        if ((ff & 1) ? svB : svA) save_field(c, frow(fb, a.sl.pf, f, b, B, K, 4 * F), 4 * F, ff * F + o, yv);
                        write_operand_f16(c, ff == 0 ? L.opB : ff == 1 ? L.opC : ff == 2 ? L.aw0 : L.aw1, yv);
                        signal_operand(c, B_FOPND + ff);
                    }
                    ++pcall;
                    
                    wait_acc(c);  load_acc(c, TC_B, yv);
                    const float bb2 = P[bo.f2b + o];
                    if (m_ffn) {                                                           // the Dropout closing the FFN (transformer.py:68)
                        float mf[KH];
                        load_field(c, m_ffn, F, o, mf);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) x[kk] = x1[kk] + (yv[kk] + bb2) * mf[kk];
                    } else {
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) x[kk] = x1[kk] + yv[kk] + bb2;
                    }
                    if (svB) save_field(c, frow(fb, a.sl.px2, f, b, B, K, F), F, o, x);
                }
                layer_norm(c, x, h, P[po.lnf_w + o], P[po.lnf_b + o], d.ln_eps);
                PPH(24);
                UPH(20);
            }
        }
        if (dbg) for (int i = 0; i < 60; ++i) if (sdbg[i]) a.dbg[i] += sdbg[i];
    }
    // ---- teardown ----
    __syncwarp();
    fence_before_sync();
    __syncthreads();
    if (CN > 1) { cluster_arrive(); cluster_wait(); }          // nobody exits while its inbox may still be written
    if (warp == W_MMA) tmem_dealloc(tb, TC_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int savi_umma_mha_fits(int K, int heads) {
    if (heads < 1 || F % heads || (F / heads) % 4) return 0;
    const int KR = (K + 3) & ~3, ka = mha_ka(K);
    const int win_bwd = 5 * OPB;                              // opA .. aw1
    (void)KR;
    // forward: q, k tiles in aw0 | aw1; the two attention matrices and the v tile in opB | opC
    return heads * K * K <= 8 * NCT && mha_bwd_bytes(K, heads) <= win_bwd &&
           2 * K * MHA_LD * 4 <= 2 * OPB && (2 * heads * K * ka + K * MHA_LD) * 4 <= 2 * OPB;
}

int savi_fwd_umma_smem_bytes(const Dims& d) {
    return plan_smem(d.K, d.CN, false).total;
}

cudaError_t savi_launch_fwd_umma(const FwdArgs& a, const unsigned char* wimg, const WImg& wi, cudaStream_t st) {
    FwdUArgs ua;
    ua.a = a; ua.wimg = wimg; ua.wi = wi;
    ua.a.smem_bytes = savi_fwd_umma_smem_bytes(a.d);
    // instances: C2 (K = 24, 4 heads) and C4 (K = 11), as CTA pairs (B <= 74 clips per GPU) and as single CTAs; generic otherwise
    void (*kern)(FwdUArgs) = savi_fwd_umma_kernel<0, 0, 0>;
    const Dims& dd = ua.a.d;
    if (dd.heads == 4 && dd.K == 24) kern = dd.CN == 2 ? savi_fwd_umma_kernel<24, 2, 4> : dd.CN == 1 ? savi_fwd_umma_kernel<24, 1, 4> : kern;
    else if (dd.heads == 4 && dd.K == 11) kern = dd.CN == 2 ? savi_fwd_umma_kernel<11, 2, 4> : dd.CN == 1 ? savi_fwd_umma_kernel<11, 1, 4> : kern;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ua.a.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.d.B * a.d.CN);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = ua.a.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.d.CN; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ua);
}
