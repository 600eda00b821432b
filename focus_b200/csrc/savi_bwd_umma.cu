// tcgen05 backward clip kernel: BPTT through the whole recurrence (+ predictor) of one clip per CTA / CTA pair.
//
// The reference has no explicit backward (autograd through steve.py:52-105); the closed form is SURVEY.md
// Appendix A.2 in the folded form, restated and validated against the reference's autograd in
// oracle/savi_numpy.py.  Execution model: savi_umma_clip.cuh (feature-per-thread, transposed products).
// Every dX = dY . W product uses the "backward orientation" weight images (rows = input feature).
// Writes the staged (dY, X) operands of the weight-gradient GEMMs and the per-token coefficients
// (dL^T, W^T) + dUx of the token-parallel d_inputs kernel, exactly like the mma.sync kernel (savi_bwd.cu).
#include "savi_umma_clip.cuh"

using namespace uc;

#define UPH(id) do { if (dbg) { const long long t_ = clock64(); dbg[id] += t_ - ph_last; ph_last = t_; } } while (0)

// TMEM columns (64 each)
enum { TB_A = 0, TB_B = 64, TB_HP = 128, TB_DQK = 192, TB_S0 = 256 /* 4 x 64: [S logits 32 | G 32] per warpgroup */,
       TB_F0 = TB_S0,                         // predictor: the four ffn.2^T tiles live where the token-pass accumulators are
       TB_COLS = 512 };
enum { B_HP = B_OPND2 };                      // W_hh^T product complete (its operands may be overwritten)

struct BwdUArgs {
    BwdArgs a;
    long long* trace;              // development (SAVI_DX_TRACE): globaltimer of CTA 0 at start / end
    const unsigned char* wimg;
    WImg wi;
};

struct TokState { uint32_t nseq;                      // issuer: sequence number of the next tile; compute: of the step's first tile
                  uint32_t own_s, own_a; };           // compute warpgroup: its waits on B_SFULL4 / B_AFREE4 so far

// operand buffers: X0..X3 = opA, opB, opC, aw0 of the shared plan; aw1 + scratch hold the float2 LayerNorm scratch
__device__ __forceinline__ int xop(const Smem& L, int i) { return i == 0 ? L.opA : i == 1 ? L.opB : i == 2 ? L.opC : L.aw0; }

// ------------------------------------------------------------------------------------------------
// issuer: attention-step backward products of this CTA's token tiles
//   P1(i): [S | G] = xhat_i . [qk_hi | dUx_hi]  +  xhat_i . [qk_lo | dUx_lo]     (two N = 64 MMAs per k-step into one 64-column
//          accumulator).  S and G share their A operand (the token tile), and one tcgen05.mma costs 45 / 51 / 65 cycles at
//          N = 32 / 64 / 128 (tools/umma_rate.cu: a ~40-cycle floor per instruction, the 4 KB A read), so pairing them per
//          precision half instead of four N = 32 MMAs takes the tile's first products from 1440 to 816 cycles; the four
//          tiles' first products are issued back to back, so this is on the serial chain of every step.
//   P2(i): DQK   += xhat_i^T . [dL_hi | dL_lo]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void issue_token_pass_bwd(Ring& r, bool el, uint64_t* bars, uint32_t tb, int ntile, TokState& ts,
                                                     uint32_t hi_op, uint32_t lo_op, uint32_t dl0, uint32_t dl1) {
    int ts0[4] = {0, 0, 0, 0};                              // ring stage of the tile's first token block, per in-flight tile
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
        const uint32_t dl = dlo_mn(g ? dl1 : dl0, BLK);
        if (el) {
            mma_lo(tb + TB_DQK, x0, dl, IDESC_MN_MN64, j > 0 ? 1u : 0u);
#pragma unroll
            for (int kt = 1; kt < 8; ++kt) mma_lo(tb + TB_DQK, x0 + kt * 128, dl + kt * 128, IDESC_MN_MN64, 1u);
            mma_commit(&r.empty[st]);
            mma_commit(&r.empty[st + 1]);
            mma_commit(&bars[B_AFREE4 + ((n + 2u) & 3u)]);                // the dL tile's next writer may proceed
        }
        __syncwarp();
    };
    const uint32_t hi0 = dlo_mn(hi_op, BLK), lo0 = dlo_mn(lo_op, BLK);
    for (int i = 0; i < ntile; ++i) {
        const uint32_t n = n0 + (uint32_t)i, w = n & 3u;
        mbar_wait(&bars[B_SFREE4 + w], ((n >> 2) & 1u) ^ 1u);           // warpgroup w has drained its previous logits
        fence_after_sync();
        ts0[w] = r.stage;
        const uint32_t acc_sg = tb + TB_S0 + 64u * w;
#pragma unroll
        for (int db = 0; db < 2; ++db) {
            mbar_wait(&r.full[r.stage], r.phase);
            fence_after_sync();
            const uint32_t a = dlo_k(rb + r.stage * BLK);
            if (el) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {          // hi and lo halves accumulate into the SAME 64 columns [S | G]
                    const uint32_t acc = (db > 0 || k4 > 0) ? 1u : 0u;
                    const uint32_t ko = (uint32_t)(db * 4 + k4) * 128u;
                    mma_lo(acc_sg, a + k4 * 2, hi0 + ko, IDESC_K_MN64, acc);
                    mma_lo(acc_sg, a + k4 * 2, lo0 + ko, IDESC_K_MN64, 1u);
                }
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
// compute threads: recompute P, form dL and W for this warpgroup's token tiles (thread = token, all K <= 24 slots of the
// token in registers; tile ownership as in the forward: savi_umma_clip.cuh, "Token pass")
//   dP = (G - c)/S (+ grad_attn);  dL = P (dP - <P, dP>);  W = (P + eps)/S          
// cv: shared [64] floats: c[k] at [k], 1/S[k] at [32 + k].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void softmax_bwd_tiles(const Ctx& c, const Dims& d, int ntile, int tile0, const bf16* gattn_frame,
                                                  unsigned char* coef, const float* cv, int dl0_off, int dl1_off, TokState& ts, long long* dbg) {
    constexpr float LOG2E = 1.4426950408889634f;
    long long ph_last = clock64();
    const int K = c.K;
    const uint32_t scol = c.tb + c.tlane + TB_S0 + 64u * (uint32_t)c.wg;
    const uint32_t gcol = scol + 32u;
    const uint32_t sw = (uint32_t)(c.o & 7);
    for (int i = 0; i < ntile; ++i) {
        const uint32_t nseq = ts.nseq + (uint32_t)i;
        if ((nseq & 3u) != (uint32_t)c.wg) continue;
        const int g = (int)(nseq & 1u);
        unsigned char* dlrow = c.sm + (g ? dl1_off : dl0_off) + c.o * 128;
        const int n = (tile0 + i) * 128 + c.o;
        const bool valid = n < d.N;
        // upstream gradient of the attention map (last iteration only), kept packed (bf16 pairs) until it is consumed;
        // issued before the wait: latency hidden behind the MMAs
        uint32_t gap[KTOK / 2];
#pragma unroll
        for (int j = 0; j < KTOK / 2; ++j) gap[j] = 0u;
        if (gattn_frame && valid) {
            const bf16* row = gattn_frame + (size_t)n * K;
            if ((K & 7) == 0) {
#pragma unroll
                for (int j = 0; j < KTOK / 8; ++j) {
                    if (8 * j < K) {
                        const uint4 v = *reinterpret_cast<const uint4*>(row + 8 * j);
                        gap[4 * j] = v.x; gap[4 * j + 1] = v.y; gap[4 * j + 2] = v.z; gap[4 * j + 3] = v.w;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < KTOK / 2; ++j) {
                    const uint32_t lo = (2 * j < K) ? (uint32_t)__bfloat16_as_ushort(row[2 * j]) : 0u;
                    const uint32_t hi = (2 * j + 1 < K) ? (uint32_t)__bfloat16_as_ushort(row[2 * j + 1]) : 0u;
                    gap[j] = lo | (hi << 16);
                }
            }
        }
        mbar_wait(&c.bars[B_SFULL4 + c.wg], ts.own_s & 1u);
        ++ts.own_s;
        fence_after_sync();
        UPH(55);
        float l[KTOK], gq[KTOK];
        tmem_ld16(scol, l); tmem_ld8(scol + 16, l + 16);
        tmem_ld16(gcol, gq); tmem_ld8(gcol + 16, gq + 16);
        tmem_wait_ld();
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
        float dot = 0.f;
#pragma unroll
        for (int s = 0; s < KTOK; ++s) {
            if (s < K) {
                l[s] *= scale;                                                           // P
                const __nv_bfloat162 gp = *reinterpret_cast<const __nv_bfloat162*>(&gap[s >> 1]);
                const float ga = (s & 1) ? __high2float(gp) : __low2float(gp);
                const float dp = (gq[s] - cv[s]) * cv[32 + s] + ga;                      // dP
                gq[s] = dp;
                dot = fmaf(l[s], dp, dot);
            }
        }
        UPH(56);
        if (nseq > 1u) { mbar_wait(&c.bars[B_AFREE4 + c.wg], ts.own_a & 1u); ++ts.own_a; }   // the second product of tile n - 2 has read this dL tile
        UPH(57);
#pragma unroll
        for (int s = 0; s < KTOK; s += 8) {
            uint32_t hv[4], lv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float dl2[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ss = s + 2 * e + q;
                    const bool ok = valid && ss < K;
                    const float pv = l[ss];
                    dl2[q] = ok ? pv * (gq[ss] - dot) : 0.f;                                 // dL
                    l[ss] = ok ? (pv + d.eps) * cv[32 + ss] : 0.f;                           // W
                    gq[ss] = dl2[q];
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(dl2[0], dl2[1]);
                const float2 f2 = __bfloat1622float2(h2);
                hv[e] = *reinterpret_cast<const uint32_t*>(&h2);
                lv[e] = pack_bf16x2(dl2[0] - f2.x, dl2[1] - f2.y);
            }
            const uint32_t ch = (uint32_t)(s >> 3);
            *reinterpret_cast<uint4*>(dlrow + ((ch ^ sw) << 4)) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<uint4*>(dlrow + (((4u + ch) ^ sw) << 4)) = make_uint4(lv[0], lv[1], lv[2], lv[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&c.bars[B_AREADY + g]);
        UPH(58);
        {
            // coefficients of the d_inputs kernel: this token's row of the iteration's K-major operand block,
            // [128 tokens][32 dL | 32 W] bf16, SWIZZLE_128B (savi_dx_umma.cu fetches the block with one bulk copy);
            // slots [24, 32) belong to nobody and rows past N (ragged last tile) to no token: written as zeros, the d_inputs
            // kernel also sums the block's columns over its rows on the tensor cores
            unsigned char* row = coef + (size_t)(tile0 + i) * d.I * 16384 + (size_t)c.o * 128;
#pragma unroll
            for (int s = 0; s < KTOK; s += 8) {
                const uint32_t ch = (uint32_t)(s >> 3);
                *reinterpret_cast<uint4*>(row + ((ch ^ sw) << 4)) =
                    make_uint4(pack_bf16x2(gq[s], gq[s + 1]), pack_bf16x2(gq[s + 2], gq[s + 3]), pack_bf16x2(gq[s + 4], gq[s + 5]), pack_bf16x2(gq[s + 6], gq[s + 7]));
                *reinterpret_cast<uint4*>(row + (((4u + ch) ^ sw) << 4)) =
                    make_uint4(pack_bf16x2(l[s], l[s + 1]), pack_bf16x2(l[s + 2], l[s + 3]), pack_bf16x2(l[s + 4], l[s + 5]), pack_bf16x2(l[s + 6], l[s + 7]));
            }
#pragma unroll
            for (uint32_t ch = KTOK / 8; ch < 4; ++ch) {
                *reinterpret_cast<uint4*>(row + ((ch ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(row + (((4u + ch) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        UPH(59);
    }
    ts.nseq += (uint32_t)ntile;
}

static __device__ __noinline__ void ln_bwd_reduce(const float2* scr, float2* red, int warp, int lane, int K) {
    bar_sync_compute();
#pragma unroll 1
    for (int k = warp; k < K; k += NCW) {
        const float4 a = *reinterpret_cast<const float4*>(scr + k * F + lane * 4);
        const float4 b = *reinterpret_cast<const float4*>(scr + k * F + lane * 4 + 2);
        const float s1 = warp_sum((a.x + a.z) + (b.x + b.z)), s2 = warp_sum((a.y + a.w) + (b.y + b.w));
        if (lane == 0) red[k] = make_float2(s1 * (1.0f / F), s2 * (1.0f / F));
    }
    bar_sync_compute();
}
// LayerNorm backward in the feature-per-thread layout.  dy: gradient of the LN output; x: its input (this thread's column);
// st: (mean, rstd) of each slot row.  Accumulates d gamma / d beta of this feature; returns dx.
__device__ __forceinline__ void ln_bwd(const Ctx& c, const float (&dy)[KH], const float (&x)[KH], const float2* st, float gamma,
                                       float& dgam, float& dbet, float (&dx)[KH], int scr_off) {
    float2* scr = reinterpret_cast<float2*>(c.sm + scr_off);                 // [K][128] (dz, dz * xhat)
    float2* red = reinterpret_cast<float2*>(c.sm + c.L.stats);               // [32] means of the two sums
    float xh[KH], dz[KH];
    float2 s_[KH];
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        if (kk < c.nk) {
            s_[kk] = st[c.k0 + kk];
            xh[kk] = (x[kk] - s_[kk].x) * s_[kk].y;
            dz[kk] = dy[kk] * gamma;
            dgam = fmaf(dy[kk], xh[kk], dgam); dbet += dy[kk];
            scr[(c.k0 + kk) * F + c.o] = make_float2(dz[kk], dz[kk] * xh[kk]);
        } else { xh[kk] = 0.f; dz[kk] = 0.f; s_[kk] = make_float2(0.f, 0.f); }
    }
    ln_bwd_reduce(scr, red, c.warp, c.lane, c.K);
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        if (kk < c.nk) { const float2 r = red[c.k0 + kk]; dx[kk] = s_[kk].y * (dz[kk] - r.x - xh[kk] * r.y); }
        else dx[kk] = 0.f;
    }
}

// Backward of the predictor's attention core in shared memory (transformer.py:34-47).  dQ is w.r.t. the UNSCALED projection.
constexpr int ATT_PT = 8;     // attention-matrix elements per thread: heads * K * K <= 8 * 512
__device__ __forceinline__ void mha_core_bwd(const Ctx& c, int H, float hscale, const float (&dO)[KH], const float (&Qv)[KH], const float (&Kv)[KH],
                                             const float (&Vv)[KH], const float (&attv)[ATT_PT], float (&dQ)[KH], float (&dK)[KH], float (&dV)[KH],
                                             const float* matt /* nullable: the forward's dropout mask of the probabilities [H][K][K] */,
                                             long long* dbg, long long& ph_last) {
    // layout as in the forward core (savi_fwd_umma.cu: mha_core): 16-byte aligned rows, attention rows padded to ka
    const int K = c.K, ld = MHA_LD, ka = mha_ka(K), dh = F / H;
    float* sO = reinterpret_cast<float*>(c.sm + c.L.opA);             // opA .. scratch are contiguous (5 x 16 KB + scratch)
    float* sQ = sO + K * ld; float* sK = sQ + K * ld; float* sV = sK + K * ld;
    float* sA = sV + K * ld;                                          // [H][query][ka key]   probabilities
    float* sD = sA + H * K * ka;                                      // [H][query][ka key]   d attention -> d logits
    float* sDT = sD + H * K * ka;                                     // [H][key][ka query]   d logits, transposed
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) {
        const int k = c.k0 + kk;
        if (kk < c.nk) { sO[k * ld + c.o] = dO[kk]; sQ[k * ld + c.o] = Qv[kk]; sK[k * ld + c.o] = Kv[kk]; sV[k * ld + c.o] = Vv[kk]; }
    }
    {   // element idx = tid + e * NCT of the [H*K][K] matrices -> padded row / column, without a division per element
        int q = c.tid / K, r = c.tid - q * K;
        const int dq = NCT / K, dr = NCT - dq * K;
#pragma unroll
        for (int e = 0; e < ATT_PT; ++e) {
            if (q < H * K) sA[q * ka + r] = attv[e];
            q += dq; r += dr;
            if (r >= K) { r -= K; ++q; }
        }
    }
    bar_sync_compute();
    UPH(45);
    // d attention = dO_h . V_h^T: 4 threads per (head, query) row, <= MHA_JP keys each with independent accumulators
#pragma unroll 1
    for (int row = c.tid >> 2; row < H * K; row += NCT / 4) {
        const int jq = c.tid & 3, JP = (K + 3) >> 2, j0 = jq * JP;
        const int h = row / K, i = row - h * K;
        const float* a = sO + i * ld + h * dh;
        const float* b = sV + h * dh;
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
        // O = (att . m) V: d att = (dO V^T) . m
        for (int jj = 0; jj < MHA_JP; ++jj) if (jj < JP && j0 + jj < K) sD[row * ka + j0 + jj] = matt ? acc[jj] * matt[row * K + j0 + jj] : acc[jj];
    }
    bar_sync_compute();
    UPH(46);
    // softmax backward: one warp per (head, query) row, lane = key; three rows in flight per warp
#pragma unroll 1
    for (int row0 = c.warp; row0 < H * K; row0 += 3 * NCW) {
        float av[3], dav[3], dot[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int row = row0 + r * NCW;
            const bool on = row < H * K && c.lane < K;
            av[r] = on ? sA[row * ka + c.lane] : 0.f; dav[r] = on ? sD[row * ka + c.lane] : 0.f;
            dot[r] = av[r] * dav[r];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < 3; ++r) dot[r] += __shfl_xor_sync(0xffffffffu, dot[r], o);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int row = row0 + r * NCW;
            if (row < H * K && c.lane < K) {
                const float dl = av[r] * (dav[r] - dot[r]);
                sD[row * ka + c.lane] = dl; sDT[((row / K) * K + c.lane) * ka + row % K] = dl;
                if (matt) sA[row * ka + c.lane] = av[r] * matt[row * K + c.lane];     // from here on only d V = (att . m)^T dO reads the probabilities
            }
        }
    }
    bar_sync_compute();
    UPH(47);
    const int hb = (c.o / dh) * K * ka + c.k0;
    const float* dlt = sDT + hb;                                      // row j: d logits [i][j] over this warpgroup's 8 slots i
    const float* dlg = sD + hb;                                       // row j: d logits [j][i]
    const float* at = sA + hb;                                        // row j: attention [j][i]
    float sq[KH], sk[KH], sv[KH];
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) { sq[kk] = 0.f; sk[kk] = 0.f; sv[kk] = 0.f; }
    if (c.nk > 0) {
#pragma unroll 2
        for (int j = 0; j < K; ++j) {
            const float kj = sK[j * ld + c.o], qj = sQ[j * ld + c.o], oj = sO[j * ld + c.o];
            float t[KH], u[KH], w[KH];                                // warp-uniform 16-byte reads
            *reinterpret_cast<float4*>(t) = ld4(dlt + j * ka); *reinterpret_cast<float4*>(t + 4) = ld4(dlt + j * ka + 4);
            *reinterpret_cast<float4*>(u) = ld4(dlg + j * ka); *reinterpret_cast<float4*>(u + 4) = ld4(dlg + j * ka + 4);
            *reinterpret_cast<float4*>(w) = ld4(at + j * ka); *reinterpret_cast<float4*>(w + 4) = ld4(at + j * ka + 4);
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) {
                sq[kk] = fmaf(t[kk], kj, sq[kk]);
                sk[kk] = fmaf(u[kk], qj, sk[kk]);
                sv[kk] = fmaf(w[kk], oj, sv[kk]);
            }
        }
    }
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) { dQ[kk] = sq[kk] * hscale; dK[kk] = sk[kk]; dV[kk] = sv[kk]; }
    bar_sync_compute();
    UPH(48);
}

__device__ __forceinline__ float sum8(const Ctx& c, const float (&v)[KH]) {
    float s = 0.f;
#pragma unroll
    for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) s += v[kk];
    return s;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// KFIX / CNFIX / HFIX: slots, CTAs per clip and predictor heads when the launch site knows them at compile time (0: read them from
// the shape).  The MOVi configurations get their own instances: with K fixed, the per-slot predicates of the token threads and
// the row / K index arithmetic of the predictor fold away; everything else runs the generic <0, 0, 0> instance.
template <int KFIX, int CNFIX, int HFIX>
__global__ void __launch_bounds__(NTHREADS, 1) savi_bwd_umma_kernel(const __grid_constant__ BwdUArgs ua) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const BwdArgs& a = ua.a;
    const Dims& d = a.d;
    const ParamOff& po = a.po;
    unsigned char* sm = smem_raw;
    if ((smem_u32(sm) & 1023u) != 0u) __trap();                // SWIZZLE_128B operands need a 1024-byte aligned base
    // (warp index through a lane-0 broadcast: ptxas then treats it, and every role branch on it, as warp-uniform)
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int CN = CNFIX ? CNFIX : d.CN, b = blockIdx.x / CN, rank = blockIdx.x % CN;
    const int K = KFIX ? KFIX : d.K, B = d.B, KP = KFIX ? ((KFIX + 3) & ~3) : d.KP;
    const int HEADS = HFIX ? HFIX : d.heads;
    const Smem L = plan_smem(K, CN, true);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const bool lead = (rank == 0);
    // both CTAs of a pair carry bit-identical slot-state gradients: the staged records are split between them
    const bool svA = lead, svB = (rank == CN - 1);
    const int per = (d.NTILE + CN - 1) / CN;
    const int tile0 = min(d.NTILE, rank * per), ntile = min(d.NTILE, tile0 + per) - tile0;
    const float* P = a.packed;
    float* G = a.grad_params;
    const float* fb = reinterpret_cast<const float*>(a.saved + a.sl.fbase);
    const unsigned char* ximg = a.saved + a.sl.ximg;
    float* W = a.ws;
    unsigned char* coef_base = reinterpret_cast<unsigned char*>(W) + a.wl.coef;
    const int scr2 = L.opC;                                   // float2 [K][128] LayerNorm-backward scratch: opC | aw0 (= X2 | X3, idle at every ln_bwd call)
    float* cv = reinterpret_cast<float*>(sm + L.stats + 256);  // [64]: c | 1/S

    // ---- one-time setup ----
    for (int i = tid * 16; i < L.bars; i += NTHREADS * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < L.nst; ++s) { mbar_init(&bars[B_FULL + s], 1); mbar_init(&bars[B_EMPTY + s], 1); }
        mbar_init(&bars[B_OPND], NCW); mbar_init(&bars[B_ACC], 1); mbar_init(&bars[B_TOK], 1); mbar_init(&bars[B_HP], 1);
        for (int g = 0; g < 2; ++g) {
            mbar_init(&bars[B_AREADY + g], 4);                                         // one warpgroup (4 warps) per tile
            mbar_init(&bars[B_INBOX + g], NCW);
        }
        for (int f = 0; f < 4; ++f) { mbar_init(&bars[B_FACC + f], 1); mbar_init(&bars[B_FOPND + f], NCW); mbar_init(&bars[B_SFULL4 + f], 1); mbar_init(&bars[B_AFREE4 + f], 1); mbar_init(&bars[B_SFREE4 + f], 4); }
        mbar_init_fence();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TB_COLS);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (CN > 1) { cluster_arrive(); cluster_wait(); }
    const uint32_t tb = *tmem_slot;
    // every CTA of this grid is resident from here on: the dependent d_inputs grid may take the idle SMs (it waits on the frame flags)
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (ua.trace && tid == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_)); atomicMin(ua.trace, t_); }

    Ring ring;
    ring.base = sm + L.ring; ring.full = &bars[B_FULL]; ring.empty = &bars[B_EMPTY]; ring.nst = L.nst; ring.stage = 0; ring.phase = 0;
    ring.wait_cycles = nullptr;

    if (warp == W_PROD) {
        // =====================================================================================
        // producer (mirrors the issuer's consumption order)
        // =====================================================================================
        if (lane == 0) {
            const unsigned char* Wi = ua.wimg;
            const WImg& wi = ua.wi;
            for (int t = d.T - 1; t >= 0; --t) {
                if (t < d.T - 1) {
                    for (int j = d.blocks - 1; j >= 0; --j) {
                        const WImgBlock& wb = wi.blkT[j];
                        prod_blocks(ring, Wi + wb.f2, 8 * NBW);                       // ffn.2^T: 4 row tiles x K = 128
                        prod_blocks(ring, Wi + wb.f1 + 2 * NBW * BLK, 6 * NBW);             // ffn.0^T: contraction chunks 1, 2, 3, then 0
                        prod_blocks(ring, Wi + wb.f1, 2 * NBW);
                        prod_blocks(ring, Wi + wb.po, 2 * NBW);
                        prod_blocks(ring, Wi + wb.pq, 2 * NBW); prod_blocks(ring, Wi + wb.pk, 2 * NBW); prod_blocks(ring, Wi + wb.pv, 2 * NBW);
                    }
                }
                const unsigned char* xf = ximg + ((size_t)(b * d.T + t) * d.NTILE + tile0) * 2 * BLK;
                for (int it = d.I - 1; it >= 0; --it) {
                    // (this otherwise idle thread also warms L2: kept out of the compute warps' instruction stream)
                    const int64_t s = (int64_t)t * d.I + it;
                    float* fbw = const_cast<float*>(fb);
                    if (s > 0) {
                        // The saved records are read in reverse order of their creation and do not fit L2 as a whole:
                        // pull the previous step's rows (the next ones this kernel needs) into L2 while this step runs.
                        const int64_t sp = s - 1;
                        const uint32_t rowb = (uint32_t)K * F * 4u;
                        prefetch_l2(frow(fbw, a.sl.r, sp, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.z, sp, b, B, K, F), rowb);
                        prefetch_l2(frow(fbw, a.sl.n, sp, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.ghn, sp, b, B, K, F), rowb);
                        prefetch_l2(frow(fbw, a.sl.hp, sp, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.qk, sp, b, B, K, F), rowb);
                        prefetch_l2(frow(fbw, a.sl.ux, sp, b, B, K, F), rowb);
                        const int itp = (it > 0) ? it - 1 : d.I - 1, tp = (it > 0) ? t : t - 1;
                        if (itp < d.I - 1) {
                            const int64_t smp = (int64_t)tp * (d.I - 1) + itp;
                            prefetch_l2(frow(fbw, a.sl.a, smp, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.hg, smp, b, B, K, F), rowb);
                        }
                        if (it == 0 && t > 0) {                                                // the predictor link of frame t-1 comes next
                            prefetch_l2(fb + a.sl.px0 + ((size_t)(t - 1) * B + b) * K * F, rowb);
                            for (int j = 0; j < d.blocks; ++j) {
                                const int64_t f = (int64_t)j * (d.T - 1) + (t - 1);
                                prefetch_l2(frow(fbw, a.sl.px2, f, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.px1, f, b, B, K, F), rowb);
                                prefetch_l2(frow(fbw, a.sl.pq, f, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.pk, f, b, B, K, F), rowb);
                                prefetch_l2(frow(fbw, a.sl.pv, f, b, B, K, F), rowb); prefetch_l2(frow(fbw, a.sl.pf, f, b, B, K, 4 * F), 4 * rowb);
                                const int natt = HEADS * K * K;                               // saved attention matrices of the block
                                if ((natt & 3) == 0) prefetch_l2(fb + a.sl.patt + (f * B + b) * (int64_t)natt, (uint32_t)natt * 4u);
                            }
                        }
                    }
                    if (it < d.I - 1) { prod_blocks(ring, Wi + wi.w2T, 2 * NBW); prod_blocks(ring, Wi + wi.w1T, 2 * NBW); }
                    prod_blocks(ring, Wi + wi.wgT, 6 * NBW);
                    prod_blocks(ring, Wi + wi.whhT, 6 * NBW);
                    prod_token_blocks(ring, xf, 2 * ntile, 16);
                    prod_blocks(ring, Wi + wi.wqkT, 2 * NBW);
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
            const uint32_t X0 = smem_u32(sm + L.opA), X1 = smem_u32(sm + L.opB), X2 = smem_u32(sm + L.opC), X3 = smem_u32(sm + L.aw0);
            auto wait_opnd = [&]() { mbar_wait(&bars[B_OPND], ph_opnd); ph_opnd ^= 1u; fence_after_sync(); };
            for (int t = d.T - 1; t >= 0; --t) {
                if (t < d.T - 1) {
                    for (int j = d.blocks - 1; j >= 0; --j) {
                        wait_opnd();                                                       // d x2 in X0
                        for (int f = 0; f < 4; ++f) { issue_linear(ring, el, X0, tb + TB_F0 + 64 * f, 1, 2, false); if (el) mma_commit(&bars[B_FACC + f]); }   // d f = ffn.2^T d x2
                        // d l2 = ffn.0^T d f: chunks 1, 2, 3 (X1..X3) then chunk 0 (X0, rewritten after all four tiles above completed)
                        for (int q = 0; q < 4; ++q) {
                            const int f = (q + 1) & 3;
                            mbar_wait(&bars[B_FOPND + f], pcall & 1u); fence_after_sync();
                            issue_linear(ring, el, f == 0 ? X0 : f == 1 ? X1 : f == 2 ? X2 : X3, tb + TB_A, 1, 2, q > 0);
                        }
                        if (el) mma_commit(&bars[B_ACC]);
                        ++pcall;
                        wait_opnd();                                                       // d x1 in X0
                        issue_linear(ring, el, X0, tb + TB_B, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);        // dO = proj_o^T d x1
                        wait_opnd();                                                       // dQ, dK, dV in X1, X2, X3
                        issue_linear(ring, el, X1, tb + TB_A, 1, 2, false);
                        issue_linear(ring, el, X2, tb + TB_A, 1, 2, true);
                        issue_linear(ring, el, X3, tb + TB_A, 1, 2, true); if (el) mma_commit(&bars[B_ACC]);         // dy
                    }
                }
                for (int it = d.I - 1; it >= 0; --it) {
                    if (it < d.I - 1) {
                        wait_opnd();                                                       // dh in X0
                        issue_linear(ring, el, X0, tb + TB_A, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);        // d a = mlp.2^T dh
                        wait_opnd();                                                       // d a in X1
                        issue_linear(ring, el, X1, tb + TB_B, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);        // d m = mlp.0^T d a
                    }
                    wait_opnd();                                                           // dr, dz, dn, dn*r in X0..X3
                    issue_linear(ring, el, X0, tb + TB_B, 1, 2, false);
                    issue_linear(ring, el, X1, tb + TB_B, 1, 2, true);
                    issue_linear(ring, el, X2, tb + TB_B, 1, 2, true); if (el) mma_commit(&bars[B_ACC]);             // dUx = (W_ih W_v)^T dgi  (folded: dU is never formed)
                    issue_linear(ring, el, X0, tb + TB_HP, 1, 2, false);
                    issue_linear(ring, el, X1, tb + TB_HP, 1, 2, true);
                    issue_linear(ring, el, X3, tb + TB_HP, 1, 2, true); if (el) mma_commit(&bars[B_HP]);             // W_hh^T dgh (consumed at the end of the step)
                    wait_opnd();                                                           // [qk | dUx] hi halves in X0, lo halves in X1
                    issue_token_pass_bwd(ring, el, bars, tb, ntile, ts, X0, X1, X2, X3);
                    wait_opnd();                                                           // d qk in X0
                    issue_linear(ring, el, X0, tb + TB_B, 1, 2, false); if (el) mma_commit(&bars[B_ACC]);            // d s~ = wqk^T d qk  (folded: dq is never formed)
                }
            }
        }
    } else {
        // =====================================================================================
        // compute threads
        // =====================================================================================
        Ctx c;
        ctx_init(c, tid, K, sm, L, tb, bars);
        const int o = c.o;
        TokState ts = {0, 0, 0};
        long long* sdbg = reinterpret_cast<long long*>(sm + L.bars + NBAR * 8 + 16);
        if (a.dbg && blockIdx.x == 0 && tid == 0) for (int i = 0; i < 64; ++i) sdbg[i] = 0;
#ifdef SAVI_PHASE_PROFILE
        long long* dbg = (a.dbg && blockIdx.x == 0 && tid == 0) ? sdbg : nullptr;
#else
        long long* const dbg = nullptr;      // production build: the ~30 UPH probes fold away (tools/phase_times.py builds its own library)
#endif
        long long ph_last = clock64();
        const float g_s = P[po.ln_s_w + o], g_m = P[po.ln_m_w + o];
        const float hscale = 1.0f / sqrtf((float)(F / HEADS));
        // per-feature parameter-gradient accumulators over all steps of this clip (added to the flat buffer at the end)
        float a_gs = 0.f, a_bs = 0.f, a_gm = 0.f, a_bm = 0.f, a_b1 = 0.f, a_b2 = 0.f;
        float a_dr = 0.f, a_dz = 0.f, a_dn = 0.f, a_dnr = 0.f;
        float dh[KH];
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) dh[kk] = 0.f;
        uint32_t step = 0, pcall = 0, hpcall = 0;
        for (int t = d.T - 1; t >= 0; --t) {
            if (t < d.T - 1) {
                // ---- predictor backward (transformer.py:106-114, 70-86, 22-49) ----
                const float* px0 = fb + a.sl.px0 + ((size_t)t * B + b) * K * F;
                const float* x_last = d.blocks > 0 ? frow(const_cast<float*>(fb), a.sl.px2, (int64_t)(d.blocks - 1) * (d.T - 1) + t, b, B, K, F) : px0;
                float xin[KH], t0[KH];
                load_field(c, x_last, F, o, xin);
                slot_stats(c, xin, d.ln_eps);
                {
                    float dg = 0.f, dbt = 0.f;
                    ln_bwd(c, dh, xin, reinterpret_cast<const float2*>(sm + L.stats) + 0, P[po.lnf_w + o], dg, dbt, t0, scr2);
                    if (lead) { atomicAdd(G + po.lnf_w + o, dg); atomicAdd(G + po.lnf_b + o, dbt); }
                }
                for (int j = d.blocks - 1; j >= 0; --j) {
                    const int64_t f = (int64_t)j * (d.T - 1) + t;
                    const BlockOff& bo = po.blk[j];
                    float* fbw = const_cast<float*>(fb);
                    const float* p_q = frow(fbw, a.sl.pq, f, b, B, K, F);
                    const float* p_k = frow(fbw, a.sl.pk, f, b, B, K, F);
                    const float* p_v = frow(fbw, a.sl.pv, f, b, B, K, F);
                    const float* p_x1 = frow(fbw, a.sl.px1, f, b, B, K, F);
                    const float* p_f = frow(fbw, a.sl.pf, f, b, B, K, 4 * F);
                    const float* p_att = fb + a.sl.patt + (f * B + b) * ((int64_t)HEADS * K * K);
                    UPH(37);
                    // training-mode dropout masks of this block evaluation (nullptr otherwise): savi_args.h, DropLayout
                    const DropLayout dlay = savi_dropout_layout(d);
                    const float* m_att = a.drop ? a.drop + dlay.att + (f * B + b) * ((int64_t)HEADS * K * K) : nullptr;
                    const float* m_out = a.drop ? a.drop + dlay.out + (f * B + b) * ((int64_t)K * F) : nullptr;
                    const float* m_ffn = a.drop ? a.drop + dlay.ffn + (f * B + b) * ((int64_t)K * F) : nullptr;
                    // t0 = d x2; the FFN branch sees d x2 . m_ffn (staged for d W2 / d b2, pushed through ffn.2^T), the residual keeps t0
                    float tm[KH];
                    if (m_ffn) {
                        float mf[KH];
                        load_field(c, m_ffn, F, o, mf);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) tm[kk] = t0[kk] * mf[kk];
                    } else {
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) tm[kk] = t0[kk];
                    }
                    if (svB) save_field(c, frow(W, a.wl.pdx2, f, b, B, K, F), F, o, tm);
                    if (lead) atomicAdd(G + bo.f2b + o, sum8(c, tm));
                    write_operand(c, xop(L, 0), tm);
                    signal_operand(c);
                    // d f = (ffn.2^T d x2) masked by relu; chunks 1..3 go to X1..X3 at once, chunk 0 to X0 after all four tiles are done
                    float df0[KH];
#pragma unroll
                    for (int ff = 0; ff < 4; ++ff) {
                        float v[KH], fm[KH];
                        load_field(c, p_f, 4 * F, ff * F + o, fm);                         // issued before the wait: latency hidden behind the MMAs
                        mbar_wait(&bars[B_FACC + ff], pcall & 1u); fence_after_sync();
                        load_acc(c, TB_F0 + 64 * ff, v);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) v[kk] = (fm[kk] > 0.f) ? v[kk] : 0.f;
                        if ((ff & 1) ? svB : svA) { save_field(c, frow(W, a.wl.pdf, f, b, B, K, 4 * F), 4 * F, ff * F + o, v); atomicAdd(G + bo.f1b + ff * F + o, sum8(c, v)); }
                        if (ff == 0) {
#pragma unroll
                            for (int kk = 0; kk < KH; ++kk) df0[kk] = v[kk];
                        } else {
                            write_operand(c, xop(L, ff), v);
                            signal_operand(c, B_FOPND + ff);
                        }
                    }
                    write_operand(c, xop(L, 0), df0);
                    signal_operand(c, B_FOPND + 0);
                    ++pcall;
                    UPH(38);
                    float t1[KH], dx1[KH], x1v[KH];
                    load_field(c, p_x1, F, o, x1v);
                    wait_acc(c); load_acc(c, TB_A, t1);                                    // d l2
                    UPH(39);
                    slot_stats(c, x1v, d.ln_eps);
                    {
                        float dg = 0.f, dbt = 0.f;
                        ln_bwd(c, t1, x1v, reinterpret_cast<const float2*>(sm + L.stats), P[bo.ln2_w + o], dg, dbt, dx1, scr2);
                        if (lead) { atomicAdd(G + bo.ln2_w + o, dg); atomicAdd(G + bo.ln2_b + o, dbt); }
                    }
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) dx1[kk] += t0[kk];
                    {   // the proj_o branch sees d x1 . m_out (staged for d W_o, pushed through proj_o^T); the residual below keeps d x1
                        float dm[KH];
                        if (m_out) {
                            float mo[KH];
                            load_field(c, m_out, F, o, mo);
#pragma unroll
                            for (int kk = 0; kk < KH; ++kk) dm[kk] = dx1[kk] * mo[kk];
                        } else {
#pragma unroll
                            for (int kk = 0; kk < KH; ++kk) dm[kk] = dx1[kk];
                        }
                        if (svB) save_field(c, frow(W, a.wl.pdx1, f, b, B, K, F), F, o, dm);
                        write_operand(c, xop(L, 0), dm);
                    }
                    signal_operand(c);
                    UPH(40);
                    float dO[KH], dQ[KH], dK[KH], dV[KH];
                    {
                        float qv[KH], kv[KH], vv[KH], attv[ATT_PT];
                        load_field(c, p_q, F, o, qv); load_field(c, p_k, F, o, kv); load_field(c, p_v, F, o, vv);
#pragma unroll
                        for (int e = 0; e < ATT_PT; ++e) { const int idx = tid + e * NCT; attv[e] = (idx < HEADS * K * K) ? p_att[idx] : 0.f; }
                        wait_acc(c); load_acc(c, TB_B, dO);
                        UPH(41);
                        mha_core_bwd(c, HEADS, hscale, dO, qv, kv, vv, attv, dQ, dK, dV, m_att, dbg, ph_last);
                        UPH(42);
                    }
                    if (svA) {
                        save_field(c, frow(W, a.wl.pdq, f, b, B, K, F), F, o, dQ);
                        save_field(c, frow(W, a.wl.pdv, f, b, B, K, F), F, o, dV);
                    }
                    if (svB) save_field(c, frow(W, a.wl.pdk, f, b, B, K, F), F, o, dK);
                    write_operand(c, xop(L, 1), dQ); write_operand(c, xop(L, 2), dK); write_operand(c, xop(L, 3), dV);
                    signal_operand(c);
                    const float* x_in = (j == 0) ? px0 : frow(fbw, a.sl.px2, (int64_t)(j - 1) * (d.T - 1) + t, b, B, K, F);
                    load_field(c, x_in, F, o, xin);
                    UPH(43);
                    wait_acc(c); load_acc(c, TB_A, t1);                                    // dy
                    UPH(44);
                    if (j == 0) {
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) t1[kk] += dx1[kk];                 // residual taken from the normalised input
                    }
                    slot_stats(c, xin, d.ln_eps);
                    {
                        float dg = 0.f, dbt = 0.f;
                        ln_bwd(c, t1, xin, reinterpret_cast<const float2*>(sm + L.stats), P[bo.ln1_w + o], dg, dbt, t0, scr2);
                        if (lead) { atomicAdd(G + bo.ln1_w + o, dg); atomicAdd(G + bo.ln1_b + o, dbt); }
                    }
                    if (j != 0) {
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) t0[kk] += dx1[kk];
                    }
                }
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) dh[kk] = t0[kk];
            }
            {
                float gsl[KH];
                load_field(c, a.grad_slots + ((size_t)b * d.T + t) * K * F, F, o, gsl);
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) dh[kk] += gsl[kk];
            }
            for (int it = d.I - 1; it >= 0; --it, ++step) {
                const int64_t s = (int64_t)t * d.I + it;
                float* fbw = const_cast<float*>(fb);
                UPH(25);
                if (it < d.I - 1) {
                    // ---- residual MLP backward (steve.py:92-93) ----
                    const int64_t smi = (int64_t)t * (d.I - 1) + it;
                    float am[KH], hg[KH], v[KH];
                    load_field(c, frow(fbw, a.sl.a, smi, b, B, K, F), F, o, am);
                    load_field(c, frow(fbw, a.sl.hg, smi, b, B, K, F), F, o, hg);
                    // (staged records are stored after the operand hand-over of their phase, off the serial chain)
                    write_operand(c, xop(L, 0), dh);
                    signal_operand(c);
                    if (svA) save_field(c, frow(W, a.wl.dhm, smi, b, B, K, F), F, o, dh);
                    a_b2 += sum8(c, dh);
                    wait_acc(c); load_acc(c, TB_A, v);
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) v[kk] = (am[kk] > 0.f) ? v[kk] : 0.f;    // d a
                    write_operand(c, xop(L, 1), v);
                    signal_operand(c);
                    if (svB) save_field(c, frow(W, a.wl.da, smi, b, B, K, F), F, o, v);
                    a_b1 += sum8(c, v);
                    wait_acc(c); load_acc(c, TB_B, v);                                     // d m
                    float dxm[KH];
                    ln_bwd(c, v, hg, reinterpret_cast<const float2*>(fb + a.sl.lnm) + (smi * B + b) * K, g_m, a_gm, a_bm, dxm, scr2);
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) dh[kk] += dxm[kk];
                }
                UPH(26);
                // ---- GRUCell backward (steve.py:87-89) ----
                {
                    float r_[KH], z_[KH], n_[KH], hn[KH], hp[KH];
                    load_field(c, frow(fbw, a.sl.r, s, b, B, K, F), F, o, r_);
                    load_field(c, frow(fbw, a.sl.z, s, b, B, K, F), F, o, z_);
                    load_field(c, frow(fbw, a.sl.n, s, b, B, K, F), F, o, n_);
                    load_field(c, frow(fbw, a.sl.ghn, s, b, B, K, F), F, o, hn);
                    load_field(c, frow(fbw, a.sl.hp, s, b, B, K, F), F, o, hp);
                    float dr[KH], dz[KH], dn[KH], dnr[KH];
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) {
                        const float g = dh[kk];
                        dn[kk] = g * (1.0f - z_[kk]) * (1.0f - n_[kk] * n_[kk]);
                        dz[kk] = g * (hp[kk] - n_[kk]) * z_[kk] * (1.0f - z_[kk]);
                        dr[kk] = dn[kk] * hn[kk] * r_[kk] * (1.0f - r_[kk]);
                        dnr[kk] = dn[kk] * r_[kk];
                        dh[kk] = g * z_[kk];
                    }
                    write_operand(c, xop(L, 0), dr); write_operand(c, xop(L, 1), dz);
                    write_operand(c, xop(L, 2), dn); write_operand(c, xop(L, 3), dnr);
                    signal_operand(c);
                    if (svA) {
                        float* gi = frow(W, a.wl.dgi, s, b, B, K, 3 * F);
                        save_field(c, gi, 3 * F, o, dr); save_field(c, gi, 3 * F, F + o, dz); save_field(c, gi, 3 * F, 2 * F + o, dn);
                    }
                    if (svB) {
                        float* gh = frow(W, a.wl.dgh, s, b, B, K, 3 * F);
                        save_field(c, gh, 3 * F, o, dr); save_field(c, gh, 3 * F, F + o, dz); save_field(c, gh, 3 * F, 2 * F + o, dnr);
                    }
                    a_dr += sum8(c, dr); a_dz += sum8(c, dz); a_dn += sum8(c, dn); a_dnr += sum8(c, dnr);
                }
                UPH(27);
                float v[KH], dux[KH];
                // operands of the token pass that do not depend on dUx: qk (saved), Ux, 1/S (loads issued before the wait)
                float qk[KH], ux[KH];
                load_field(c, frow(fbw, a.sl.qk, s, b, B, K, F), F, o, qk);
                load_field(c, frow(fbw, a.sl.ux, s, b, B, K, F), F, o, ux);
                const float ssv = (tid < K) ? fb[a.sl.ssum + (s * B + b) * KP + tid] : 1.f;
                if (tid < 32) cv[32 + tid] = (tid < K) ? 1.0f / ssv : 0.f;
                UPH(28);
                wait_acc(c); load_acc(c, TB_B, dux);                                       // dUx = dgi (W_ih W_v)
                UPH(29);
                // c[k] = <dUx[k,:], Ux[k,:]>: reduce over the 128 features through the scratch tile
                {
                    float* scr = reinterpret_cast<float*>(sm + L.scratch);
#pragma unroll
                    for (int kk = 0; kk < KH; ++kk) if (kk < c.nk) scr[(c.k0 + kk) * F + o] = dux[kk] * ux[kk];
                    mbar_wait(&bars[B_HP], hpcall & 1u);                                   // W_hh^T product done: X0, X1, X3 may be rewritten
                    ++hpcall;
                    write_operand_halves(c, xop(L, 0), xop(L, 1), qk, dux);      // X0 = [qk_hi | dUx_hi], X1 = [qk_lo | dUx_lo]
                    bar_sync_compute();
#pragma unroll 1
                    for (int k = c.warp; k < K; k += NCW) {
                        const float4 x = ld4(scr + k * F + lane * 4);
                        const float sdot = warp_sum((x.x + x.y) + (x.z + x.w));
                        if (lane == 0) cv[k] = sdot;
                    }
                    bar_sync_compute();
                }
                signal_operand(c);
                if (svB) save_field(c, frow(W, a.wl.duxs, s, b, B, K, F), F, o, dux);
                UPH(30);
                // ---- attention step backward over the token tiles ----
                const bf16* ga = (a.grad_attn && it == d.I - 1) ? reinterpret_cast<const bf16*>(a.grad_attn) + ((size_t)b * d.T + t) * d.N * K : nullptr;
                unsigned char* coef = coef_base + (((size_t)b * d.T + t) * d.NTILE * d.I + it) * 16384;
                softmax_bwd_tiles(c, d, ntile, tile0, ga, coef, cv, xop(L, 2), xop(L, 3), ts, dbg);
                UPH(31);
                mbar_wait(&bars[B_TOK], step & 1u);
                fence_after_sync();
                UPH(32);
                float dqk[KH];
                load_acc(c, TB_DQK, dqk);
                if (CN > 1) {
                    // inbox layout [feature o][KR8 slots]: two 16-byte remote stores per thread (as in the forward)
                    float* ib = reinterpret_cast<float*>(sm + L.inbox);
                    const int KR8 = (K + 7) & ~7;
                    const uint32_t peer = rank ^ 1u;
                    // single inbox: the peer must have consumed what this CTA sent in the previous step (it says so on OUR barrier)
                    if (step > 0) mbar_wait_cluster(&bars[B_INBOX + 1], (step - 1) & 1u);
                    if (c.nk > 0) {
                        const uint32_t rb = map_to_rank(ib, peer) + (uint32_t)(o * KR8 + c.k0) * 4u;
                        st_cluster_f4(rb, make_float4(dqk[0], dqk[1], dqk[2], dqk[3]));
                        st_cluster_f4(rb + 16u, make_float4(dqk[4], dqk[5], dqk[6], dqk[7]));
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(map_to_rank(&bars[B_INBOX], peer));
                    mbar_wait_cluster(&bars[B_INBOX], step & 1u);
                    if (c.nk > 0) {
                        float pn[KH];
                        *reinterpret_cast<float4*>(pn) = ld4(ib + o * KR8 + c.k0); *reinterpret_cast<float4*>(pn + 4) = ld4(ib + o * KR8 + c.k0 + 4);
#pragma unroll
                        for (int kk = 0; kk < KH; ++kk) dqk[kk] = lead ? dqk[kk] + pn[kk] : pn[kk] + dqk[kk];
                        loaded_before_release(pn[0]); loaded_before_release(pn[4]);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(map_to_rank(&bars[B_INBOX + 1], peer));     // consumed: the peer may send again
                }
                UPH(33);
                write_operand(c, xop(L, 0), dqk);
                signal_operand(c);
                if (svA) save_field(c, frow(W, a.wl.dqk, s, b, B, K, F), F, o, dqk);
                UPH(34);
                float hp[KH], dst[KH], hpv[KH];
                load_field(c, frow(fbw, a.sl.hp, s, b, B, K, F), F, o, hp);
                load_acc(c, TB_HP, hpv);                                                   // W_hh^T dgh (B_HP was observed above)
                wait_acc(c); load_acc(c, TB_B, v);                                         // d s~
                UPH(35);
                ln_bwd(c, v, hp, reinterpret_cast<const float2*>(fb + a.sl.lns) + (s * B + b) * K, g_s, a_gs, a_bs, dst, scr2);
#pragma unroll
                for (int kk = 0; kk < KH; ++kk) dh[kk] += hpv[kk] + dst[kk];
                UPH(36);
            }
            // frame t is fully staged (coefficient blocks, d(Ux) rows): release it to the d_inputs kernel, which runs
            // concurrently on the SMs this grid leaves idle (programmatic dependent launch, savi_dx_umma.cu)
            bar_sync_compute();
            if (tid == 0) {
                __threadfence();
                atomicAdd(reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(a.ws) + a.wl.flags) + b * d.T + t, 1);
            }
        }
        // ---- slot initialisation backward (steve.py:56-57) and the accumulated vector-parameter gradients ----
        if (lead) {
            const float es = expf(P[po.slot_log_sigma + o]);
            float nz[KH];
            load_field(c, a.noise + (size_t)b * K * F, F, o, nz);
            float smu = 0.f, sls = 0.f;
#pragma unroll
            for (int kk = 0; kk < KH; ++kk) {
                if (kk < c.nk) {
                    smu += dh[kk]; sls = fmaf(dh[kk] * es, nz[kk], sls);
                    if (a.grad_noise) a.grad_noise[((size_t)b * K + c.k0 + kk) * F + o] = dh[kk] * es;
                }
            }
            if (c.nk > 0) {
                atomicAdd(G + po.slot_mu + o, smu); atomicAdd(G + po.slot_log_sigma + o, sls);
                atomicAdd(G + po.ln_s_w + o, a_gs); atomicAdd(G + po.ln_s_b + o, a_bs);
                atomicAdd(G + po.ln_m_w + o, a_gm); atomicAdd(G + po.ln_m_b + o, a_bm);
                atomicAdd(G + po.b1 + o, a_b1); atomicAdd(G + po.b2 + o, a_b2);
                atomicAdd(G + po.bih + o, a_dr); atomicAdd(G + po.bih + F + o, a_dz); atomicAdd(G + po.bih + 2 * F + o, a_dn);
                atomicAdd(G + po.bhh + o, a_dr); atomicAdd(G + po.bhh + F + o, a_dz); atomicAdd(G + po.bhh + 2 * F + o, a_dnr);
            }
        }
        if (dbg) for (int i = 0; i < 60; ++i) if (sdbg[i]) a.dbg[i] += sdbg[i];
    }
    // ---- teardown ----
    __syncwarp();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {     // every record of this CTA is written: count it for the weight-gradient kernel (a programmatic dependent too)
        __threadfence();
        atomicAdd(reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(a.ws) + a.wl.flags) + d.B * d.T, 1);
    }
    if (ua.trace && tid == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_)); atomicMax(ua.trace + 1, t_); }
    if (CN > 1) { cluster_arrive(); cluster_wait(); }
    if (warp == W_MMA) tmem_dealloc(tb, TB_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int savi_bwd_umma_smem_bytes(const Dims& d) {
    return plan_smem(d.K, d.CN, true).total;
}

cudaError_t savi_launch_bwd_umma(const BwdArgs& a, const unsigned char* wimg, const WImg& wi, cudaStream_t st) {
    BwdUArgs ua;
    ua.a = a; ua.wimg = wimg; ua.wi = wi;
    ua.trace = (a.dbg && savi_options().dx_trace) ? a.dbg + 64 + 3 * 4096 : nullptr;      // [0] = min start (preset to LLONG_MAX), [1] = max end
    if (ua.trace) ua.a.dbg = nullptr;                                                         // no phase counters in a trace run
    ua.a.smem_bytes = savi_bwd_umma_smem_bytes(a.d);
    // instances: C2 (K = 24, 4 heads) and C4 (K = 11), as CTA pairs (B <= 74 clips per GPU) and as single CTAs; generic otherwise
    void (*kern)(BwdUArgs) = savi_bwd_umma_kernel<0, 0, 0>;
    const Dims& dd = ua.a.d;
    if (dd.heads == 4 && dd.K == 24) kern = dd.CN == 2 ? savi_bwd_umma_kernel<24, 2, 4> : dd.CN == 1 ? savi_bwd_umma_kernel<24, 1, 4> : kern;
    else if (dd.heads == 4 && dd.K == 11) kern = dd.CN == 2 ? savi_bwd_umma_kernel<11, 2, 4> : dd.CN == 1 ? savi_bwd_umma_kernel<11, 1, 4> : kern;
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
