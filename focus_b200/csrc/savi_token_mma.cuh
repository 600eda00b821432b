// Tensor-core attention step (bf16 token stream), forward and backward, for one CTA's share of
// the tokens of frame (b,t).  Follows steve.py:76-83 in the folded form.
//
// Per 128-token tile (8 tokens per warp, 16 warps; tiles double-buffered through shared memory
// with cp.async):
//   phase 1  L^T[slots x 8 tokens] = qk . xhat^T       (A = qk hi/lo from smem, B = xhat tile)
//            softmax over the slot axis: the column of one token lives in the 8 lanes that share
//            lane%4 -> three __shfl_xor (4, 8, 16);  A = P + eps written as bf16 hi/lo [slot][token]
//   phase 2  numx^T[D x slots] += xhat^T . A            (A-operand = xhat tile via ldmatrix.trans;
//            a warp owns one 16-row slice of D and one slice of the tile's tokens, the accumulators
//            stay in registers for the whole step and are combined once at its end)
// Backward adds G^T = dUx . xhat^T in phase 1 (same B fragments), forms dL and W = A/S, stages them
// to global for the token-parallel d_inputs kernel, and accumulates d(qk)^T += xhat^T . dL in phase 2.
#pragma once
#include "savi_mma.cuh"

constexpr int TMMA_TN = 128;                     // tokens per tile
constexpr int TMMA_TPW = TMMA_TN / NW;           // tokens per warp in phase 1 (8 = one mma n-tile)
static_assert(TMMA_TPW == 8, "phase 1 assumes one 8-token n-tile per warp");

__host__ __device__ __forceinline__ int tmma_xs(int D) { return (D + 8) * 2; }          // token/qk row stride, bytes
constexpr int TMMA_ATS = (TMMA_TN + 8) * 2;                                              // [slot][token] row stride, bytes

// token slices of a tile in phase 2: as many as there are spare warps, at least 16 tokens each
__host__ __device__ __forceinline__ int tmma_kg(int D) {
    const int md_n = D >> 4;
    int kg = md_n >= NW ? 1 : NW / md_n;
    return kg > TMMA_TN / 16 ? TMMA_TN / 16 : kg;
}
__host__ __device__ __forceinline__ size_t tmma_tile_region(int D, int K, int stages) {
    size_t tiles = (size_t)stages * TMMA_TN * tmma_xs(D);
    const size_t comb = (size_t)tmma_kg(D) * K * D * 4;      // end-of-step combine buffer aliases the tiles
    return tiles < comb ? comb : tiles;
}
// shared memory: qk hi/lo (+ dUx hi/lo in backward), token tiles, three [slot][token] tiles,
// per-warp slot sums, c / 1/S vectors, grad_attn tile (bwd)
__host__ __device__ __forceinline__ size_t tmma_smem_bytes(int MT, int D, int K, int stages, bool bwd) {
    size_t b = (size_t)(bwd ? 4 : 2) * MT * 16 * tmma_xs(D) + tmma_tile_region(D, K, stages) +
               (size_t)3 * MT * 16 * TMMA_ATS + (size_t)NW * MT * 16 * 4 + 2 * 64 * 4;
    if (bwd) b += (size_t)TMMA_TN * K * 2 + 16;
    return (b + 15) / 16 * 16;
}

// global fp32 [K][D] -> shared bf16 hi/lo [MT*16][D+8] (rows >= K zero)
template <int MT>
__device__ __forceinline__ void stage_rows_hilo(unsigned char* hi, unsigned char* lo, const float* src, int K, int D) {
    const int xs = tmma_xs(D), d4 = D >> 2;
    for (int idx = threadIdx.x; idx < MT * 16 * d4; idx += NT) {
        const int r = idx / d4, c = (idx - r * d4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < K) v = ld4(src + (size_t)r * D + c);
        bf16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
        __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
        __nv_bfloat162 e = __halves2bfloat162(l0, l1), f = __halves2bfloat162(l2, l3);
        uint2 ph, pl;
        ph.x = *reinterpret_cast<uint32_t*>(&a); ph.y = *reinterpret_cast<uint32_t*>(&b);
        pl.x = *reinterpret_cast<uint32_t*>(&e); pl.y = *reinterpret_cast<uint32_t*>(&f);
        *reinterpret_cast<uint2*>(hi + (size_t)r * xs + c * 2) = ph;
        *reinterpret_cast<uint2*>(lo + (size_t)r * xs + c * 2) = pl;
    }
}

// rows [n0, n0+tn) of the frame -> shared tile (cp.async, not waited); rows >= tn zero-filled
template <int NTHREADS>
__device__ __forceinline__ void tmma_issue_tile(unsigned char* xt, const bf16* __restrict__ xh, int n0, int tn, int D) {
    const int xs = tmma_xs(D), chunks = D >> 3;
    const unsigned char* gb = reinterpret_cast<const unsigned char*>(xh + (size_t)n0 * D);
    for (int idx = threadIdx.x; idx < TMMA_TN * chunks; idx += NTHREADS) {
        const int r = idx / chunks, c = idx - r * chunks;
        if (r < tn) cp_async16(xt + (size_t)r * xs + c * 16, gb + (size_t)r * D * 2 + c * 16);
        else *reinterpret_cast<uint4*>(xt + (size_t)r * xs + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

// max / sum over the slot axis of a token column: in-thread values first, then the 8 lanes sharing lane%4
__device__ __forceinline__ float colmax(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
}
__device__ __forceinline__ float colsum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v + __shfl_xor_sync(0xffffffffu, v, 16);
}

// L^T (and optionally G^T) for this warp's 8 tokens: acc[m][4] with
// e: slot = m*16 + g + 8*(e>>1), token = warp*8 + 2q + (e&1)
template <int MT, bool WITH_G>
__device__ __forceinline__ void tmma_phase1_mma(float (&lt)[MT][4], float (&gt)[MT][4], const unsigned char* xt,
                                                const unsigned char* qkh, const unsigned char* qkl,
                                                const unsigned char* duh, const unsigned char* dul, int D, int warp, int lane) {
    const int xs = tmma_xs(D), mi = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e) { lt[m][e] = 0.f; if (WITH_G) gt[m][e] = 0.f; }
    const unsigned char* xb = xt + (size_t)(warp * TMMA_TPW + rr) * xs + (mi & 1) * 16;     // x2: lanes 0-15 give the rows
    const size_t aoff = (size_t)((mi & 1) * 8 + rr) * xs + (mi >> 1) * 16;
#pragma unroll 2
    for (int kk = 0; kk < D; kk += 16) {
        uint32_t b[2];
        ldsm_x2(b, xb + kk * 2);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            uint32_t a[4];
            ldsm_x4(a, qkh + (size_t)m * 16 * xs + aoff + kk * 2);
            mma16816(lt[m], a[0], a[1], a[2], a[3], b[0], b[1]);
            ldsm_x4(a, qkl + (size_t)m * 16 * xs + aoff + kk * 2);
            mma16816(lt[m], a[0], a[1], a[2], a[3], b[0], b[1]);
            if (WITH_G) {
                ldsm_x4(a, duh + (size_t)m * 16 * xs + aoff + kk * 2);
                mma16816(gt[m], a[0], a[1], a[2], a[3], b[0], b[1]);
                ldsm_x4(a, dul + (size_t)m * 16 * xs + aoff + kk * 2);
                mma16816(gt[m], a[0], a[1], a[2], a[3], b[0], b[1]);
            }
        }
    }
}

// in-place softmax over slots of lt; slots >= K -> 0.  exp2 with pre-scaled logits (1 FFMA + MUFU).
template <int MT>
__device__ __forceinline__ void tmma_softmax(float (&lt)[MT][4], int K, int g) {
    constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll
    for (int par = 0; par < 2; ++par) {
        float mx = -INFINITY;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (m * 16 + g + 8 * h < K) mx = fmaxf(mx, lt[m][h * 2 + par]);
        mx = colmax(mx) * LOG2E;
        float sum = 0.f;
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float e = (m * 16 + g + 8 * h < K) ? exp2f(fmaf(lt[m][h * 2 + par], LOG2E, -mx)) : 0.f;
                lt[m][h * 2 + par] = e; sum += e;
            }
        const float inv = 1.0f / colsum(sum);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) lt[m][h * 2 + par] *= inv;
    }
}

// Phase-2 work split: D/16 row slices x (NW / (D/16)) token slices of the tile.
struct P2Map { int md, k_lo, k_hi, kgi; bool active; };
__device__ __forceinline__ P2Map tmma_p2map(int D, int warp) {
    const int md_n = D >> 4;
    const int kg = tmma_kg(D);
    P2Map m;
    m.active = warp < md_n * kg;
    m.md = warp % md_n;
    m.kgi = warp / md_n;
    const int span = TMMA_TN / kg;
    m.k_lo = m.kgi * span; m.k_hi = m.k_lo + span;
    return m;
}

// acc[ns] += xhat^T (16-row slice md of D) . C[tokens k_lo..k_hi x slots], C stored [slot][token] bf16 (hi + lo)
template <int MT>
__device__ __forceinline__ void tmma_phase2(float (&acc)[2 * MT][4], const unsigned char* xt, const unsigned char* ct,
                                            const unsigned char* ct_lo, int D, int NS, const P2Map& pm, int lane) {
    const int xs = tmma_xs(D), mi = lane >> 3, rr = lane & 7;
    if (!pm.active) return;
    for (int kt = pm.k_lo; kt < pm.k_hi; kt += 16) {
        uint32_t a[4];
        ldsm_x4_t(a, xt + (size_t)(kt + (mi >> 1) * 8 + rr) * xs + (pm.md * 16 + (mi & 1) * 8) * 2);
#pragma unroll
        for (int ns = 0; ns < 2 * MT; ++ns) {
            if (ns < NS) {
                uint32_t b[2];
                ldsm_x2(b, ct + (size_t)(ns * 8 + rr) * TMMA_ATS + (kt + (mi & 1) * 8) * 2);
                mma16816(acc[ns], a[0], a[1], a[2], a[3], b[0], b[1]);
                ldsm_x2(b, ct_lo + (size_t)(ns * 8 + rr) * TMMA_ATS + (kt + (mi & 1) * 8) * 2);
                mma16816(acc[ns], a[0], a[1], a[2], a[3], b[0], b[1]);
            }
        }
    }
}

// End of step: per-warp accumulators (rows d, cols slots) -> comb[kgroup][slot*D + d] in shared memory, then
// out[slot*D + d] = sum over k-groups (fixed order).  `comb` aliases the token tiles (callers sync first).
template <int MT>
__device__ __forceinline__ void tmma_reduce_store(const float (&acc)[2 * MT][4], float* comb, float* out, int D, int K, int NS,
                                                  const P2Map& pm, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int kg = tmma_kg(D);
    if (pm.active) {
        float* cb = comb + (size_t)pm.kgi * K * D;
#pragma unroll
        for (int ns = 0; ns < 2 * MT; ++ns) {
            if (ns < NS) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int slot = ns * 8 + q * 2 + (e & 1), dd = pm.md * 16 + g + 8 * (e >> 1);
                    if (slot < K) cb[(size_t)slot * D + dd] = acc[ns][e];
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x * 4; i < K * D; i += NT * 4) {
        float4 s = ld4(comb + i);
        for (int r = 1; r < kg; ++r) { const float4 t = ld4(comb + (size_t)r * K * D + i); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        st4(out + i, s);
    }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int MT>
__device__ __noinline__ void token_pass_fwd_mma(const Dims& d, const bf16* __restrict__ xh, int n_lo, int n_hi, const float* qk_g,
                                                float* part_g, bf16* attn_g, unsigned char* smem, int stages, long long* dbg = nullptr) {
    long long ph_last = clock64();
#define TP_PH(id) do { if (dbg) { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); \
    atomicAdd(reinterpret_cast<unsigned long long*>(dbg + (id)), (unsigned long long)(t_ - ph_last)); ph_last = t_; } } } while (0)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int K = d.K, D = d.D, xs = tmma_xs(D), NS = (K + 7) >> 3;
    unsigned char* qkh = smem;
    unsigned char* qkl = qkh + (size_t)MT * 16 * xs;
    unsigned char* xt0 = qkl + (size_t)MT * 16 * xs;
    unsigned char* at = xt0 + tmma_tile_region(D, K, stages);
    unsigned char* atl = at + (size_t)MT * 16 * TMMA_ATS;                        // lo part of A (A = hi + lo)
    unsigned char* pt = atl + (size_t)MT * 16 * TMMA_ATS;
    float* ssum_w = reinterpret_cast<float*>(pt + (size_t)MT * 16 * TMMA_ATS);   // [NW][MT*16]
    const P2Map pm = tmma_p2map(D, warp);

    __syncthreads();
    stage_rows_hilo<MT>(qkh, qkl, qk_g, K, D);
    float acc2[2 * MT][4];
#pragma unroll
    for (int b = 0; b < 2 * MT; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[b][e] = 0.f;
    float ssum[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) { ssum[m][0] = 0.f; ssum[m][1] = 0.f; }

    const int ntile = (n_hi - n_lo + TMMA_TN - 1) / TMMA_TN;
    if (ntile > 0) tmma_issue_tile<NT>(xt0, xh, n_lo, min(TMMA_TN, n_hi - n_lo), D);
    TP_PH(50);
    for (int ti = 0; ti < ntile; ++ti) {
        const int n0 = n_lo + ti * TMMA_TN, tn = min(TMMA_TN, n_hi - n0);
        unsigned char* xt = xt0 + (size_t)(stages > 1 ? (ti & 1) : 0) * TMMA_TN * xs;
        cp_async_wait_all();
        __syncthreads();                                           // tile ti landed; everyone is past phase 2 of ti-1
        TP_PH(51);
        if (stages > 1 && ti + 1 < ntile)
            tmma_issue_tile<NT>(xt0 + (size_t)((ti + 1) & 1) * TMMA_TN * xs, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        // ---- phase 1: logits, softmax over slots, A = P + eps ----
        float lt[MT][4], dummy[MT][4];
        tmma_phase1_mma<MT, false>(lt, dummy, xt, qkh, qkl, nullptr, nullptr, D, warp, lane);
        TP_PH(52);
        tmma_softmax<MT>(lt, K, g);
        TP_PH(53);
        const int tl = warp * TMMA_TPW + q * 2;                    // this thread's token pair within the tile
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = m * 16 + g + 8 * h;
                const float p0 = lt[m][h * 2], p1 = lt[m][h * 2 + 1];
                const bool s_ok = slot < K;
                const float f0 = (s_ok && tl < tn) ? p0 + d.eps : 0.f, f1 = (s_ok && tl + 1 < tn) ? p1 + d.eps : 0.f;
                bf16 a0, a1, l0, l1;
                split_bf16(f0, a0, l0); split_bf16(f1, a1, l1);
                // the token sum uses exactly the weights the tensor cores see (hi + lo)
                ssum[m][h] += (__bfloat162float(a0) + __bfloat162float(l0)) + (__bfloat162float(a1) + __bfloat162float(l1));
                *reinterpret_cast<__nv_bfloat162*>(at + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(a0, a1);
                *reinterpret_cast<__nv_bfloat162*>(atl + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(l0, l1);
                if (attn_g) *reinterpret_cast<uint32_t*>(pt + (size_t)slot * TMMA_ATS + tl * 2) = pack_bf16(p0, p1);
            }
        __syncthreads();
        TP_PH(54);
        if (attn_g) {                                              // P -> attn_out[n][k] (steve.py:77, 96): contiguous tn*K block
            bf16* ab = attn_g + (size_t)n0 * K;
            for (int i = tid; i < tn * K; i += NT) {
                const int n = i / K, k = i - n * K;
                ab[i] = *reinterpret_cast<const bf16*>(pt + (size_t)k * TMMA_ATS + n * 2);
            }
        }
        TP_PH(55);
        // ---- phase 2: numx^T += xhat^T . A ----
        tmma_phase2<MT>(acc2, xt, at, atl, D, NS, pm, lane);
        TP_PH(56);
        if (stages == 1) {
            __syncthreads();
            if (ti + 1 < ntile) tmma_issue_tile<NT>(xt0, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        }
    }
    cp_async_wait_all();
    __syncthreads();                                               // token tiles are dead: reuse them as the combine buffer
    tmma_reduce_store<MT>(acc2, reinterpret_cast<float*>(xt0), part_g, D, K, NS, pm, lane);
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float s = ssum[m][h];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) ssum_w[warp * MT * 16 + m * 16 + g + 8 * h] = s;
        }
    __syncthreads();
    for (int k = tid; k < d.KP; k += NT) {
        float s = 0.f;
        if (k < K) for (int w = 0; w < NW; ++w) s += ssum_w[w * MT * 16 + k];
        part_g[(size_t)K * D + k] = s;
    }
    TP_PH(57);
#undef TP_PH
}

// ---------------------------------------------------------------------------
// backward.  coef_g: staged [2][MT*16][N] bf16 (dL^T then W^T, token index contiguous) of this (b,t,i)
// ---------------------------------------------------------------------------
template <int MT>
__device__ __noinline__ void token_pass_bwd_mma(const Dims& d, const bf16* __restrict__ xh, int n_lo, int n_hi, const float* qk_g,
                                                const float* dux_g, const float* cvec_g, const bf16* __restrict__ gattn,
                                                bf16* coef_g, float* part_g, unsigned char* smem, int stages) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int K = d.K, KP = d.KP, D = d.D, N = d.N, xs = tmma_xs(D), NS = (K + 7) >> 3;
    unsigned char* qkh = smem;
    unsigned char* qkl = qkh + (size_t)MT * 16 * xs;
    unsigned char* duh = qkl + (size_t)MT * 16 * xs;
    unsigned char* dul = duh + (size_t)MT * 16 * xs;
    unsigned char* xt0 = dul + (size_t)MT * 16 * xs;
    unsigned char* dlt = xt0 + tmma_tile_region(D, K, stages);                // dL^T [slot][token] (hi)
    unsigned char* wtt = dlt + (size_t)MT * 16 * TMMA_ATS;                    // W^T
    unsigned char* dll = wtt + (size_t)MT * 16 * TMMA_ATS;                    // lo part of dL
    float* cv = reinterpret_cast<float*>(dll + (size_t)MT * 16 * TMMA_ATS + (size_t)NW * MT * 16 * 4);   // [2][64]: c, 1/S
    bf16* gs = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(cv) + 2 * 64 * 4);                 // [TN][K] grad_attn tile
    const P2Map pm = tmma_p2map(D, warp);

    __syncthreads();
    stage_rows_hilo<MT>(qkh, qkl, qk_g, K, D);
    stage_rows_hilo<MT>(duh, dul, dux_g, K, D);
    for (int i = tid; i < 128; i += NT) {
        const int k = i & 63, which = i >> 6;
        cv[i] = (k < K) ? cvec_g[which * KP + k] : 0.f;
    }
    float acc2[2 * MT][4];
#pragma unroll
    for (int b = 0; b < 2 * MT; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc2[b][e] = 0.f;

    const int ntile = (n_hi - n_lo + TMMA_TN - 1) / TMMA_TN;
    if (ntile > 0) tmma_issue_tile<NT>(xt0, xh, n_lo, min(TMMA_TN, n_hi - n_lo), D);
    for (int ti = 0; ti < ntile; ++ti) {
        const int n0 = n_lo + ti * TMMA_TN, tn = min(TMMA_TN, n_hi - n0);
        unsigned char* xt = xt0 + (size_t)(stages > 1 ? (ti & 1) : 0) * TMMA_TN * xs;
        if (gattn) {                                               // this tile's grad_attn rows are contiguous in global
            const bf16* src = gattn + (size_t)n0 * K;
            for (int i = tid; i < tn * K; i += NT) gs[i] = src[i];
        }
        cp_async_wait_all();
        __syncthreads();
        if (stages > 1 && ti + 1 < ntile)
            tmma_issue_tile<NT>(xt0 + (size_t)((ti + 1) & 1) * TMMA_TN * xs, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        float lt[MT][4], gt[MT][4];
        tmma_phase1_mma<MT, true>(lt, gt, xt, qkh, qkl, duh, dul, D, warp, lane);
        tmma_softmax<MT>(lt, K, g);                                // lt = P
        // dP = (G - c)/S (+ grad_attn);  dL = P (dP - <P, dP>);  W = (P + eps)/S
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int tl = warp * TMMA_TPW + q * 2 + par;
            float dot = 0.f;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int slot = m * 16 + g + 8 * h, e = h * 2 + par;
                    float dp = 0.f;
                    if (slot < K) {
                        dp = (gt[m][e] - cv[slot]) * cv[64 + slot];
                        if (gattn && tl < tn) dp += __bfloat162float(gs[(size_t)tl * K + slot]);
                    }
                    gt[m][e] = dp;
                    dot = fmaf(lt[m][e], dp, dot);
                }
            dot = colsum(dot);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int slot = m * 16 + g + 8 * h, e = h * 2 + par;
                    const bool ok = slot < K && tl < tn;
                    const float p = lt[m][e];
                    gt[m][e] = ok ? p * (gt[m][e] - dot) : 0.f;               // dL
                    lt[m][e] = ok ? (p + d.eps) * cv[64 + slot] : 0.f;        // W
                }
        }
        {
            const int tl = warp * TMMA_TPW + q * 2;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int slot = m * 16 + g + 8 * h;
                    bf16 d0, d1, e0, e1;
                    split_bf16(gt[m][h * 2], d0, e0); split_bf16(gt[m][h * 2 + 1], d1, e1);
                    *reinterpret_cast<__nv_bfloat162*>(dlt + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(d0, d1);
                    *reinterpret_cast<__nv_bfloat162*>(dll + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(e0, e1);
                    *reinterpret_cast<uint32_t*>(wtt + (size_t)slot * TMMA_ATS + tl * 2) = pack_bf16(lt[m][h * 2], lt[m][h * 2 + 1]);
                }
        }
        __syncthreads();
        // stage dL^T and W^T for the d_inputs kernel: rows of 128 tokens are contiguous in global
        {
            const int cpr = TMMA_TN / 8;                                   // 16-byte chunks per row
            for (int idx = tid; idx < 2 * MT * 16 * cpr; idx += NT) {
                const int row = idx / cpr, c = idx - row * cpr;           // row < 2*MT*16: [dL | W] x slot
                const unsigned char* src = (row < MT * 16 ? dlt + (size_t)row * TMMA_ATS : wtt + (size_t)(row - MT * 16) * TMMA_ATS) + c * 16;
                bf16* dst = coef_g + (size_t)row * N + n0 + c * 8;
                if (n0 + c * 8 + 8 <= n_hi && ((N & 7) == 0)) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
                else {
                    const bf16* s2 = reinterpret_cast<const bf16*>(src);
                    for (int e = 0; e < 8; ++e) if (n0 + c * 8 + e < n_hi) dst[e] = s2[e];
                }
            }
        }
        tmma_phase2<MT>(acc2, xt, dlt, dll, D, NS, pm, lane);     // d(qk)^T += xhat^T . dL (hi + lo)
        if (stages == 1) {
            __syncthreads();
            if (ti + 1 < ntile) tmma_issue_tile<NT>(xt0, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    tmma_reduce_store<MT>(acc2, reinterpret_cast<float*>(xt0), part_g, D, K, NS, pm, lane);
}
