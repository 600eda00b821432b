// Tensor-core attention step (bf16 token stream), forward and backward, for one CTA's share of
// the tokens of frame (b,t).  Follows steve.py:76-83 in the folded form.
//
// Per 128-token tile (16 tokens per warp, double-buffered through shared memory with cp.async):
//   phase 1  L^T[slots x 16 tokens] = qk . xhat^T      (A = qk hi/lo from smem, B = xhat tile)
//            softmax over the slot axis: the column of one token lives in the 8 lanes that share
//            lane%4 -> three __shfl_xor (4, 8, 16);  A = P + eps written as bf16 [slot][token]
//   phase 2  numx^T[D x slots] += xhat^T . A            (A-operand = xhat tile via ldmatrix.trans,
//            each warp owns 16-row slices of D, so the accumulators stay in registers all step)
// Backward adds G^T = dUx . xhat^T in phase 1 (same B fragments), forms dL and W = A/S, stages them
// to global for the token-parallel d_inputs kernel, and accumulates d(qk)^T += xhat^T . dL in phase 2.
#pragma once
#include "savi_mma.cuh"

constexpr int TMMA_TN = 128;                     // tokens per tile = 8 warps x 16

__host__ __device__ __forceinline__ int tmma_xs(int D) { return (D + 8) * 2; }          // token/qk row stride, bytes
constexpr int TMMA_ATS = (TMMA_TN + 8) * 2;                                              // [slot][token] row stride, bytes

// shared memory: qk hi/lo (+ dUx hi/lo in backward), `stages` token tiles, 2 (fwd) / 3 (bwd) [slot][token] tiles,
// per-warp slot sums, grad_attn tile (bwd)
__host__ __device__ __forceinline__ size_t tmma_smem_bytes(int MT, int D, int K, int stages, bool bwd) {
    size_t b = (size_t)(bwd ? 4 : 2) * MT * 16 * tmma_xs(D) + (size_t)stages * TMMA_TN * tmma_xs(D) +
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
__device__ __forceinline__ void tmma_issue_tile(unsigned char* xt, const bf16* __restrict__ xh, int n0, int tn, int D) {
    const int xs = tmma_xs(D), chunks = D >> 3;
    const unsigned char* gb = reinterpret_cast<const unsigned char*>(xh + (size_t)n0 * D);
    for (int idx = threadIdx.x; idx < TMMA_TN * chunks; idx += NT) {
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

// L^T (and optionally G^T) for this warp's 16 tokens: acc[m][nt][4]
template <int MT, bool WITH_G>
__device__ __forceinline__ void tmma_phase1_mma(float (&lt)[MT][2][4], float (&gt)[MT][2][4], const unsigned char* xt,
                                                const unsigned char* qkh, const unsigned char* qkl,
                                                const unsigned char* duh, const unsigned char* dul, int D, int warp, int lane) {
    const int xs = tmma_xs(D), mi = lane >> 3, rr = lane & 7;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) { lt[m][n][e] = 0.f; if (WITH_G) gt[m][n][e] = 0.f; }
    const unsigned char* xb = xt + (size_t)(warp * 16 + (mi >> 1) * 8 + rr) * xs + (mi & 1) * 16;
    const size_t aoff = (size_t)((mi & 1) * 8 + rr) * xs + (mi >> 1) * 16;
    for (int kk = 0; kk < D; kk += 16) {
        uint32_t b[4];
        ldsm_x4(b, xb + kk * 2);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            uint32_t a[4];
            ldsm_x4(a, qkh + (size_t)m * 16 * xs + aoff + kk * 2);
            mma16816(lt[m][0], a[0], a[1], a[2], a[3], b[0], b[1]);
            mma16816(lt[m][1], a[0], a[1], a[2], a[3], b[2], b[3]);
            ldsm_x4(a, qkl + (size_t)m * 16 * xs + aoff + kk * 2);
            mma16816(lt[m][0], a[0], a[1], a[2], a[3], b[0], b[1]);
            mma16816(lt[m][1], a[0], a[1], a[2], a[3], b[2], b[3]);
            if (WITH_G) {
                ldsm_x4(a, duh + (size_t)m * 16 * xs + aoff + kk * 2);
                mma16816(gt[m][0], a[0], a[1], a[2], a[3], b[0], b[1]);
                mma16816(gt[m][1], a[0], a[1], a[2], a[3], b[2], b[3]);
                ldsm_x4(a, dul + (size_t)m * 16 * xs + aoff + kk * 2);
                mma16816(gt[m][0], a[0], a[1], a[2], a[3], b[0], b[1]);
                mma16816(gt[m][1], a[0], a[1], a[2], a[3], b[2], b[3]);
            }
        }
    }
}

// in-place softmax over slots of lt (slot = m*16 + g + 8*(e>>1), token column = (nt, e&1)); slots >= K -> 0
template <int MT>
__device__ __forceinline__ void tmma_softmax(float (&lt)[MT][2][4], int K, int g) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            float mx = -INFINITY;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (m * 16 + g + 8 * h < K) mx = fmaxf(mx, lt[m][nt][h * 2 + par]);
            mx = colmax(mx);
            float sum = 0.f;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float e = (m * 16 + g + 8 * h < K) ? expf(lt[m][nt][h * 2 + par] - mx) : 0.f;
                    lt[m][nt][h * 2 + par] = e; sum += e;
                }
            const float inv = 1.0f / colsum(sum);
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) lt[m][nt][h * 2 + par] *= inv;
        }
}

// acc2[mi][ns] += xhat^T (this warp's 16-row D slices) . C[tokens x slots], C stored [slot][token] bf16
template <int MT>
__device__ __forceinline__ void tmma_phase2(float (&acc)[2][2 * MT][4], const unsigned char* xt, const unsigned char* ct,
                                            const unsigned char* ct_lo, int D, int NS, int warp, int lane) {
    const int xs = tmma_xs(D), mi = lane >> 3, rr = lane & 7, md_n = D >> 4;
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
        const int md = warp + w2 * NW;
        if (md < md_n) {
            for (int kt = 0; kt < TMMA_TN; kt += 16) {
                uint32_t a[4];
                ldsm_x4_t(a, xt + (size_t)(kt + (mi >> 1) * 8 + rr) * xs + (md * 16 + (mi & 1) * 8) * 2);
#pragma unroll
                for (int ns = 0; ns < 2 * MT; ++ns) {
                    if (ns < NS) {
                        uint32_t b[2];
                        ldsm_x2(b, ct + (size_t)(ns * 8 + rr) * TMMA_ATS + (kt + (mi & 1) * 8) * 2);
                        mma16816(acc[w2][ns], a[0], a[1], a[2], a[3], b[0], b[1]);
                        ldsm_x2(b, ct_lo + (size_t)(ns * 8 + rr) * TMMA_ATS + (kt + (mi & 1) * 8) * 2);
                        mma16816(acc[w2][ns], a[0], a[1], a[2], a[3], b[0], b[1]);
                    }
                }
            }
        }
    }
}

// acc2 (rows d, cols slots) -> out[slot*D + d] for slot < K
template <int MT>
__device__ __forceinline__ void tmma_store_acc(const float (&acc)[2][2 * MT][4], float* out, int D, int K, int NS, int warp, int lane) {
    const int g = lane >> 2, q = lane & 3, md_n = D >> 4;
#pragma unroll
    for (int w2 = 0; w2 < 2; ++w2) {
        const int md = warp + w2 * NW;
        if (md < md_n) {
#pragma unroll
            for (int ns = 0; ns < 2 * MT; ++ns) {
                if (ns < NS) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int slot = ns * 8 + q * 2 + (e & 1), dd = md * 16 + g + 8 * (e >> 1);
                        if (slot < K) out[(size_t)slot * D + dd] = acc[w2][ns][e];
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int MT>
__device__ void token_pass_fwd_mma(const Dims& d, const bf16* __restrict__ xh, int n_lo, int n_hi, const float* qk_g,
                                   float* part_g, bf16* attn_g, unsigned char* smem, int stages) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int K = d.K, D = d.D, xs = tmma_xs(D), NS = (K + 7) >> 3;
    unsigned char* qkh = smem;
    unsigned char* qkl = qkh + (size_t)MT * 16 * xs;
    unsigned char* xt0 = qkl + (size_t)MT * 16 * xs;
    unsigned char* at = xt0 + (size_t)stages * TMMA_TN * xs;
    unsigned char* atl = at + (size_t)MT * 16 * TMMA_ATS;                        // lo part of A (A = hi + lo)
    unsigned char* pt = atl + (size_t)MT * 16 * TMMA_ATS;
    float* ssum_w = reinterpret_cast<float*>(pt + (size_t)MT * 16 * TMMA_ATS);       // [NW][MT*16]

    __syncthreads();
    stage_rows_hilo<MT>(qkh, qkl, qk_g, K, D);
    float acc2[2][2 * MT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2 * MT; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[a][b][e] = 0.f;
    float ssum[MT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m) { ssum[m][0] = 0.f; ssum[m][1] = 0.f; }

    const int ntile = (n_hi - n_lo + TMMA_TN - 1) / TMMA_TN;
    if (ntile > 0) tmma_issue_tile(xt0, xh, n_lo, min(TMMA_TN, n_hi - n_lo), D);
    for (int ti = 0; ti < ntile; ++ti) {
        const int n0 = n_lo + ti * TMMA_TN, tn = min(TMMA_TN, n_hi - n0);
        unsigned char* xt = xt0 + (size_t)(stages > 1 ? (ti & 1) : 0) * TMMA_TN * xs;
        cp_async_wait_all();
        __syncthreads();                                           // tile ti landed; everyone is past phase 2 of ti-1
        if (stages > 1 && ti + 1 < ntile)
            tmma_issue_tile(xt0 + (size_t)((ti + 1) & 1) * TMMA_TN * xs, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        // ---- phase 1: logits, softmax over slots, A = P + eps ----
        float lt[MT][2][4], dummy[MT][2][4];
        tmma_phase1_mma<MT, false>(lt, dummy, xt, qkh, qkl, nullptr, nullptr, D, warp, lane);
        tmma_softmax<MT>(lt, K, g);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = m * 16 + g + 8 * h;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int tl = warp * 16 + nt * 8 + q * 2;                     // token within the tile
                    const float p0 = lt[m][nt][h * 2], p1 = lt[m][nt][h * 2 + 1];
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
            }
        __syncthreads();
        if (attn_g) {                                              // P -> attn_out[n][k] (steve.py:77, 96)
            for (int n = tid; n < tn; n += NT) {
                bf16* ar = attn_g + (size_t)(n0 + n) * K;
                for (int k = 0; k < K; ++k) ar[k] = *reinterpret_cast<const bf16*>(pt + (size_t)k * TMMA_ATS + n * 2);
            }
        }
        // ---- phase 2: numx^T += xhat^T . A ----
        tmma_phase2<MT>(acc2, xt, at, atl, D, NS, warp, lane);
        if (stages == 1) {
            __syncthreads();
            if (ti + 1 < ntile) tmma_issue_tile(xt0, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        }
    }
    // ---- per-step outputs: partial sum_n A x and sum_n A of this CTA ----
    tmma_store_acc<MT>(acc2, part_g, D, K, NS, warp, lane);
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
}

// ---------------------------------------------------------------------------
// backward.  coef_g: staged [2][MT*16][N] bf16 (dL^T then W^T, token index contiguous) of this (b,t,i)
// ---------------------------------------------------------------------------
template <int MT>
__device__ void token_pass_bwd_mma(const Dims& d, const bf16* __restrict__ xh, int n_lo, int n_hi, const float* qk_g,
                                   const float* dux_g, const float* cvec_g, const bf16* __restrict__ gattn,
                                   bf16* coef_g, float* part_g, unsigned char* smem, int stages) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int K = d.K, KP = d.KP, D = d.D, N = d.N, xs = tmma_xs(D), NS = (K + 7) >> 3;
    unsigned char* qkh = smem;
    unsigned char* qkl = qkh + (size_t)MT * 16 * xs;
    unsigned char* duh = qkl + (size_t)MT * 16 * xs;
    unsigned char* dul = duh + (size_t)MT * 16 * xs;
    unsigned char* xt0 = dul + (size_t)MT * 16 * xs;
    unsigned char* dlt = xt0 + (size_t)stages * TMMA_TN * xs;                 // dL^T [slot][token]
    unsigned char* wtt = dlt + (size_t)MT * 16 * TMMA_ATS;                    // W^T
    unsigned char* dll = wtt + (size_t)MT * 16 * TMMA_ATS;                    // lo part of dL
    float* cv = reinterpret_cast<float*>(dll + (size_t)MT * 16 * TMMA_ATS + (size_t)NW * MT * 16 * 4);   // [2][64]: c, 1/S
    bf16* gs = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(cv) + 2 * 64 * 4);                 // [TN][K] grad_attn tile

    __syncthreads();
    stage_rows_hilo<MT>(qkh, qkl, qk_g, K, D);
    stage_rows_hilo<MT>(duh, dul, dux_g, K, D);
    for (int i = tid; i < 128; i += NT) {
        const int k = i & 63, which = i >> 6;
        cv[i] = (k < K) ? cvec_g[which * KP + k] : 0.f;
    }
    float acc2[2][2 * MT][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2 * MT; ++b)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2[a][b][e] = 0.f;

    const int ntile = (n_hi - n_lo + TMMA_TN - 1) / TMMA_TN;
    if (ntile > 0) tmma_issue_tile(xt0, xh, n_lo, min(TMMA_TN, n_hi - n_lo), D);
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
            tmma_issue_tile(xt0 + (size_t)((ti + 1) & 1) * TMMA_TN * xs, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        float lt[MT][2][4], gt[MT][2][4];
        tmma_phase1_mma<MT, true>(lt, gt, xt, qkh, qkl, duh, dul, D, warp, lane);
        tmma_softmax<MT>(lt, K, g);                                // lt = P
        // dP = (G - c)/S (+ grad_attn);  dL = P (dP - <P, dP>);  W = (P + eps)/S
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int tl = warp * 16 + nt * 8 + q * 2 + par;
                float dot = 0.f;
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int slot = m * 16 + g + 8 * h, e = h * 2 + par;
                        float dp = 0.f;
                        if (slot < K) {
                            dp = (gt[m][nt][e] - cv[slot]) * cv[64 + slot];
                            if (gattn && tl < tn) dp += __bfloat162float(gs[(size_t)tl * K + slot]);
                        }
                        gt[m][nt][e] = dp;
                        dot = fmaf(lt[m][nt][e], dp, dot);
                    }
                dot = colsum(dot);
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int slot = m * 16 + g + 8 * h, e = h * 2 + par;
                        const bool ok = slot < K && tl < tn;
                        const float p = lt[m][nt][e];
                        gt[m][nt][e] = ok ? p * (gt[m][nt][e] - dot) : 0.f;               // dL
                        lt[m][nt][e] = ok ? (p + d.eps) * cv[64 + slot] : 0.f;            // W
                    }
            }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int slot = m * 16 + g + 8 * h;
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int tl = warp * 16 + nt * 8 + q * 2;
                    bf16 d0, d1, e0, e1;
                    split_bf16(gt[m][nt][h * 2], d0, e0); split_bf16(gt[m][nt][h * 2 + 1], d1, e1);
                    *reinterpret_cast<__nv_bfloat162*>(dlt + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(d0, d1);
                    *reinterpret_cast<__nv_bfloat162*>(dll + (size_t)slot * TMMA_ATS + tl * 2) = __halves2bfloat162(e0, e1);
                    *reinterpret_cast<uint32_t*>(wtt + (size_t)slot * TMMA_ATS + tl * 2) = pack_bf16(lt[m][nt][h * 2], lt[m][nt][h * 2 + 1]);
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
        tmma_phase2<MT>(acc2, xt, dlt, dll, D, NS, warp, lane);    // d(qk)^T += xhat^T . dL (hi + lo)
        if (stages == 1) {
            __syncthreads();
            if (ti + 1 < ntile) tmma_issue_tile(xt0, xh, n0 + TMMA_TN, min(TMMA_TN, n_hi - n0 - TMMA_TN), D);
        }
    }
    tmma_store_acc<MT>(acc2, part_g, D, K, NS, warp, lane);
}
