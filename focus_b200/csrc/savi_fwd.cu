// Forward kernels: token LayerNorm prologue and the per-clip recurrence kernel.
//
// Reference semantics: /root/reference/slowfast/models/STEVE/steve.py:52-105
// (SlotAttentionVideo.forward) and transformer.py:22-49, 70-86, 106-114.
#include <cstdlib>
#include "savi_token_mma.cuh"
#include "savi_args.h"

// ---------------------------------------------------------------------------
// K1: xhat = LayerNorm_D(inputs) (steve.py:60), one warp per token.
// Streams `inputs` once from HBM with 16-byte accesses, writes xhat in the token
// dtype plus (mean, rstd) per token for the backward.
// ---------------------------------------------------------------------------
// tcgen05 path (ximg != nullptr): xhat is ALSO written as [frame][tile][D/64][128 rows][64] SWIZZLE_128B operand
// blocks (savi_umma.cuh), rows past N zero-filled, so the clip kernels fetch a token tile with two 16 KB bulk copies.
__device__ __forceinline__ unsigned char* ximg_chunk(unsigned char* ximg, int64_t frame, int n, int d, int NTILE, int D) {
    const int tile = n >> 7, r = n & 127, db = d >> 6, dc = d & 63;
    return ximg + ((frame * NTILE + tile) * (D >> 6) + db) * (int64_t)16384 + r * 128 + (((dc >> 3) ^ (r & 7)) << 4) + (dc & 7) * 2;
}
template <typename TokT>
__global__ void __launch_bounds__(NT) ln_tokens_fwd_kernel(const TokT* __restrict__ x, TokT* __restrict__ xhat,
                                                           float2* __restrict__ stats, const float* __restrict__ g,
                                                           const float* __restrict__ b, int64_t rows, int D, float eps,
                                                           unsigned char* __restrict__ ximg, int N, int NTILE) {
    constexpr int VEC = Tok<TokT>::VEC;
    const int lane = threadIdx.x & 31;
    const int chunks = D / VEC;                       // 16-byte chunks per row (<= 64 for D <= 512)
    const int64_t warp0 = (int64_t)blockIdx.x * NW + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * NW;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const TokT* xr = x + row * D;
        float v[2][VEC];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = lane + 32 * j;
            if (c < chunks) {
                Tok<TokT>::load(xr + c * VEC, v[j]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) s += v[j][e];
            }
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = lane + 32 * j;
            if (c < chunks) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) { float t = v[j][e] - mean; q = fmaf(t, t, q); }
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
        if (lane == 0) stats[row] = make_float2(mean, rstd);
        TokT* yr = xhat + row * D;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = lane + 32 * j;
            if (c < chunks) {
#pragma unroll
                for (int e = 0; e < VEC; e += 4) {
                    const int d = c * VEC + e;
                    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + d));
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + d));
                    float4 o;
                    o.x = (v[j][e + 0] - mean) * rstd * gg.x + bb.x;
                    o.y = (v[j][e + 1] - mean) * rstd * gg.y + bb.y;
                    o.z = (v[j][e + 2] - mean) * rstd * gg.z + bb.z;
                    o.w = (v[j][e + 3] - mean) * rstd * gg.w + bb.w;
                    if (xhat) Tok<TokT>::store4(yr + d, o);
                    if (ximg) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
                        uint2 t; t.x = *reinterpret_cast<unsigned*>(&p0); t.y = *reinterpret_cast<unsigned*>(&p1);
                        *reinterpret_cast<uint2*>(ximg_chunk(ximg, row / N, (int)(row % N), d, NTILE, D)) = t;
                    }
                }
            }
        }
    }
    if (ximg && (N & 127)) {          // zero the padding rows of every frame's last tile
        const int pad = NTILE * 128 - N, c16 = D >> 3;
        const int64_t frames = rows / N, items = frames * pad * c16;
        for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < items; i += (int64_t)gridDim.x * NT) {
            const int c = (int)(i % c16);
            const int64_t t = i / c16;
            const int n = N + (int)(t % pad);
            *reinterpret_cast<uint4*>(ximg_chunk(ximg, t / pad, n, c * 8, NTILE, D)) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// tcgen05 path, D = 128 bf16: one CTA per (frame, 128-token tile), 16 lanes per token (one 16-byte chunk each), two tokens per
// warp and pass, four passes of a warp loaded up front (4 x 16 B in flight per thread at 64 registers -> 4 CTAs per SM; no index
// divisions in the loop).
// Output only as SWIZZLE_128B operand blocks: reads 256 B and writes 256 B per token with full 16-byte accesses.
static __global__ void __launch_bounds__(256, 4) ln_tokens_fwd_img128_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats,
                                                                   const float* __restrict__ g, const float* __restrict__ b,
                                                                   float eps, unsigned char* __restrict__ ximg, int N, int NTILE) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 15, half = lane >> 4;
    const int frame = blockIdx.x / NTILE, tile = blockIdx.x - frame * NTILE;
    float gg[8], bb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { gg[e] = __ldg(g + sub * 8 + e); bb[e] = __ldg(b + sub * 8 + e); }
    const __nv_bfloat16* xf = x + (size_t)frame * N * 128;
#pragma unroll 1
    for (int it0 = 0; it0 < 8; it0 += 4) {              // two batches of four rows: 4 x 16 B in flight per thread at 64 registers
    uint4 raw[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int n = tile * 128 + (it0 + it) * 16 + warp * 2 + half;
        raw[it] = (n < N) ? __ldg(reinterpret_cast<const uint4*>(xf + (size_t)n * 128 + sub * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int n = tile * 128 + (it0 + it) * 16 + warp * 2 + half;
        const bool real = n < N;
        float v[8];
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&raw[it]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(hp[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[e];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 128.0f);
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float t = v[e] - mean; q = fmaf(t, t, q); }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = 1.0f / sqrtf(q * (1.0f / 128.0f) + eps);
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (real) {
            if (sub == 0) stats[(size_t)frame * N + n] = make_float2(mean, rstd);
            float y[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
            __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]), p1 = __floats2bfloat162_rn(y[2], y[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(y[4], y[5]), p3 = __floats2bfloat162_rn(y[6], y[7]);
            out.x = *reinterpret_cast<unsigned*>(&p0); out.y = *reinterpret_cast<unsigned*>(&p1);
            out.z = *reinterpret_cast<unsigned*>(&p2); out.w = *reinterpret_cast<unsigned*>(&p3);
        }
        *reinterpret_cast<uint4*>(ximg_chunk(ximg, frame, n, sub * 8, NTILE, 128)) = out;   // padding rows are zero-filled
    }
    }
}

// ---------------------------------------------------------------------------
// Attention step over this CTA's share of the tokens of frame (b,t)
// (steve.py:76-83 in the folded form):
//   L = xhat qk^T ; P = softmax_K(L) ; A = P + eps ; part = [sum_n A x , sum_n A]
// One thread per token for the logits/softmax (all K logits in registers), then a
// register-tiled outer-product accumulation over the tile.
// ---------------------------------------------------------------------------
template <typename TokT, int KMAX>
__device__ __noinline__ void token_pass_fwd(const Dims& d, const TokT* __restrict__ xh, int n_lo, int n_hi,
                               const float* qk_g, float* part_g, TokT* attn_g, unsigned char* smem, int TN) {
    const int tid = threadIdx.x, K = d.K, KP = d.KP, D = d.D;
    const int cst = coef_stride(KP), xst = tile_stride_bytes(D, sizeof(TokT));
    float* qk_s = reinterpret_cast<float*>(smem);                  // [KMAX][D], rows >= K are zero
    float* acc_s = qk_s + (size_t)KMAX * D;                        // [KP][D]
    float* ssum_s = acc_s + (size_t)KP * D;                        // [KP]
    float* coef = ssum_s + KP;                                     // [TN][cst]
    unsigned char* xs = reinterpret_cast<unsigned char*>(coef + (size_t)TN * cst);

    __syncthreads();
    for (int i = tid * 4; i < KMAX * D; i += NT * 4)
        st4(qk_s + i, (i < K * D) ? ld4(qk_g + i) : make_float4(0.f, 0.f, 0.f, 0.f));
    for (int i = tid; i < KP * D + KP; i += NT) acc_s[i] = 0.f;    // acc_s and ssum_s are contiguous

    for (int n0 = n_lo; n0 < n_hi; n0 += TN) {
        const int tn = min(TN, n_hi - n0);
        __syncthreads();
        load_token_tile<TokT>(xs, xst, xh, n0, tn, D);
        __syncthreads();
        for (int n = tid; n < tn; n += NT) {
            float acc[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
            const TokT* xr = reinterpret_cast<const TokT*>(xs + (size_t)n * xst);
            constexpr int VEC = Tok<TokT>::VEC;
            for (int c = 0; c < D; c += VEC) {
                float xv[VEC];
                Tok<TokT>::load(xr + c, xv);
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
#pragma unroll
                    for (int e = 0; e < VEC; e += 4) {
                        const float4 w = ld4(qk_s + (size_t)k * D + c + e);
                        acc[k] = fmaf(xv[e], w.x, acc[k]); acc[k] = fmaf(xv[e + 1], w.y, acc[k]);
                        acc[k] = fmaf(xv[e + 2], w.z, acc[k]); acc[k] = fmaf(xv[e + 3], w.w, acc[k]);
                    }
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) mx = fmaxf(mx, acc[k]);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) { acc[k] = (k < K) ? expf(acc[k] - mx) : 0.f; sum += acc[k]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) acc[k] *= inv;          // P (softmax over slots, steve.py:77)
            if (attn_g) {
                TokT* ar = attn_g + (size_t)(n0 + n) * K;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) if (k < K) ar[k] = Tok<TokT>::from_f(acc[k]);
            }
            float* cr = coef + (size_t)n * cst;
#pragma unroll
            for (int k = 0; k < KMAX; k += 4) {
                if (k < KP) {
                    float4 a;
                    a.x = (k + 0 < K) ? acc[k + 0] + d.eps : 0.f;   // A = P + eps (steve.py:81)
                    a.y = (k + 1 < K) ? acc[k + 1] + d.eps : 0.f;
                    a.z = (k + 2 < K) ? acc[k + 2] + d.eps : 0.f;
                    a.w = (k + 3 < K) ? acc[k + 3] + d.eps : 0.f;
                    st4(cr + k, a);
                }
            }
        }
        __syncthreads();
        tile_outer_accum<TokT>(acc_s, ssum_s, coef, cst, xs, xst, tn, KP, D);
    }
    __syncthreads();
    for (int i = tid; i < K * D; i += NT) part_g[i] = acc_s[i];
    for (int i = tid; i < KP; i += NT) part_g[(size_t)K * D + i] = ssum_s[i];
}

// Multi-head self-attention core of the predictor on one clip (transformer.py:34-47).
// Q is already scaled by dh^-1/2.  att: [H][K][K] (global), O: [K][Ds].
// m (nullable): training-mode dropout mask of the probabilities [H][K][K] (0 or 1/(1-p)); `att` keeps the UNDROPPED softmax.
static __device__ __noinline__ void mha_core_fwd(const float* Q, const float* Kk, const float* V, float* att, float* O,
                             int K, int Ds, int H, const float* m) {
    const int dh = Ds / H, tid = threadIdx.x;
    for (int idx = tid; idx < H * K * K; idx += NT) {
        const int j = idx % K, i = (idx / K) % K, h = idx / (K * K);
        const float* q = Q + (size_t)i * Ds + h * dh;
        const float* k = Kk + (size_t)j * Ds + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s = fmaf(q[c], k[c], s);
        att[idx] = s;
    }
    __syncthreads();
    for (int row = tid; row < H * K; row += NT) {
        float* a = att + (size_t)row * K;
        float mx = -INFINITY;
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, a[j]);
        float sum = 0.f;
        for (int j = 0; j < K; ++j) { float e = expf(a[j] - mx); a[j] = e; sum += e; }
        const float inv = 1.0f / sum;
        for (int j = 0; j < K; ++j) a[j] *= inv;
    }
    __syncthreads();
    for (int idx = tid; idx < K * Ds; idx += NT) {
        const int c = idx % Ds, i = idx / Ds, h = c / dh;
        const float* a = att + ((size_t)h * K + i) * K;
        const float* mm = m ? m + ((size_t)h * K + i) * K : nullptr;
        float s = 0.f;
        for (int j = 0; j < K; ++j) s = fmaf(mm ? a[j] * mm[j] : a[j], V[(size_t)j * Ds + c], s);
        O[idx] = s;
    }
    __syncthreads();
}

// Same computation with Q, K, V and the attention matrix staged in shared memory (rows padded to an
// odd stride: conflict-free for both the per-key and the per-feature access patterns).
static __device__ __noinline__ void mha_core_fwd_smem(const float* Q, const float* Kk, const float* V, float* att, float* O,
                                                      int K, int Ds, int H, float* arena, const float* m) {
    const int dh = Ds / H, tid = threadIdx.x, ld = Ds + 1, ka = K | 1;
    float* sQ = arena; float* sK = sQ + (size_t)K * ld; float* sV = sK + (size_t)K * ld; float* sA = sV + (size_t)K * ld;
    __syncthreads();
    for (int i = tid * 4; i < K * Ds; i += NT * 4) {
        const int r = i / Ds, c = i - r * Ds;
        const float4 a = ld4(Q + i), b = ld4(Kk + i), v = ld4(V + i);
        float* q = sQ + (size_t)r * ld + c; float* k = sK + (size_t)r * ld + c; float* w = sV + (size_t)r * ld + c;
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
        k[0] = b.x; k[1] = b.y; k[2] = b.z; k[3] = b.w;
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __syncthreads();
    for (int idx = tid; idx < H * K * K; idx += NT) {
        const int j = idx % K, i = (idx / K) % K, h = idx / (K * K);
        const float* q = sQ + (size_t)i * ld + h * dh;
        const float* k = sK + (size_t)j * ld + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s = fmaf(q[c], k[c], s);
        sA[(size_t)(h * K + i) * ka + j] = s;
    }
    __syncthreads();
    for (int row = tid; row < H * K; row += NT) {
        float* a = sA + (size_t)row * ka;
        float mx = -INFINITY;
        for (int j = 0; j < K; ++j) mx = fmaxf(mx, a[j]);
        float sum = 0.f;
        for (int j = 0; j < K; ++j) { float e = expf(a[j] - mx); a[j] = e; sum += e; }
        const float inv = 1.0f / sum;
        for (int j = 0; j < K; ++j) a[j] *= inv;
    }
    __syncthreads();
    for (int idx = tid; idx < H * K * K; idx += NT) att[idx] = sA[(size_t)(idx / K) * ka + idx % K];     // saved for backward
    if (m) {                                                  // attn_dropout (transformer.py:44): the product below uses the dropped probabilities
        __syncthreads();
        for (int idx = tid; idx < H * K * K; idx += NT) sA[(size_t)(idx / K) * ka + idx % K] *= m[idx];
        __syncthreads();
    }
    for (int idx = tid; idx < K * Ds; idx += NT) {
        const int c = idx % Ds, i = idx / Ds, h = c / dh;
        const float* a = sA + (size_t)(h * K + i) * ka;
        float s = 0.f;
        for (int j = 0; j < K; ++j) s = fmaf(a[j], sV[(size_t)j * ld + c], s);
        O[idx] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// K2: the whole T x I recurrence of one clip.  grid = B * CN CTAs, cluster CN.
// ---------------------------------------------------------------------------
template <typename TokT, int KMAX, bool MMA>
__global__ void __launch_bounds__(NT, 1) savi_fwd_kernel(const __grid_constant__ FwdArgs a) {
    extern __shared__ float4 smem4[];
    // dynamic smem: [operand buffer A | operand buffer B | arena (token tiles / staging)]
    unsigned char* smem = reinterpret_cast<unsigned char*>(smem4) + 2 * (size_t)a.op_bytes;
    float* arena = reinterpret_cast<float*>(smem);
    const Dims& d = a.d;
    const ParamOff& po = a.po;
    const int tid = threadIdx.x;
    const int CN = d.CN, b = blockIdx.x / CN, rank = blockIdx.x % CN;
    const int K = d.K, Ds = d.Ds, D = d.D, M = d.M, B = d.B, KP = d.KP;
    const int per = ((d.N + CN - 1) / CN + 15) & ~15;
    const int n_lo = min(d.N, rank * per), n_hi = min(d.N, n_lo + per);
    const float* P = a.packed;
    float* fb = reinterpret_cast<float*>(a.saved + a.sl.fbase);
    const TokT* xhat = reinterpret_cast<const TokT*>(a.saved + a.sl.xhat);
    float* cs = a.ws + a.wl.cta + (size_t)blockIdx.x * a.wl.cta_floats;    // this CTA's scratch
    float* h = cs + a.wl.h;
    float* st = cs + a.wl.st;
    float* gi = cs + a.wl.gi;
    float* gh = cs + a.wl.gh;
    const bool lead = (rank == 0);
    const int AF = a.arena_floats;
    constexpr int MT = (KMAX + 15) / 16;
    const bf16* Phi = reinterpret_cast<const bf16*>(P + po.packed_total);      // bf16 hi / lo images of the packed buffer
    const bf16* Plo = Phi + po.packed_total;
    // W_io: offset of the [in][out] fp32 copy (SIMT path), W_oi: offset of the [out][in] copy (tensor-core path)
#define LINP(Y, ldy, X, ldx, W_io, W_oi, bias, Res, ldr, R_, C_, O_, alpha, fl, pre_, out_) \
    lin<MMA, MT>(P, Phi, Plo, Y, ldy, X, ldx, W_io, W_oi, bias, Res, ldr, nullptr, 0, R_, C_, O_, alpha, fl, arena, AF, pre_, out_)
#define LIN(Y, ldy, X, ldx, W_io, W_oi, bias, Res, ldr, R_, C_, O_, alpha, fl) \
    LINP(Y, ldy, X, ldx, W_io, W_oi, bias, Res, ldr, R_, C_, O_, alpha, fl, nullptr, nullptr)
    // operand hand-over buffers (tensor-core path): a producer whose width equals op_width also leaves its
    // result there as bf16 hi/lo, the consuming linear then needs no staging pass
    OpStage opA, opB;
    constexpr int OW = 0;     // operand hand-over measured neutral at C2 and costs registers: compiled out (kept for the record)
    {
        const int ost = OW ? lin_stride(OW) : 0;
        opA.hi = reinterpret_cast<bf16*>(smem4); opA.lo = opA.hi + (size_t)MT * 16 * ost; opA.stride = ost;
        opB.hi = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(smem4) + a.op_bytes); opB.lo = opB.hi + (size_t)MT * 16 * ost; opB.stride = ost;
        if (OW) { opstage_clear(opA, MT * 16); opstage_clear(opB, MT * 16); }
    }
#define OP_IF(w_, buf_) ((OW && (w_) == OW) ? &(buf_) : nullptr)

    if (blockIdx.x == 0 && tid == 0) g_lin_dbg = a.dbg;
    long long ph_last = clock64();
    // slots0 = mu + exp(log_sigma) * noise   (steve.py:56-57)
    for (int i = tid; i < K * Ds; i += NT) {
        const int c = i % Ds;
        h[i] = P[po.slot_mu + c] + expf(P[po.slot_log_sigma + c]) * a.noise[(size_t)b * K * Ds + i];
    }
    __syncthreads();

    for (int t = 0; t < d.T; ++t) {
        const TokT* xh_t = xhat + ((size_t)b * d.T + t) * d.N * D;
        for (int it = 0; it < d.I; ++it) {
            const int64_t s = (int64_t)t * d.I + it;
            // step record: the saved arrays for the cluster leader, a private shadow for the other ranks
            float* r_hp = lead ? frow(fb, a.sl.hp, s, b, B, K, Ds) : cs + a.wl.sh_hp;
            float* r_q = lead ? frow(fb, a.sl.q, s, b, B, K, Ds) : cs + a.wl.sh_q;
            float* r_qk = lead ? frow(fb, a.sl.qk, s, b, B, K, D) : cs + a.wl.sh_qk;
            float* r_ux = lead ? frow(fb, a.sl.ux, s, b, B, K, D) : cs + a.wl.sh_ux;
            float* r_u = lead ? frow(fb, a.sl.u, s, b, B, K, Ds) : cs + a.wl.sh_u;
            float* r_r = lead ? frow(fb, a.sl.r, s, b, B, K, Ds) : cs + a.wl.sh_r;
            float* r_z = lead ? frow(fb, a.sl.z, s, b, B, K, Ds) : cs + a.wl.sh_z;
            float* r_n = lead ? frow(fb, a.sl.n, s, b, B, K, Ds) : cs + a.wl.sh_n;
            float* r_ghn = lead ? frow(fb, a.sl.ghn, s, b, B, K, Ds) : cs + a.wl.sh_ghn;
            float* r_ss = lead ? fb + a.sl.ssum + (s * B + b) * KP : cs + a.wl.sh_ssum;

            SAVI_PH(0);
            cta_copy(r_hp, h, K * Ds);                                                   // slots_prev (steve.py:71)
            cta_ln(st, Ds, h, Ds, P + po.ln_s_w, P + po.ln_s_b, K, Ds, d.ln_eps);      // :72
            SAVI_PH(1);
            LINP(r_q, Ds, st, Ds, po.wq_t, po.wq, nullptr, nullptr, 0, K, Ds, Ds, 1.0f, 0, OP_IF(Ds, opA), OP_IF(Ds, opB));   // :75
            SAVI_PH(2);
            LINP(r_qk, D, r_q, Ds, po.wk, po.wk_t, nullptr, nullptr, 0, K, Ds, D, d.qscale, 0, OP_IF(Ds, opB), nullptr);   // fold Wk and Ds^-1/2 (:61,63)

            SAVI_PH(3);
            float* part = a.ws + a.wl.part + (((size_t)b * 2 + (s & 1)) * CN + rank) * ((size_t)K * D + KP);
            TokT* attn_t = (it == d.I - 1) ? reinterpret_cast<TokT*>(a.attn_out) + ((size_t)b * d.T + t) * d.N * K : nullptr;
            if constexpr (MMA) token_pass_fwd_mma<MT>(d, xh_t, n_lo, n_hi, r_qk, part, attn_t, smem, a.stages, (blockIdx.x == 0) ? a.dbg : nullptr);
            else token_pass_fwd<TokT, KMAX>(d, xh_t, n_lo, n_hi, r_qk, part, attn_t, smem, a.TN);
            SAVI_PH(4);
            __threadfence();
            sync_clip(CN);
            SAVI_PH(5);
            // combine the ranks' partial sums in a fixed order: Ux = (sum A x) / (sum A)   (:82-83)
            {
                const float* pb = a.ws + a.wl.part + (((size_t)b * 2 + (s & 1)) * CN) * ((size_t)K * D + KP);
                const size_t pstride = (size_t)K * D + KP;
                for (int i = tid * 4; i < K * D; i += NT * 4) {           // all loads of a thread are issued before the math
                    const int k = i / D;
                    float4 num = __ldcg(reinterpret_cast<const float4*>(pb + i));
                    float den = __ldcg(pb + (size_t)K * D + k);
                    for (int r = 1; r < CN; ++r) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(pb + r * pstride + i));
                        num.x += t.x; num.y += t.y; num.z += t.z; num.w += t.w;
                        den += __ldcg(pb + r * pstride + (size_t)K * D + k);
                    }
                    const float4 ux = make_float4(num.x / den, num.y / den, num.z / den, num.w / den);
                    st4(r_ux + i, ux);
                    if constexpr (MMA) {
                        if (OW && D == OW) { const int c = i - k * D; opstage_put2(opA, k, c, ux.x, ux.y); opstage_put2(opA, k, c + 2, ux.z, ux.w); }
                    }
                    if (i % D == 0) r_ss[k] = den;
                }
                __syncthreads();
            }
            SAVI_PH(6);
            LINP(r_u, Ds, r_ux, D, po.wv_t, po.wv, nullptr, nullptr, 0, K, D, Ds, 1.0f, 0, OP_IF(D, opA), OP_IF(Ds, opB));   // updates (:83)
            SAVI_PH(7);
            LINP(gi, 3 * Ds, r_u, Ds, po.wih_t, po.wih, P + po.bih, nullptr, 0, K, Ds, 3 * Ds, 1.0f, 0, OP_IF(Ds, opB), nullptr);   // GRUCell (:87)
            SAVI_PH(8);
            LIN(gh, 3 * Ds, r_hp, Ds, po.whh_t, po.whh, P + po.bhh, nullptr, 0, K, Ds, 3 * Ds, 1.0f, 0);
            SAVI_PH(9);
            const bool mlp = (it < d.I - 1);
            float* r_hg = nullptr; float* r_a = nullptr;
            if (mlp) {
                const int64_t sm = (int64_t)t * (d.I - 1) + it;
                r_hg = lead ? frow(fb, a.sl.hg, sm, b, B, K, Ds) : cs + a.wl.sh_hg;
                r_a = lead ? frow(fb, a.sl.a, sm, b, B, K, M) : cs + a.wl.sh_a;
            }
            for (int i = tid * 4; i < K * Ds; i += NT * 4) {               // 4 features per thread, loads first
                const int k = i / Ds, c = i - k * Ds;
                const float* gik = gi + (size_t)k * 3 * Ds + c; const float* ghk = gh + (size_t)k * 3 * Ds + c;
                const float4 ir = ld4(gik), iz = ld4(gik + Ds), in_ = ld4(gik + 2 * Ds);
                const float4 hr = ld4(ghk), hz = ld4(ghk + Ds), hn4 = ld4(ghk + 2 * Ds), hp4 = ld4(r_hp + i);
                const float air[4] = {ir.x, ir.y, ir.z, ir.w}, aiz[4] = {iz.x, iz.y, iz.z, iz.w}, ain[4] = {in_.x, in_.y, in_.z, in_.w};
                const float ahr[4] = {hr.x, hr.y, hr.z, hr.w}, ahz[4] = {hz.x, hz.y, hz.z, hz.w}, ahn[4] = {hn4.x, hn4.y, hn4.z, hn4.w};
                const float ahp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
                float vr[4], vz[4], vn[4], vh[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vr[e] = sigmoidf_(air[e] + ahr[e]);
                    vz[e] = sigmoidf_(aiz[e] + ahz[e]);
                    vn[e] = tanhf(ain[e] + vr[e] * ahn[e]);
                    vh[e] = (1.0f - vz[e]) * vn[e] + vz[e] * ahp[e];
                }
                st4(r_r + i, make_float4(vr[0], vr[1], vr[2], vr[3])); st4(r_z + i, make_float4(vz[0], vz[1], vz[2], vz[3]));
                st4(r_n + i, make_float4(vn[0], vn[1], vn[2], vn[3])); st4(r_ghn + i, hn4);
                st4(h + i, make_float4(vh[0], vh[1], vh[2], vh[3]));
                if (mlp) st4(r_hg + i, make_float4(vh[0], vh[1], vh[2], vh[3]));
            }
            __syncthreads();
            SAVI_PH(10);
            if (mlp) {                                                                    // residual MLP (:92-93)
                cta_ln(st, Ds, h, Ds, P + po.ln_m_w, P + po.ln_m_b, K, Ds, d.ln_eps);
                SAVI_PH(11);
                LINP(r_a, M, st, Ds, po.w1_t, po.w1, P + po.b1, nullptr, 0, K, Ds, M, 1.0f, LIN_RELU, OP_IF(Ds, opA), OP_IF(M, opB));
                SAVI_PH(12);
                LINP(h, Ds, r_a, M, po.w2_t, po.w2, P + po.b2, r_hg, Ds, K, M, Ds, 1.0f, 0, OP_IF(M, opB), nullptr);
            }
        }
        SAVI_PH(13);
        if (lead) cta_copy(a.slots_out + ((size_t)b * d.T + t) * K * Ds, h, K * Ds);     // collect (:96-97)
        if (t < d.T - 1) {
            // predictor (:100).  The reference also evaluates it after the last frame and discards the result.
            if (lead) cta_copy(fb + a.sl.px0 + ((size_t)t * B + b) * K * Ds, h, K * Ds);
            __syncthreads();
            const float hscale = 1.0f / sqrtf((float)(Ds / d.heads));
            const float* x = h;
            for (int j = 0; j < d.blocks; ++j) {
                const int64_t f = (int64_t)j * (d.T - 1) + t;        // block j's rows are contiguous
                const BlockOff& bo = po.blk[j];
                const BlockOffT& bt = po.blkt[j];
                float* p_y = lead ? frow(fb, a.sl.py, f, b, B, K, Ds) : cs + a.wl.sh_py;
                float* p_q = lead ? frow(fb, a.sl.pq, f, b, B, K, Ds) : cs + a.wl.sh_pq;
                float* p_k = lead ? frow(fb, a.sl.pk, f, b, B, K, Ds) : cs + a.wl.sh_pk;
                float* p_v = lead ? frow(fb, a.sl.pv, f, b, B, K, Ds) : cs + a.wl.sh_pv;
                float* p_o = lead ? frow(fb, a.sl.po, f, b, B, K, Ds) : cs + a.wl.sh_po;
                float* p_x1 = lead ? frow(fb, a.sl.px1, f, b, B, K, Ds) : cs + a.wl.sh_px1;
                float* p_l2 = lead ? frow(fb, a.sl.pl2, f, b, B, K, Ds) : cs + a.wl.sh_pl2;
                float* p_f = lead ? frow(fb, a.sl.pf, f, b, B, K, 4 * Ds) : cs + a.wl.sh_pf;
                float* p_x2 = lead ? frow(fb, a.sl.px2, f, b, B, K, Ds) : cs + a.wl.sh_px2;
                float* p_att = lead ? fb + a.sl.patt + (f * B + b) * ((int64_t)d.heads * K * K) : cs + a.wl.sh_patt;
                SAVI_PH(14);
                cta_ln(p_y, Ds, x, Ds, P + bo.ln1_w, P + bo.ln1_b, K, Ds, d.ln_eps);
                SAVI_PH(15);
                LINP(p_q, Ds, p_y, Ds, bt.pq_t, bo.pq, nullptr, nullptr, 0, K, Ds, Ds, hscale, 0, OP_IF(Ds, opA), nullptr);
                LINP(p_k, Ds, p_y, Ds, bt.pk_t, bo.pk, nullptr, nullptr, 0, K, Ds, Ds, 1.0f, 0, OP_IF(Ds, opA), nullptr);
                LINP(p_v, Ds, p_y, Ds, bt.pv_t, bo.pv, nullptr, nullptr, 0, K, Ds, Ds, 1.0f, 0, OP_IF(Ds, opA), nullptr);
                SAVI_PH(16);
                // training-mode dropout masks of this block evaluation (nullptr otherwise): savi_args.h, DropLayout
                const DropLayout dl = savi_dropout_layout(d);
                const float* m_att = a.drop ? a.drop + dl.att + (f * B + b) * ((int64_t)d.heads * K * K) : nullptr;
                const float* m_out = a.drop ? a.drop + dl.out + (f * B + b) * ((int64_t)K * Ds) : nullptr;
                const float* m_ffn = a.drop ? a.drop + dl.ffn + (f * B + b) * ((int64_t)K * Ds) : nullptr;
                if (3 * K * (Ds + 1) + d.heads * K * (K | 1) <= AF) mha_core_fwd_smem(p_q, p_k, p_v, p_att, p_o, K, Ds, d.heads, arena, m_att);
                else mha_core_fwd(p_q, p_k, p_v, p_att, p_o, K, Ds, d.heads, m_att);
                SAVI_PH(17);
                // first block adds the residual to the NORMALISED input (transformer.py:75-78)
                const float* res1 = (j == 0) ? p_y : x;
                if (m_out) {                                  // x1 = residual + output_dropout(proj_o(O))   (transformer.py:48)
                    LIN(p_x1, Ds, p_o, Ds, bt.po_t, bo.po, nullptr, nullptr, 0, K, Ds, Ds, 1.0f, 0);
                    __syncthreads();
                    for (int i = tid; i < K * Ds; i += NT) p_x1[i] = res1[i] + m_out[i] * p_x1[i];
                    __syncthreads();
                } else
                LIN(p_x1, Ds, p_o, Ds, bt.po_t, bo.po, nullptr, res1, Ds, K, Ds, Ds, 1.0f, 0);
                SAVI_PH(18);
                cta_ln(p_l2, Ds, p_x1, Ds, P + bo.ln2_w, P + bo.ln2_b, K, Ds, d.ln_eps);
                SAVI_PH(19);
                LINP(p_f, 4 * Ds, p_l2, Ds, bt.f1_t, bo.f1, P + bo.f1b, nullptr, 0, K, Ds, 4 * Ds, 1.0f, LIN_RELU, OP_IF(Ds, opA), nullptr);
                SAVI_PH(20);
                if (m_ffn) {                                  // x2 = x1 + Dropout(ffn.2(f) + b2)   (transformer.py:68)
                    LIN(p_x2, Ds, p_f, 4 * Ds, bt.f2_t, bo.f2, P + bo.f2b, nullptr, 0, K, 4 * Ds, Ds, 1.0f, 0);
                    __syncthreads();
                    for (int i = tid; i < K * Ds; i += NT) p_x2[i] = p_x1[i] + m_ffn[i] * p_x2[i];
                    __syncthreads();
                } else
                LIN(p_x2, Ds, p_f, 4 * Ds, bt.f2_t, bo.f2, P + bo.f2b, p_x1, Ds, K, 4 * Ds, Ds, 1.0f, 0);
                x = p_x2;
            }
            SAVI_PH(21);
            cta_ln(st, Ds, x, Ds, P + po.lnf_w, P + po.lnf_b, K, Ds, d.ln_eps);
            cta_copy(h, st, K * Ds);
            __syncthreads();
        }
    }
#undef LIN
#undef LINP
#undef OP_IF
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
template <typename TokT>
static cudaError_t launch_ln_fwd(const FwdArgs& a, const void* inputs, cudaStream_t st) {
    const int64_t rows = (int64_t)a.d.B * a.d.T * a.d.N;
    if constexpr (sizeof(TokT) == 2) {
        if (a.d.umma && a.d.D == 128) {
            ln_tokens_fwd_img128_kernel<<<(unsigned)(a.d.B * a.d.T * a.d.NTILE), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(inputs),
                reinterpret_cast<float2*>(a.saved + a.sl.stats), a.packed + a.po.ln_in_w, a.packed + a.po.ln_in_b, a.d.ln_eps,
                a.saved + a.sl.ximg, a.d.N, a.d.NTILE);
            return cudaGetLastError();
        }
    }
    int grid = (int)((rows + NW - 1) / NW);
    if (grid > 148 * 16) grid = 148 * 16;
    ln_tokens_fwd_kernel<TokT><<<grid, NT, 0, st>>>(
        reinterpret_cast<const TokT*>(inputs),
        a.d.umma ? nullptr : reinterpret_cast<TokT*>(a.saved + a.sl.xhat),      // the tcgen05 kernels read only the blocked image
        reinterpret_cast<float2*>(a.saved + a.sl.stats), a.packed + a.po.ln_in_w, a.packed + a.po.ln_in_b, rows, a.d.D,
        a.d.ln_eps, a.d.umma ? a.saved + a.sl.ximg : nullptr, a.d.N, a.d.NTILE);
    return cudaGetLastError();
}

template <typename TokT, int KMAX, bool MMA>
static cudaError_t launch_fwd_t(const FwdArgs& a, cudaStream_t st) {
    auto kern = savi_fwd_kernel<TokT, KMAX, MMA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.d.B * a.d.CN);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = a.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.d.CN; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <typename TokT>
static cudaError_t launch_fwd_k(const FwdArgs& a, cudaStream_t st) {
    const int K = a.d.K;
    if constexpr (sizeof(TokT) == 2) {
        if (a.d.mma) {                      // tensor-core path: KMAX only selects the number of 16-slot m-tiles
            if (K <= 16) return launch_fwd_t<TokT, 16, true>(a, st);
            if (K <= 32) return launch_fwd_t<TokT, 32, true>(a, st);
            return launch_fwd_t<TokT, 64, true>(a, st);
        }
    }
    if (K <= 8) return launch_fwd_t<TokT, 8, false>(a, st);
    if (K <= 16) return launch_fwd_t<TokT, 16, false>(a, st);
    if (K <= 24) return launch_fwd_t<TokT, 24, false>(a, st);
    if (K <= 32) return launch_fwd_t<TokT, 32, false>(a, st);
    return launch_fwd_t<TokT, 64, false>(a, st);
}

#ifndef SAVI_TOK
#error "compile with -DSAVI_TOK=float -DSAVI_SUFFIX=f32 (or __nv_bfloat16 / bf16)"
#endif
#define SAVI_CAT2(a, b) a##b
#define SAVI_CAT(a, b) SAVI_CAT2(a, b)

cudaError_t SAVI_CAT(savi_launch_forward_, SAVI_SUFFIX)(const FwdArgs& a, const void* inputs, cudaStream_t st, int* launches) {
    savi_prof_begin(1, st);
    cudaError_t e = launch_ln_fwd<SAVI_TOK>(a, inputs, st);
    savi_prof_end(1, st);
    if (e != cudaSuccess) return e;
    savi_prof_begin(2, st);
    if (a.d.umma) {
        WImg wi;
        savi_wimg_layout(a.d.D, a.d.Ds, a.d.M, a.d.blocks, wi);
        e = savi_launch_fwd_umma(a, reinterpret_cast<const unsigned char*>(a.packed) + savi_wimg_base(a.po.packed_total), wi, st);
    } else
    e = launch_fwd_k<SAVI_TOK>(a, st);
    savi_prof_end(2, st);
    *launches += 2;
    return e;
}
