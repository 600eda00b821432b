// Device-side building blocks shared by the forward and backward clip kernels.
//
// Execution model: one CTA (or one thread-block cluster of CN CTAs) owns one clip
// for the whole T x I recurrence, so the sequential slot-state chain needs no
// grid-wide synchronisation and no kernel launches between steps.  Slot-side
// tensors are [K, C] fp32 with K <= 64 rows; they live in global memory (they are
// what backward needs anyway) and are staged through shared memory by the CTA.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cooperative_groups.h>
#include "savi_layout.h"

namespace cg = cooperative_groups;

constexpr int NT = 512;            // threads per CTA (16 warps: the recurrence is latency-bound, warps are the only way to hide it)
constexpr int NW = NT / 32;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- token element access (fp32 or bf16 storage, fp32 math) -----------------
template <typename T> struct Tok;
template <> struct Tok<float> {
    static constexpr int VEC = 4;                      // elements per 16-byte chunk
    __device__ static __forceinline__ void load(const float* p, float* v) {
        float4 t = ld4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static __forceinline__ float4 load4(const float* p) { return ld4(p); }
    __device__ static __forceinline__ void store4(float* p, float4 v) { st4(p, v); }
    __device__ static __forceinline__ float to_f(float x) { return x; }
    __device__ static __forceinline__ float from_f(float x) { return x; }
};
template <> struct Tok<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void load(const __nv_bfloat16* p, float* v) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
    __device__ static __forceinline__ float4 load4(const __nv_bfloat16* p) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
        float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    __device__ static __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 t; t.x = *reinterpret_cast<unsigned*>(&a); t.y = *reinterpret_cast<unsigned*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
    __device__ static __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
    __device__ static __forceinline__ __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Row stride (floats) of the [tokens, K] coefficient tiles: a multiple of 4 whose
// quarter is odd, so both per-token float4 row access and float4 broadcast reads
// along K are bank-conflict free.
__host__ __device__ __forceinline__ int coef_stride(int KP) { return ((KP / 4) & 1) ? KP : KP + 4; }
// Row stride (bytes) of a token tile in shared memory: +16 B skews rows across banks.
__host__ __device__ __forceinline__ int tile_stride_bytes(int D, int tok_bytes) { return D * tok_bytes + 16; }

__device__ __forceinline__ void sync_clip(int CN) {
    if (CN > 1) cg::this_cluster().sync();   // barrier.cluster arrive.release / wait.acquire: orders global writes too
    else __syncthreads();
}

// ---------------------------------------------------------------------------
// Y[r][o] = act( Res[r][o] + bias[o] + alpha * sum_c X[r][c] * W[c][o] ),  r < R, o < O
//   X   : global [R, ldx] fp32 (written earlier by this CTA), staged through `arena`
//   W   : global [C, ldw] fp32, "in x out" (out contiguous) -> coalesced across threads
//   Res : optional residual (may alias Y), Mask : optional, output zeroed where Mask <= 0
// Requires C % 4 == 0, O % 2 == 0, ldx % 4 == 0.
// ---------------------------------------------------------------------------
constexpr int LIN_RELU = 1;

static __device__ __noinline__ void cta_linear(float* Y, int ldy, const float* X, int ldx, const float* __restrict__ W, int ldw,
                           const float* __restrict__ bias, const float* Res, int ldr, const float* Mask, int ldm,
                           int R, int C, int O, float alpha, int flags, float* arena, int arena_floats) {
    const int tid = threadIdx.x;
    int rc_max = arena_floats / C;
    if (rc_max >= 8) rc_max &= ~7;
    if (rc_max > R) rc_max = R;
    const int nop = O >> 1;
    for (int r0 = 0; r0 < R; r0 += rc_max) {
        const int rc = min(rc_max, R - r0);
        __syncthreads();
        const int c4n = C >> 2;
        for (int idx = tid; idx < rc * c4n; idx += NT) {
            int r = idx / c4n, c4 = idx - r * c4n;
            st4(arena + (size_t)idx * 4, ld4(X + (size_t)(r0 + r) * ldx + c4 * 4));
        }
        __syncthreads();
        const int nrb = (rc + 7) >> 3;
        for (int item = tid; item < nrb * nop; item += NT) {
            const int rb = item / nop, o = (item - rb * nop) * 2;
            const int nr = min(8, rc - rb * 8);
            const float* xb = arena + (size_t)rb * 8 * C;
            float a0[8], a1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
            const float* wp = W + o;
            for (int c = 0; c < C; c += 4) {
                const float2 w0 = __ldg(reinterpret_cast<const float2*>(wp + (size_t)(c + 0) * ldw));
                const float2 w1 = __ldg(reinterpret_cast<const float2*>(wp + (size_t)(c + 1) * ldw));
                const float2 w2 = __ldg(reinterpret_cast<const float2*>(wp + (size_t)(c + 2) * ldw));
                const float2 w3 = __ldg(reinterpret_cast<const float2*>(wp + (size_t)(c + 3) * ldw));
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < nr) {
                        const float4 x = ld4(xb + i * C + c);
                        a0[i] = fmaf(x.x, w0.x, a0[i]); a1[i] = fmaf(x.x, w0.y, a1[i]);
                        a0[i] = fmaf(x.y, w1.x, a0[i]); a1[i] = fmaf(x.y, w1.y, a1[i]);
                        a0[i] = fmaf(x.z, w2.x, a0[i]); a1[i] = fmaf(x.z, w2.y, a1[i]);
                        a0[i] = fmaf(x.w, w3.x, a0[i]); a1[i] = fmaf(x.w, w3.y, a1[i]);
                    }
                }
            }
            const float b0 = bias ? __ldg(bias + o) : 0.f, b1 = bias ? __ldg(bias + o + 1) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < nr) {
                    const int r = r0 + rb * 8 + i;
                    float y0 = alpha * a0[i] + b0, y1 = alpha * a1[i] + b1;
                    if (Res) { y0 += Res[(size_t)r * ldr + o]; y1 += Res[(size_t)r * ldr + o + 1]; }
                    if (flags & LIN_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                    if (Mask) {
                        if (!(Mask[(size_t)r * ldm + o] > 0.f)) y0 = 0.f;
                        if (!(Mask[(size_t)r * ldm + o + 1] > 0.f)) y1 = 0.f;
                    }
                    *reinterpret_cast<float2*>(Y + (size_t)r * ldy + o) = make_float2(y0, y1);
                }
            }
        }
    }
    __syncthreads();
}

// LayerNorm over the last dim, one warp per row (torch semantics: biased variance).
// Every element is loaded ONCE (all loads of a row are issued back to back, then reduced from
// registers): the rows live in global memory and each dependent re-read would cost an L2 round trip.
template <int VPT>   // values per lane: C <= 32 * VPT
static __device__ __forceinline__ void cta_ln_t(float* Y, int ldy, const float* X, int ldx, const float* __restrict__ g,
                                                 const float* __restrict__ b, int R, int C, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < R; r += NW) {
        const float* x = X + (size_t)r * ldx;
        float v[VPT];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) { const int c = lane + 32 * j; v[j] = (c < C) ? x[c] : 0.f; s += v[j]; }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) { const int c = lane + 32 * j; const float t = (c < C) ? v[j] - mean : 0.f; q = fmaf(t, t, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int c = lane + 32 * j;
            if (c < C) Y[(size_t)r * ldy + c] = (v[j] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
        }
    }
    __syncthreads();
}
static __device__ __noinline__ void cta_ln(float* Y, int ldy, const float* X, int ldx, const float* __restrict__ g,
                                           const float* __restrict__ b, int R, int C, float eps) {
    if (C <= 128) cta_ln_t<4>(Y, ldy, X, ldx, g, b, R, C, eps);
    else if (C <= 256) cta_ln_t<8>(Y, ldy, X, ldx, g, b, R, C, eps);
    else cta_ln_t<16>(Y, ldy, X, ldx, g, b, R, C, eps);
}

// dX[r] = (Res ? Res[r] : 0) + LayerNorm backward of dY through X;  optionally
// accumulates d gamma / d beta into global (atomics, one add per warp and column).  C <= 512.
template <int VPT>
static __device__ __forceinline__ void cta_ln_bwd_t(float* dX, int lddx, const float* Res, int ldr, const float* dY, int lddy,
                                                     const float* X, int ldx, const float* __restrict__ g, float* dg_glob, float* db_glob,
                                                     int R, int C, float eps, bool do_param_grads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ag[VPT], ab[VPT], gm[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) { ag[j] = 0.f; ab[j] = 0.f; const int c = lane + 32 * j; gm[j] = (c < C) ? __ldg(g + c) : 0.f; }
    for (int r = warp; r < R; r += NW) {
        const float* x = X + (size_t)r * ldx;
        const float* dy = dY + (size_t)r * lddy;
        float xv[VPT], dv[VPT], rv[VPT];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {                      // all loads of the row first
            const int c = lane + 32 * j;
            xv[j] = (c < C) ? x[c] : 0.f;
            dv[j] = (c < C) ? dy[c] : 0.f;
            rv[j] = (Res && c < C) ? Res[(size_t)r * ldr + c] : 0.f;
            s += xv[j];
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) { const int c = lane + 32 * j; const float t = (c < C) ? xv[j] - mean : 0.f; q = fmaf(t, t, q); }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int c = lane + 32 * j;
            const float z = (c < C) ? (xv[j] - mean) * rstd : 0.f;
            const float dz = dv[j] * gm[j];
            s1 += dz; s2 = fmaf(dz, z, s2);
            ag[j] = fmaf(dv[j], z, ag[j]); ab[j] += dv[j];
            xv[j] = z; dv[j] = dz;
        }
        s1 = warp_sum(s1) / (float)C; s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int c = lane + 32 * j;
            if (c < C) dX[(size_t)r * lddx + c] = rstd * (dv[j] - s1 - xv[j] * s2) + rv[j];
        }
    }
    if (do_param_grads) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const int c = lane + 32 * j;
            if (c < C) { atomicAdd(dg_glob + c, ag[j]); atomicAdd(db_glob + c, ab[j]); }
        }
    }
    __syncthreads();
}
static __device__ __noinline__ void cta_ln_bwd(float* dX, int lddx, const float* Res, int ldr, const float* dY, int lddy,
                                               const float* X, int ldx, const float* __restrict__ g, float* dg_glob, float* db_glob,
                                               int R, int C, float eps, bool do_param_grads) {
    if (C <= 128) cta_ln_bwd_t<4>(dX, lddx, Res, ldr, dY, lddy, X, ldx, g, dg_glob, db_glob, R, C, eps, do_param_grads);
    else if (C <= 256) cta_ln_bwd_t<8>(dX, lddx, Res, ldr, dY, lddy, X, ldx, g, dg_glob, db_glob, R, C, eps, do_param_grads);
    else cta_ln_bwd_t<16>(dX, lddx, Res, ldr, dY, lddy, X, ldx, g, dg_glob, db_glob, R, C, eps, do_param_grads);
}

// dst[o] += sum_r X[r][o]   (bias gradients), one thread per column.
static __device__ __noinline__ void cta_colsum_atomic(float* dst, const float* X, int ldx, int R, int O) {
    for (int o = threadIdx.x; o < O; o += NT) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int r = 0;
        for (; r + 3 < R; r += 4) {
            const float a0 = X[(size_t)r * ldx + o], a1 = X[(size_t)(r + 1) * ldx + o];
            const float a2 = X[(size_t)(r + 2) * ldx + o], a3 = X[(size_t)(r + 3) * ldx + o];
            s0 += a0; s1 += a1; s2 += a2; s3 += a3;
        }
        for (; r < R; ++r) s0 += X[(size_t)r * ldx + o];
        atomicAdd(dst + o, (s0 + s1) + (s2 + s3));
    }
}

__device__ __forceinline__ void cta_copy(float* dst, const float* src, int n) {   // n % 4 == 0, 16B aligned
    for (int i = threadIdx.x * 4; i < n; i += NT * 4) st4(dst + i, ld4(src + i));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------
// Token tile staging: rows [n0, n0+tn) of a [N, D] token matrix -> shared memory
// ---------------------------------------------------------------------------
template <typename TokT>
__device__ __forceinline__ void load_token_tile(unsigned char* xs, int xstride, const TokT* __restrict__ g,
                                                int n0, int tn, int D) {
    const int chunks = D * (int)sizeof(TokT) / 16;
    const unsigned char* gb = reinterpret_cast<const unsigned char*>(g + (size_t)n0 * D);
    const size_t grow = (size_t)D * sizeof(TokT);
    for (int idx = threadIdx.x; idx < tn * chunks; idx += NT) {
        int r = idx / chunks, c = idx - r * chunks;
        cp_async16(xs + (size_t)r * xstride + c * 16, gb + r * grow + c * 16);
    }
    cp_async_wait_all();
}

// acc_s[k][d] += sum_{n<tn} coef[n][k] * x[n][d]     (k < KP, d < D)
// ssum[k]    += sum_n coef[n][k]                      (when ssum != nullptr)
// Each thread owns a 4 (slots) x 4 (features) register tile.
template <typename TokT>
__device__ __forceinline__ void tile_outer_accum(float* acc_s, float* ssum, const float* coef, int cstride,
                                                 const unsigned char* xs, int xstride, int tn, int KP, int D) {
    const int dgn = D >> 2, items = (KP >> 2) * dgn;
    for (int item = threadIdx.x; item < items; item += NT) {
        const int kg = item / dgn, dg = item - kg * dgn;
        float a[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
        float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* cp = coef + kg * 4;
        const unsigned char* xp = xs + (size_t)dg * 4 * sizeof(TokT);
#pragma unroll 4
        for (int n = 0; n < tn; ++n) {
            const float4 c = ld4(cp + n * cstride);
            const float4 x = Tok<TokT>::load4(reinterpret_cast<const TokT*>(xp + (size_t)n * xstride));
            a[0][0] = fmaf(c.x, x.x, a[0][0]); a[0][1] = fmaf(c.x, x.y, a[0][1]); a[0][2] = fmaf(c.x, x.z, a[0][2]); a[0][3] = fmaf(c.x, x.w, a[0][3]);
            a[1][0] = fmaf(c.y, x.x, a[1][0]); a[1][1] = fmaf(c.y, x.y, a[1][1]); a[1][2] = fmaf(c.y, x.z, a[1][2]); a[1][3] = fmaf(c.y, x.w, a[1][3]);
            a[2][0] = fmaf(c.z, x.x, a[2][0]); a[2][1] = fmaf(c.z, x.y, a[2][1]); a[2][2] = fmaf(c.z, x.z, a[2][2]); a[2][3] = fmaf(c.z, x.w, a[2][3]);
            a[3][0] = fmaf(c.w, x.x, a[3][0]); a[3][1] = fmaf(c.w, x.y, a[3][1]); a[3][2] = fmaf(c.w, x.z, a[3][2]); a[3][3] = fmaf(c.w, x.w, a[3][3]);
            if (dg == 0) { sa.x += c.x; sa.y += c.y; sa.z += c.z; sa.w += c.w; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* o = acc_s + (size_t)(kg * 4 + i) * D + dg * 4;
            float4 t = ld4(o);
            t.x += a[i][0]; t.y += a[i][1]; t.z += a[i][2]; t.w += a[i][3];
            st4(o, t);
        }
        if (ssum && dg == 0) { ssum[kg * 4 + 0] += sa.x; ssum[kg * 4 + 1] += sa.y; ssum[kg * 4 + 2] += sa.z; ssum[kg * 4 + 3] += sa.w; }
    }
}

// Field-array row pointer: rows of one (step, clip) block
__device__ __forceinline__ float* frow(float* fbase, int64_t off, int64_t step, int b, int B, int K, int width) {
    return fbase + off + ((step * B + b) * (int64_t)K) * width;
}
