// Tensor-core (mma.sync m16n8k16, bf16 x bf16 -> fp32) building blocks of the bf16 mode.
//
// Why mma.sync and not tcgen05 here: every contraction on this path is skinny — one side is the
// K <= 64 slot rows of ONE clip (M = 16..64), the tile lives for a few hundred cycles, and a
// softmax over the slot axis sits between the two token-side products.  tcgen05 needs M >= 64/128
// per instruction and a TMEM round trip around the softmax; the warp-level fragments of mma.sync
// keep a token's slot column inside a warp, so the softmax is three shuffles.
//
// Accuracy: operands that carry fp32 state (slot activations, weights, qk, dUx) are split into
// bf16 hi + lo parts and multiplied with three MMAs (hi*hi + hi*lo + lo*hi, error ~2^-16), because
// this recurrence amplifies operand rounding ~10-50x (rounding the weights to bf16 alone moves the
// slots by 3e-2 at BASELINE config 1 — measured on the oracle).  Only the token stream xhat, the
// attention weights and the staged backward coefficients are single bf16.
#pragma once
#include "savi_dev.cuh"

typedef __nv_bfloat16 bf16;

// development aid: sub-phase cycle counters of cta_linear_mma (slots 58..62 of the phase buffer), CTA 0 only
__device__ long long* g_lin_dbg = nullptr;
#define LIN_PH(id) do { if (ldbg) { if (threadIdx.x == 0) { long long t_ = clock64(); \
    atomicAdd(reinterpret_cast<unsigned long long*>(ldbg + (id)), (unsigned long long)(t_ - lt_)); lt_ = t_; } } } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// Shared-memory row stride (in bf16 elements) for a [rows, C] operand read as 16-byte chunks by
// quarter-warps that span two rows (lanes g, g+1 x 4 chunks): stride bytes = 64 (mod 128).
__host__ __device__ __forceinline__ int lin_stride(int C) {
    const int cp = (C + 31) & ~31;                 // contraction padded to the 32-column MMA chunk
    return ((cp * 2) % 128 == 64) ? cp : cp + 32;
}
// bytes of staging arena cta_linear_mma needs for an [R, C] operand (hi + lo)
__host__ __device__ __forceinline__ size_t lin_mma_smem(int MT, int C) { return (size_t)2 * MT * 16 * lin_stride(C) * 2; }

// Optional second destination of a linear layer / LayerNorm: the result is also written, already
// split into bf16 hi/lo, into a shared-memory operand buffer, so the NEXT tensor-core linear finds
// its A operand in place and skips the global round trip + staging pass.
struct OpStage {
    bf16* hi; bf16* lo; int stride;      // [rows][stride] bf16, stride in elements
};
__device__ __forceinline__ void opstage_put2(const OpStage& s, int r, int c, float y0, float y1) {
    bf16 h0, h1, l0, l1;
    split_bf16(y0, h0, l0); split_bf16(y1, h1, l1);
    *reinterpret_cast<__nv_bfloat162*>(s.hi + (size_t)r * s.stride + c) = __halves2bfloat162(h0, h1);
    *reinterpret_cast<__nv_bfloat162*>(s.lo + (size_t)r * s.stride + c) = __halves2bfloat162(l0, l1);
}
// rows [R, 16*MT) and columns [C, stride) of an operand buffer must be zero: done once per kernel
__device__ __forceinline__ void opstage_clear(const OpStage& s, int rows) {
    uint32_t* a = reinterpret_cast<uint32_t*>(s.hi); uint32_t* b = reinterpret_cast<uint32_t*>(s.lo);
    for (int i = threadIdx.x; i < rows * s.stride / 2; i += NT) { a[i] = 0u; b[i] = 0u; }
}

// ---------------------------------------------------------------------------
// Y[r][o] = act( Res[r][o] + bias[o] + alpha * sum_c X[r][c] * W[o][c] ),  r < R <= 16*MT, o < O
//   X        : global fp32 [R, ldx], staged into shared memory as bf16 hi/lo -- unless `pre` names an
//              operand buffer a previous layer already filled (stride must be lin_stride(C))
//   Whi, Wlo : global bf16 [O, C] (contraction index contiguous), streamed from L2 straight
//              into B fragments with 16-byte loads (k-permuted: both operands use the same
//              lane -> column map, so no ldmatrix is needed on the weight side).  The loads of a
//              warp's first tile are issued BEFORE the operand staging, and every later tile is
//              prefetched into registers while the current one is multiplied.
//   out      : optional operand buffer that also receives the result (for the next layer)
// Requires C % 8 == 0, O % 8 == 0.
// ---------------------------------------------------------------------------
template <int MT>
__device__ __noinline__ void cta_linear_mma(float* Y, int ldy, const float* X, int ldx, const bf16* __restrict__ Whi,
                                            const bf16* __restrict__ Wlo, const float* __restrict__ bias, const float* Res, int ldr,
                                            const float* Mask, int ldm, int R, int C, int O, float alpha, int flags, void* arena,
                                            const OpStage* pre, const OpStage* out) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int cp = (C + 31) & ~31, st = pre ? pre->stride : lin_stride(C);
    long long* ldbg = (blockIdx.x == 0) ? g_lin_dbg : nullptr;
    long long lt_ = clock64();
    const int ntiles = O >> 3, nseg = (cp + 127) >> 7;
    const bf16* xh = pre ? pre->hi : reinterpret_cast<const bf16*>(arena);
    const bf16* xl = pre ? pre->lo : xh + (size_t)MT * 16 * st;

    auto load_b = [&](int nt, int seg, uint4 (&bh)[4], uint4 (&bl)[4]) {
        const int o = nt * 8 + g;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            const int col = seg * 128 + ch * 32 + q * 8;
            if (col < C) { bh[ch] = ldg128(Whi + (size_t)o * C + col); bl[ch] = ldg128(Wlo + (size_t)o * C + col); }
            else { bh[ch] = make_uint4(0u, 0u, 0u, 0u); bl[ch] = bh[ch]; }
        }
    };
    if (!pre) {
        __syncthreads();
        bf16* wh = reinterpret_cast<bf16*>(arena);
        bf16* wl = wh + (size_t)MT * 16 * st;
        for (int idx = tid; idx < MT * 16 * (cp >> 2); idx += NT) {
            const int r = idx / (cp >> 2), c = (idx - r * (cp >> 2)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < R && c < C) v = ld4(X + (size_t)r * ldx + c);
            bf16 h0, h1, h2, h3, l0, l1, l2, l3;
            split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
            __nv_bfloat162 hh0 = __halves2bfloat162(h0, h1), hh1 = __halves2bfloat162(h2, h3);
            __nv_bfloat162 ll0 = __halves2bfloat162(l0, l1), ll1 = __halves2bfloat162(l2, l3);
            uint2 ph, pl;
            ph.x = *reinterpret_cast<uint32_t*>(&hh0); ph.y = *reinterpret_cast<uint32_t*>(&hh1);
            pl.x = *reinterpret_cast<uint32_t*>(&ll0); pl.y = *reinterpret_cast<uint32_t*>(&ll1);
            *reinterpret_cast<uint2*>(wh + (size_t)r * st + c) = ph;
            *reinterpret_cast<uint2*>(wl + (size_t)r * st + c) = pl;
        }
    }
    LIN_PH(58);
    __syncthreads();
    LIN_PH(59);
    constexpr int NG = 1;
    for (int nt0 = warp * NG; nt0 < ntiles; nt0 += NW * NG) {
        float acc[NG][MT][4];
#pragma unroll
        for (int n = 0; n < NG; ++n)
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[n][m][e] = 0.f;
        for (int cseg = 0; cseg < cp; cseg += 128) {
            uint4 bh[4][NG], bl[4][NG];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int col = cseg + ch * 32 + q * 8;
#pragma unroll
                for (int n = 0; n < NG; ++n) {
                    const int o = (nt0 + n) * 8 + g;
                    if (col < C && nt0 + n < ntiles) {
                        bh[ch][n] = ldg128(Whi + (size_t)o * C + col);
                        bl[ch][n] = ldg128(Wlo + (size_t)o * C + col);
                    } else {
                        bh[ch][n] = make_uint4(0u, 0u, 0u, 0u); bl[ch][n] = bh[ch][n];
                    }
                }
            }
            LIN_PH(60);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int col = cseg + ch * 32 + q * 8;
                if (cseg + ch * 32 < cp) {
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const uint4 ah0 = *reinterpret_cast<const uint4*>(xh + (size_t)(m * 16 + g) * st + col);
                        const uint4 ah1 = *reinterpret_cast<const uint4*>(xh + (size_t)(m * 16 + g + 8) * st + col);
                        const uint4 al0 = *reinterpret_cast<const uint4*>(xl + (size_t)(m * 16 + g) * st + col);
                        const uint4 al1 = *reinterpret_cast<const uint4*>(xl + (size_t)(m * 16 + g + 8) * st + col);
#pragma unroll
                        for (int n = 0; n < NG; ++n) {
                            // physical columns col+0..3 -> first k16 MMA, col+4..7 -> second (same map for A and B)
                            mma16816(acc[n][m], ah0.x, ah1.x, ah0.y, ah1.y, bh[ch][n].x, bh[ch][n].y);
                            mma16816(acc[n][m], ah0.z, ah1.z, ah0.w, ah1.w, bh[ch][n].z, bh[ch][n].w);
                            mma16816(acc[n][m], ah0.x, ah1.x, ah0.y, ah1.y, bl[ch][n].x, bl[ch][n].y);
                            mma16816(acc[n][m], ah0.z, ah1.z, ah0.w, ah1.w, bl[ch][n].z, bl[ch][n].w);
                            mma16816(acc[n][m], al0.x, al1.x, al0.y, al1.y, bh[ch][n].x, bh[ch][n].y);
                            mma16816(acc[n][m], al0.z, al1.z, al0.w, al1.w, bh[ch][n].z, bh[ch][n].w);
                        }
                    }
                }
            }
        }
        LIN_PH(61);
#pragma unroll
        for (int n = 0; n < NG; ++n) {
            if (nt0 + n < ntiles) {
                const int o = (nt0 + n) * 8 + q * 2;
                const float b0 = bias ? __ldg(bias + o) : 0.f, b1 = bias ? __ldg(bias + o + 1) : 0.f;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
#pragma unroll
                    for (int hrow = 0; hrow < 2; ++hrow) {
                        const int r = m * 16 + g + hrow * 8;
                        if (r < R) {
                            float y0 = alpha * acc[n][m][hrow * 2] + b0, y1 = alpha * acc[n][m][hrow * 2 + 1] + b1;
                            if (Res) { const float2 rr = *reinterpret_cast<const float2*>(Res + (size_t)r * ldr + o); y0 += rr.x; y1 += rr.y; }
                            if (flags & LIN_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                            if (Mask) {
                                const float2 mm = *reinterpret_cast<const float2*>(Mask + (size_t)r * ldm + o);
                                if (!(mm.x > 0.f)) y0 = 0.f;
                                if (!(mm.y > 0.f)) y1 = 0.f;
                            }
                            *reinterpret_cast<float2*>(Y + (size_t)r * ldy + o) = make_float2(y0, y1);
                            if (out) opstage_put2(*out, r, o, y0, y1);
                        }
                    }
                }
            }
        }
    }
    LIN_PH(62);
    __syncthreads();
    LIN_PH(63);
}

// LayerNorm (as cta_ln) that can also leave its result, split into bf16 hi/lo, in an operand buffer.
static __device__ __noinline__ void cta_ln_op(float* Y, int ldy, const float* X, int ldx, const float* __restrict__ g,
                                              const float* __restrict__ b, int R, int C, float eps, const OpStage* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < R; r += NW) {
        const float* x = X + (size_t)r * ldx;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += x[c];
        const float mean = warp_sum(s) / (float)C;
        float v = 0.f;
        for (int c = lane; c < C; c += 32) { float t = x[c] - mean; v = fmaf(t, t, v); }
        const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)C + eps);
        for (int c = lane; c < C; c += 32) {
            const float y = (x[c] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
            Y[(size_t)r * ldy + c] = y;
            if (out) {
                bf16 hh, ll;
                split_bf16(y, hh, ll);
                out->hi[(size_t)r * out->stride + c] = hh;
                out->lo[(size_t)r * out->stride + c] = ll;
            }
        }
    }
    __syncthreads();
}

// Dispatch of one slot-side linear layer.  off_io: offset of the fp32 [C][O] ("in x out") copy used by the
// SIMT path; off_oi: offset of the [O][C] copy whose bf16 hi/lo images feed the tensor-core path.
template <bool MMA, int MT>
__device__ __noinline__ void lin(const float* P, const bf16* Phi, const bf16* Plo, float* Y, int ldy, const float* X,
                                    int ldx, int off_io, int off_oi, const float* bias, const float* Res, int ldr,
                                    const float* Mask, int ldm, int R, int C, int O, float alpha, int flags, float* arena,
                                    int arena_floats, const OpStage* pre = nullptr, const OpStage* out = nullptr) {
    if constexpr (MMA) {
        if ((pre || lin_mma_smem(MT, C) <= (size_t)arena_floats * 4) && (O & 7) == 0 && (C & 7) == 0) {
            cta_linear_mma<MT>(Y, ldy, X, ldx, Phi + off_oi, Plo + off_oi, bias, Res, ldr, Mask, ldm, R, C, O, alpha, flags, arena, pre, out);
            return;
        }
    }
    cta_linear(Y, ldy, X, ldx, P + off_io, O, bias, Res, ldr, Mask, ldm, R, C, O, alpha, flags, arena, arena_floats);
}
