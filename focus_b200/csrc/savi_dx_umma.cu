// d_inputs of the tcgen05 path, token-parallel (outside the sequential recurrence):
//   dxhat[n,:] = sum_i  dL_i[n,:] . qk_i  +  W_i[n,:] . dUx_i            (folded form)
//   d_inputs   = LayerNorm backward of dxhat through norm_inputs (steve.py:60), plus d gamma / d beta.
// The clip kernel stages the per-token coefficients as ready-made K-major SWIZZLE_128B operand blocks
// (one 16 KB block per iteration: [128 tokens][32 dL | 32 W]), so a 128-token tile is fetched with one bulk copy and
// multiplied by the [64 I x 128] right-hand side (qk_i, dUx_i rows, bf16) with 4 I tcgen05.mma (M = N = 128).
// Epilogue: thread = token (accumulator row = TMEM lane), so the LayerNorm-backward row sums are thread-local;
// the d gamma / d beta column sums over tokens are two more products with a ones matrix, accumulated in TMEM over
// all tiles of the CTA.  HBM traffic = coefficients + inputs + d_inputs, each once.
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"

using namespace umma;
typedef __nv_bfloat16 bf16;

namespace {
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
constexpr int DX_THREADS = 160;                 // warps 0-3: epilogue (thread = token); warp 4: loader + MMA issuer
constexpr int D128 = 128;
constexpr uint32_t IDESC_A_K_B_MN_128 = idesc_bf16(128, 128, false, true);
constexpr uint32_t IDESC_A_MN_B_MN_128 = idesc_bf16(128, 128, true, true);
enum { TD_ACC0 = 0, TD_ACC1 = 128, TD_SUM1 = 256, TD_SUM2 = 384, TD_COLS = 512 };

struct DxUArgs {
    const bf16* x;              // inputs [B*T][N][128]
    const float2* stats;        // [B*T*N] mean, rstd
    const unsigned char* coef;  // [B*T][NTILE][I][16 KB] coefficient blocks
    const float* qk;            // saved field array: rows ((t*I+i)*B + b)*K + k, width 128
    const float* dux;           // staged field array, same row order
    const float* gamma;
    bf16* dx;
    float* dgamma; float* dbeta;
    int B, T, N, K, I, NTILE, tiles_per_cta;
};

// shared memory plan (bytes): rhs [2 blocks][64 I rows][128 B] | coef tiles x2 [I][16 KB] | dxh, dxhz operands [2 blocks][128 rows][128 B] each | ones 4 KB | gamma, bars
__host__ __device__ inline int dx_rhs_bytes(int I) { return 2 * 64 * I * 128; }
__host__ __device__ inline int dx_smem_total(int I) { return dx_rhs_bytes(I) + 2 * I * 16384 + 2 * 32768 + 4096 + 1024; }

__global__ void __launch_bounds__(DX_THREADS, 1) dx_umma_kernel(const __grid_constant__ DxUArgs a) {
    extern __shared__ __align__(1024) unsigned char sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.I, K = a.K, N = a.N, KT = 64 * I;
    const int f = blockIdx.y, b = f / a.T, t = f - b * a.T;
    unsigned char* rhs = sm;
    unsigned char* ct = rhs + dx_rhs_bytes(I);                 // 2 stages of I blocks
    unsigned char* s1op = ct + 2 * I * 16384;
    unsigned char* s2op = s1op + 32768;
    unsigned char* ones = s2op + 32768;
    float* gam = reinterpret_cast<float*>(ones + 4096);        // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 4096 + 512);   // full[2], accfull[2], accfree[2], sumrdy, sumfree
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    enum { BF = 0, BAF = 2, BAE = 4, BSR = 6, BSF = 7 };
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(a.NTILE, tile_lo + a.tiles_per_cta), nt = tile_hi - tile_lo;

    if (tid == 0) {
        mbar_init(&bars[BF], 1); mbar_init(&bars[BF + 1], 1);
        mbar_init(&bars[BAF], 1); mbar_init(&bars[BAF + 1], 1);
        mbar_init(&bars[BAE], 4); mbar_init(&bars[BAE + 1], 4);
        mbar_init(&bars[BSR], 4); mbar_init(&bars[BSF], 1);
        mbar_init_fence();
    }
    if (warp == 4) tmem_alloc(tmem_slot, TD_COLS);
    for (int i = tid; i < 2048; i += DX_THREADS) reinterpret_cast<uint16_t*>(ones)[i] = 0x3F80;
    for (int i = tid; i < 128; i += DX_THREADS) gam[i] = a.gamma[i];
    // right-hand side rows: iteration i -> rows [64 i, 64 i + 64): qk_i (32 rows, zero beyond K) then dUx_i; MN-major, two 64-column blocks
    for (int idx = tid; idx < KT * 16; idx += DX_THREADS) {
        const int k = idx >> 4, c8 = (idx & 15) * 8;            // row, first of 8 columns
        const int i = k >> 6, which = (k >> 5) & 1, s = k & 31;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (s < K) {
            const float* src = (which == 0 ? a.qk : a.dux) + ((((size_t)t * I + i) * a.B + b) * K + s) * D128 + c8;
            const float4 p = ld4(src), q = ld4(src + 4);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(p.x, p.y), h1 = __floats2bfloat162_rn(p.z, p.w);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(q.x, q.y), h3 = __floats2bfloat162_rn(q.z, q.w);
            v.x = *reinterpret_cast<unsigned*>(&h0); v.y = *reinterpret_cast<unsigned*>(&h1);
            v.z = *reinterpret_cast<unsigned*>(&h2); v.w = *reinterpret_cast<unsigned*>(&h3);
        }
        *reinterpret_cast<uint4*>(rhs + (c8 >> 6) * (KT * 128) + k * 128 + ((((c8 & 63) >> 3) ^ (k & 7)) << 4)) = v;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;
    const unsigned char* coef_f = a.coef + ((size_t)f * a.NTILE + tile_lo) * I * 16384;

    if (warp == 4) {
        // ---- loader + issuer (warp-uniform, one elected lane issues) ----
        const bool el = elect_one();
        if (el && nt > 0) { mbar_expect_tx(&bars[BF], I * 16384); bulk_g2s(ct, coef_f, I * 16384, &bars[BF]); }
        __syncwarp();
        const uint32_t rhs0 = dlo_mn(smem_u32(rhs), KT * 128);
        const uint32_t one0 = dlo_mn(smem_u32(ones), 2048);
        const uint32_t s1d = dlo_mn(smem_u32(s1op), 16384), s2d = dlo_mn(smem_u32(s2op), 16384);
        for (int j = 0; j < nt; ++j) {
            const int st = j & 1;
            // prefetch the next tile's coefficient blocks (its stage was released when the MMAs of tile j-1 completed)
            if (j + 1 < nt) {
                if (j >= 1) mbar_wait(&bars[BAF + (st ^ 1)], ((j - 1) >> 1) & 1u);
                if (el) { mbar_expect_tx(&bars[BF + (st ^ 1)], I * 16384); bulk_g2s(ct + (st ^ 1) * I * 16384, coef_f + (size_t)(j + 1) * I * 16384, I * 16384, &bars[BF + (st ^ 1)]); }
                __syncwarp();
            }
            mbar_wait(&bars[BF + st], (j >> 1) & 1u);                     // coefficients landed
            mbar_wait(&bars[BAE + st], ((j >> 1) & 1u) ^ 1u);             // accumulator drained by the epilogue of tile j-2
            fence_after_sync();
            const uint32_t a0 = dlo_k(smem_u32(ct) + st * I * 16384);
            if (el) {
                for (int ks = 0; ks < 4 * I; ++ks)
                    mma_lo(tb + (st ? TD_ACC1 : TD_ACC0), a0 + (ks >> 2) * 1024 + (ks & 3) * 2, rhs0 + ks * 128, IDESC_A_K_B_MN_128, ks > 0 ? 1u : 0u);
                mma_commit(&bars[BAF + st]);
            }
            __syncwarp();
            // column sums of the previous tile (its operands were written by the epilogue)
            if (j >= 1) {
                mbar_wait(&bars[BSR], (j - 1) & 1u);
                fence_after_sync();
                if (el) {
                    for (int kt = 0; kt < 8; ++kt) {
                        mma_lo(tb + TD_SUM1, one0, s1d + kt * 128, IDESC_A_MN_B_MN_128, (j > 1 || kt > 0) ? 1u : 0u);
                        mma_lo(tb + TD_SUM2, one0, s2d + kt * 128, IDESC_A_MN_B_MN_128, (j > 1 || kt > 0) ? 1u : 0u);
                    }
                    mma_commit(&bars[BSF]);
                }
                __syncwarp();
            }
        }
        if (nt > 0) {
            mbar_wait(&bars[BSR], (nt - 1) & 1u);
            fence_after_sync();
            if (el) {
                for (int kt = 0; kt < 8; ++kt) {
                    mma_lo(tb + TD_SUM1, one0, s1d + kt * 128, IDESC_A_MN_B_MN_128, (nt > 1 || kt > 0) ? 1u : 0u);
                    mma_lo(tb + TD_SUM2, one0, s2d + kt * 128, IDESC_A_MN_B_MN_128, (nt > 1 || kt > 0) ? 1u : 0u);
                }
                mma_commit(&bars[BSF]);
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: thread = token ----
        const uint32_t tl = (uint32_t)(warp * 32) << 16;
        const int r = tid;                                                // row inside the tile
        for (int j = 0; j < nt; ++j) {
            const int st = j & 1;
            const int n = (tile_lo + j) * 128 + r;
            const bool valid = n < N;
            const float2 ms = valid ? a.stats[(size_t)f * N + n] : make_float2(0.f, 0.f);
            const bf16* xrow = a.x + ((size_t)f * N + (valid ? n : 0)) * D128;
            bf16* orow = a.dx + ((size_t)f * N + (valid ? n : 0)) * D128;
            mbar_wait(&bars[BAF + st], (j >> 1) & 1u);
            fence_after_sync();
            if (j >= 1) mbar_wait(&bars[BSF], (j - 1) & 1u);              // the previous tile's column-sum products have read their operands
            const uint32_t acc = tb + tl + (st ? TD_ACC1 : TD_ACC0);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {                                 // pass 1: row sums + column-sum operands
                float v[32];
                tmem_ld32(acc + c * 32, v);
                uint4 xq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) xq[q] = valid ? *reinterpret_cast<const uint4*>(xrow + c * 32 + q * 8) : make_uint4(0u, 0u, 0u, 0u);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xq[q]);
                    uint32_t o1[4], o2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xf = __bfloat1622float2(xp[e]);
                        const int col = c * 32 + q * 8 + e * 2;
                        const float z0 = (xf.x - ms.x) * ms.y, z1 = (xf.y - ms.x) * ms.y;
                        const float d0 = valid ? v[q * 8 + e * 2] : 0.f, d1 = valid ? v[q * 8 + e * 2 + 1] : 0.f;
                        const float dz0 = d0 * gam[col], dz1 = d1 * gam[col + 1];
                        s1 += dz0 + dz1; s2 = fmaf(dz0, z0, fmaf(dz1, z1, s2));
                        o1[e] = pack2(d0, d1); o2[e] = pack2(d0 * z0, d1 * z1);
                    }
                    const int col8 = c * 32 + q * 8;                      // MN-major B operand [128 token rows][128 columns], two 64-column blocks
                    const uint32_t off = (uint32_t)(col8 >> 6) * 16384u + (uint32_t)r * 128u + (((((uint32_t)col8 & 63u) >> 3) ^ ((uint32_t)r & 7u)) << 4);
                    *reinterpret_cast<uint4*>(s1op + off) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
                    *reinterpret_cast<uint4*>(s2op + off) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BSR]);
            s1 *= (1.0f / D128); s2 *= (1.0f / D128);
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {                                 // pass 2: d_inputs row
                float v[32];
                tmem_ld32(acc + c * 32, v);
                uint4 xq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) xq[q] = valid ? *reinterpret_cast<const uint4*>(xrow + c * 32 + q * 8) : make_uint4(0u, 0u, 0u, 0u);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xq[q]);
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xf = __bfloat1622float2(xp[e]);
                        const int col = c * 32 + q * 8 + e * 2;
                        const float z0 = (xf.x - ms.x) * ms.y, z1 = (xf.y - ms.x) * ms.y;
                        const float dz0 = v[q * 8 + e * 2] * gam[col], dz1 = v[q * 8 + e * 2 + 1] * gam[col + 1];
                        o[e] = pack2(ms.y * (dz0 - s1 - z0 * s2), ms.y * (dz1 - s1 - z1 * s2));
                    }
                    if (valid) *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAE + st]);
        }
        // d gamma / d beta: every row of the ones-products holds the column sums; lanes 0..127 each add one column
        if (nt > 0) {
            mbar_wait(&bars[BSF], (nt - 1) & 1u);
            fence_after_sync();
            // thread r reads row r (any row would do) but only its own column r: load the 32-column chunk that contains it
            float v[32];
            tmem_ld32(tb + tl + TD_SUM2 + (r & ~31), v);
            tmem_wait_ld();
            float dgv = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) if (e == (r & 31)) dgv = v[e];
            tmem_ld32(tb + tl + TD_SUM1 + (r & ~31), v);
            tmem_wait_ld();
            float dbv = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) if (e == (r & 31)) dbv = v[e];
            atomicAdd(a.dgamma + r, dgv);
            atomicAdd(a.dbeta + r, dbv);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tb, TD_COLS);
}
}  // namespace

cudaError_t savi_launch_dx_umma(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st) {
    const Dims& d = a.d;
    DxUArgs x;
    x.x = reinterpret_cast<const bf16*>(inputs);
    x.stats = reinterpret_cast<const float2*>(a.saved + a.sl.stats);
    x.coef = reinterpret_cast<const unsigned char*>(a.ws) + a.wl.coef;
    x.qk = reinterpret_cast<const float*>(a.saved + a.sl.fbase) + a.sl.qk;
    x.dux = a.ws + a.wl.duxs;
    x.gamma = a.packed + a.po.ln_in_w;
    x.dx = reinterpret_cast<bf16*>(grad_inputs);
    x.dgamma = a.grad_params + a.po.ln_in_w; x.dbeta = a.grad_params + a.po.ln_in_b;
    x.B = d.B; x.T = d.T; x.N = d.N; x.K = d.K; x.I = d.I; x.NTILE = d.NTILE;
    x.tiles_per_cta = d.NTILE >= 8 ? 4 : d.NTILE;                           // >= 2 CTAs per frame once frames have 8 tiles
    const int smem = dx_smem_total(d.I);
    cudaError_t e = cudaFuncSetAttribute(dx_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dim3 grid((d.NTILE + x.tiles_per_cta - 1) / x.tiles_per_cta, d.B * d.T);
    dx_umma_kernel<<<grid, DX_THREADS, smem, st>>>(x);
    return cudaGetLastError();
}
int savi_dx_umma_smem_bytes(int I) { return dx_smem_total(I); }
