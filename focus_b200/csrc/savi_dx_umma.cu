// d_inputs of the tcgen05 path, token-parallel (outside the sequential recurrence):
//   dxhat[n,:] = sum_i  dL_i[n,:] . qk_i  +  W_i[n,:] . dUx_i            (folded form)
//   d_inputs   = LayerNorm backward of dxhat through norm_inputs (steve.py:60), plus d gamma / d beta.
// The clip kernel stages the per-token coefficients as ready-made K-major SWIZZLE_128B operand blocks
// (one 16 KB block per iteration: [128 tokens][32 dL | 32 W]), so a 128-token tile C is fetched with one bulk copy and
// multiplied by the [64 I x 128] right-hand side R (qk_i, dUx_i rows, bf16) with 4 I tcgen05.mma (M = N = 128).
//
// Token rows move through the TMA engine only (the previous version's per-thread 16-byte row loads and stores, 32 lines
// per warp instruction, were 40 % of the kernel): the input tile arrives with two tensor copies as SWIZZLE_128B blocks
// [2][128 tokens][64 channels]; thread = token (= accumulator row = TMEM lane) normalises its row IN PLACE (bf16 z), which makes
// the buffer the MN-major operand Z of a second product; the LayerNorm-backward passes read z from that row, the d_inputs row
// overwrites it, and two tensor stores write the tile back.  Two warpgroups drain alternate tiles (two TMEM accumulators,
// one tile buffer each).
//
// The d gamma / d beta column sums over tokens never touch the accumulator rows: with dxhat = C R,
//   d gamma[c] = sum_n dxhat[n,c] z[n,c] = sum_k R[k,c] (Z^T C)[c,k]        d beta[c] = sum_n dxhat[n,c] = sum_k R[k,c] (C^T 1)[k]
// so two more tcgen05 products per tile, Z^T C and C^T 1, accumulate in TMEM over all tiles of the CTA and are contracted
// with R once at the end (the warp-shuffle column sums they replace were a third of the kernel's issue slots).
// HBM traffic = coefficients + inputs + d_inputs, each once.
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

using namespace umma;
typedef __nv_bfloat16 bf16;

namespace {
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
constexpr int DX_EPI_WARPS = 8;                 // two epilogue warpgroups (thread = token); tile j is drained by warpgroup j & 1
constexpr int DX_THREADS = (DX_EPI_WARPS + 1) * 32;   // + warp 8: coefficient loader + MMA issuer
constexpr int D128 = 128;
constexpr int XT_BYTES = 32768;                 // one token tile: [2 blocks][128 tokens][64 channels] bf16
constexpr uint32_t IDESC_A_K_B_MN_128 = idesc_bf16(128, 128, false, true);
constexpr uint32_t IDESC_A_MN_B_K_16 = idesc_bf16(128, 16, true, false);
// TMEM columns: two d xhat accumulators | Z^T C [128 channels x 64 I] | C^T 1 [64 I coefficient columns (two row tiles) x 16]
enum { TD_ACC0 = 0, TD_ACC1 = 128, TD_QT = 256, TD_CS = 448, TD_COLS = 512 };

struct alignas(64) DxUArgs {
    CUtensorMap tm_x, tm_dx;    // inputs / d_inputs as [B*T][N][128] bf16, box [1][128][64], SWIZZLE_128B
    const float2* stats;        // [B*T*N] mean, rstd
    const unsigned char* coef;  // [B*T][NTILE][I][16 KB] coefficient blocks
    const float* qk;            // saved field array: rows ((t*I+i)*B + b)*K + k, width 128
    const float* dux;           // staged field array, same row order
    const float* gamma;
    float* dgamma; float* dbeta;
    int B, T, N, K, I, NTILE, tiles_per_cta;
    const int* flags;           // [B*T] frame-staged counters of the clip kernel (nullptr: not overlapped)
    int flag_target;            // CTAs per clip
    int gate_last;              // development: every CTA waits for frame 0 of its clip (spins, but adds no traffic while the clip kernel runs)
    long long* trace;           // development (SAVI_DX_TRACE): per CTA globaltimer at start, after the flag wait, at the end
};

// shared memory plan (bytes): rhs [2 blocks][64 I rows][128 B] | coef tiles x2 [I][16 KB] | token tiles x2 (one per warpgroup) |
// ones [16][128 B] | C^T 1 vector [256] fp32 | gamma [128] | bars
__host__ __device__ inline int dx_rhs_bytes(int I) { return 2 * 64 * I * 128; }
__host__ __device__ inline int dx_smem_total(int I) { return dx_rhs_bytes(I) + 2 * I * 16384 + 2 * XT_BYTES + 2048 + 1024 + 1024; }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 :: "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n"
                 :: "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}

__global__ void __launch_bounds__(DX_THREADS, 1) dx_umma_kernel(const __grid_constant__ DxUArgs a) {
    extern __shared__ __align__(1024) unsigned char sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // (provably warp-uniform role index)
    const int I = a.I, K = a.K, N = a.N, KT = 64 * I;
    // frames in the order the backward clip kernel finishes them: t descending, all clips of a frame together
    const int t = a.T - 1 - (int)blockIdx.y / a.B, b = (int)blockIdx.y % a.B, f = b * a.T + t;
    unsigned char* rhs = sm;
    unsigned char* ct = rhs + dx_rhs_bytes(I);                 // 2 stages of I blocks
    unsigned char* xt = ct + 2 * I * 16384;                    // token tile of warpgroup g at xt + g * 32 KB
    unsigned char* ones = xt + 2 * XT_BYTES;
    float* csv = reinterpret_cast<float*>(ones + 2048);        // [256]
    float* gam = csv + 256;                                    // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(gam + 128);
    // coefficients landed[2], accumulator full[2] / drained[2], token tile landed[2], z rows written[2], column-sum products done[2], all done
    enum { BF = 0, BAF = 2, BAE = 4, BXF = 6, BZF = 8, BZE = 10, BFIN = 12 };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(a.NTILE, tile_lo + a.tiles_per_cta), nt = tile_hi - tile_lo;
    const unsigned char* coef_f = a.coef + ((size_t)f * a.NTILE + tile_lo) * I * 16384;

    if (tid == 0) {
        mbar_init(&bars[BF], 1); mbar_init(&bars[BF + 1], 1);
        mbar_init(&bars[BAF], 1); mbar_init(&bars[BAF + 1], 1);
        mbar_init(&bars[BAE], 4); mbar_init(&bars[BAE + 1], 4);
        mbar_init(&bars[BXF], 1); mbar_init(&bars[BXF + 1], 1);
        mbar_init(&bars[BZF], 4); mbar_init(&bars[BZF + 1], 4);
        mbar_init(&bars[BZE], 1); mbar_init(&bars[BZE + 1], 1); mbar_init(&bars[BFIN], 1);
        mbar_init_fence();
    }
    if (warp == DX_EPI_WARPS) tmem_alloc(tmem_slot, TD_COLS);
    for (int i = tid; i < 128; i += DX_THREADS) gam[i] = a.gamma[i];
    for (int i = tid; i < 512; i += DX_THREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;   // bf16 1.0 everywhere (any swizzle)
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");      // the weight-gradient kernel may fill the SMs this grid's last wave leaves idle
    long long t_start = 0;
    if (a.trace && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_start));
    __syncthreads();                                           // barriers exist before the first copy is issued
    // the first token tile of each warpgroup: the inputs do not depend on the clip kernel, fetch them before the flag wait
    if ((tid == 0 || tid == 128) && (tid >> 7) < nt) {
        const int g = tid >> 7;
        mbar_expect_tx(&bars[BXF + g], XT_BYTES);
        tma_load_3d(xt + g * XT_BYTES, &a.tm_x, 0, (tile_lo + g) * 128, f, &bars[BXF + g]);
        tma_load_3d(xt + g * XT_BYTES + 16384, &a.tm_x, 64, (tile_lo + g) * 128, f, &bars[BXF + g]);
    }
    if (tid == 0 && a.flags) {
        // launched as a programmatic dependent of the clip kernel: wait (acquire) until this frame's records are staged
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(a.flags + (a.gate_last ? b * a.T : f)) : "memory");
            if (v < a.flag_target) __nanosleep(256);
        } while (v < a.flag_target);
    }
    if (a.trace && tid == 0) {
        long long t1; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        a.trace[3 * cta] = t_start; a.trace[3 * cta + 1] = t1;
    }
    __syncthreads();                                           // the frame is staged before the first coefficient copy is issued
    if (tid == DX_EPI_WARPS * 32 && nt > 0) {                  // the first coefficient tile streams in while the right-hand side is staged
        // the coefficient blocks were written by the (possibly still running) clip kernel with generic-proxy stores and are
        // read here through the async proxy (bulk copy): order the two proxies after the acquire above
        asm volatile("fence.proxy.async.global;\n" ::: "memory");
        mbar_expect_tx(&bars[BF], I * 16384); bulk_g2s(ct, coef_f, I * 16384, &bars[BF]);
    }
    // right-hand side rows: iteration i -> rows [64 i, 64 i + 64): qk_i (32 rows, zero beyond K) then dUx_i; MN-major, two 64-column blocks
    // (all of a thread's loads are issued before the first is converted: one DRAM round trip instead of three)
    {
        constexpr int RH_IT = (64 * 3 * 16 + DX_THREADS - 1) / DX_THREADS;      // I <= 3
        float4 p[RH_IT], q[RH_IT];
#pragma unroll
        for (int it = 0; it < RH_IT; ++it) {
            const int idx = tid + it * DX_THREADS;
            const int k = idx >> 4, c8 = (idx & 15) * 8;        // row, first of 8 columns
            const int i = k >> 6, which = (k >> 5) & 1, s = k & 31;
            p[it] = q[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < KT * 16 && s < K) {
                const float* src = (which == 0 ? a.qk : a.dux) + ((((size_t)t * I + i) * a.B + b) * K + s) * D128 + c8;
                p[it] = ld4(src); q[it] = ld4(src + 4);
            }
        }
#pragma unroll
        for (int it = 0; it < RH_IT; ++it) {
            const int idx = tid + it * DX_THREADS;
            const int k = idx >> 4, c8 = (idx & 15) * 8;
            if (idx < KT * 16)
                *reinterpret_cast<uint4*>(rhs + (c8 >> 6) * (KT * 128) + k * 128 + ((((c8 & 63) >> 3) ^ (k & 7)) << 4)) =
                    make_uint4(pack2(p[it].x, p[it].y), pack2(p[it].z, p[it].w), pack2(q[it].x, q[it].y), pack2(q[it].z, q[it].w));
        }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == DX_EPI_WARPS) {
        // ---- coefficient loader + MMA issuer (warp-uniform, one lane issues) ----
        const bool el = (lane == 0);                           // lane 0 issued the first copy: keep one issuing thread
        const uint32_t rhs0 = dlo_mn(smem_u32(rhs), KT * 128);
        const uint32_t idesc_qt = idesc_bf16(128, KT, true, true);
        const uint32_t ones0 = dlo_k(smem_u32(ones));
        for (int j = 0; j < nt; ++j) {
            const int st = j & 1;
            // prefetch the next tile's coefficient blocks (its stage was released when the products of tile j-1 completed)
            if (j + 1 < nt) {
                if (j >= 1) mbar_wait(&bars[BZE + (st ^ 1)], ((j - 1) >> 1) & 1u);
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
            // column-sum products of the same tile: the coefficient blocks once more, as MN-major operands (rows = tokens = the
            // contraction index), against the normalised tile Z of the warpgroup that drains this tile
            mbar_wait(&bars[BZF + st], (j >> 1) & 1u);
            fence_after_sync();
            if (el) {
                const uint32_t cmn = dlo_mn(smem_u32(ct) + st * I * 16384, 16384), zmn = dlo_mn(smem_u32(xt) + st * XT_BYTES, 16384);
                const uint32_t first = j > 0 ? 1u : 0u;
                for (int kk = 0; kk < 8; ++kk)                            // Z^T C: [128 channels] x [64 I coefficient columns], 16 tokens per step
                    mma_lo(tb + TD_QT, zmn + kk * 128, cmn + kk * 128, idesc_qt, kk > 0 ? 1u : first);
                // C^T 1: coefficient columns in row tiles of 128 (the upper half of a partly filled tile reads whatever follows the
                // stage: those rows land in accumulator lanes nobody reads)
                for (int p = 0; p * 128 < KT; ++p)
                    for (int kk = 0; kk < 8; ++kk)
                        mma_lo(tb + TD_CS + p * 16, cmn + p * 2048 + kk * 128, ones0 + (kk & 3) * 2, IDESC_A_MN_B_K_16, kk > 0 ? 1u : first);
                mma_commit(&bars[BZE + st]);
                if (j == nt - 1) mma_commit(&bars[BFIN]);
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue: thread = token; warpgroup g drains the tiles with j & 1 == g ----
        const int g = warp >> 2;
        const uint32_t tl = (uint32_t)((warp & 3) * 32) << 16;
        const int r = tid & 127;                                          // row inside the tile
        unsigned char* xrow = xt + g * XT_BYTES + r * 128;                // chunk q (8 channels) of the row: block q >> 3, 16-byte slot (q & 7) ^ (r & 7)
        const uint32_t sw = (uint32_t)(r & 7);
        for (int j = g; j < nt; j += 2) {
            const uint32_t par = (uint32_t)(j >> 1) & 1u;
            const int n = (tile_lo + j) * 128 + r;
            const bool valid = n < N;
            const float2 ms = valid ? a.stats[(size_t)f * N + n] : make_float2(0.f, 0.f);
            // ---- normalise the row in place: z = (x - mean) rstd, bf16 (rows past N arrive as zeros and stay zero) ----
            mbar_wait(&bars[BXF + g], par);
            {
                const float nm = -ms.x * ms.y;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    uint4* slot = reinterpret_cast<uint4*>(xrow + (q >> 3) * 16384 + ((((uint32_t)q & 7u) ^ sw) << 4));
                    const uint4 raw = *slot;
                    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&raw);
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xf = __bfloat1622float2(xp[e]);
                        o[e] = pack2(fmaf(xf.x, ms.y, nm), fmaf(xf.y, ms.y, nm));
                    }
                    *slot = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BZF + g]);
            mbar_wait(&bars[BAF + g], par);
            fence_after_sync();
            const uint32_t acc = tb + tl + (g ? TD_ACC1 : TD_ACC0);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                 // pass 1: the two LayerNorm-backward row sums
                float v[32];
                tmem_ld32(acc + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 zq = *reinterpret_cast<const uint4*>(xrow + (c >> 1) * 16384 + ((((uint32_t)(c & 1) * 4u + q) ^ sw) << 4));
                    const __nv_bfloat162* zp = reinterpret_cast<const __nv_bfloat162*>(&zq);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 z = __bfloat1622float2(zp[e]);
                        const int col = c * 32 + q * 8 + e * 2, u = q * 8 + e * 2;
                        const float dz0 = v[u] * gam[col], dz1 = v[u + 1] * gam[col + 1];
                        s1 += dz0 + dz1; s2 = fmaf(dz0, z.x, fmaf(dz1, z.y, s2));
                    }
                }
            }
            s1 *= (1.0f / D128); s2 *= (1.0f / D128);
            mbar_wait(&bars[BZE + g], par);                               // the column-sum products have read z: the row may be overwritten
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                 // pass 2: d_inputs row, in place
                float v[32];
                tmem_ld32(acc + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4* slot = reinterpret_cast<uint4*>(xrow + (c >> 1) * 16384 + ((((uint32_t)(c & 1) * 4u + q) ^ sw) << 4));
                    const uint4 zq = *slot;
                    const __nv_bfloat162* zp = reinterpret_cast<const __nv_bfloat162*>(&zq);
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 z = __bfloat1622float2(zp[e]);
                        const int col = c * 32 + q * 8 + e * 2;
                        const float dz0 = v[q * 8 + e * 2] * gam[col], dz1 = v[q * 8 + e * 2 + 1] * gam[col + 1];
                        o[e] = pack2(ms.y * (dz0 - s1 - z.x * s2), ms.y * (dz1 - s1 - z.y * s2));
                    }
                    *slot = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAE + g]);
            // ---- the tile back to HBM (rows past N are clipped by the tensor map), then this warpgroup's next tile in ----
            fence_async_smem();
            asm volatile("bar.sync %0, 128;\n" :: "r"(1 + g) : "memory");
            if (r == 0) {
                tma_store_3d(&a.tm_dx, 0, (tile_lo + j) * 128, f, xt + g * XT_BYTES);
                tma_store_3d(&a.tm_dx, 64, (tile_lo + j) * 128, f, xt + g * XT_BYTES + 16384);
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                if (j + 2 < nt) {
                    asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");       // the stores have read the buffer
                    mbar_expect_tx(&bars[BXF + g], XT_BYTES);
                    tma_load_3d(xt + g * XT_BYTES, &a.tm_x, 0, (tile_lo + j + 2) * 128, f, &bars[BXF + g]);
                    tma_load_3d(xt + g * XT_BYTES + 16384, &a.tm_x, 64, (tile_lo + j + 2) * 128, f, &bars[BXF + g]);
                }
            }
        }
        if (r == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");             // d_inputs rows are in global memory
        // ---- d gamma / d beta: contract the two TMEM accumulators with the right-hand side, one atomic per column and CTA.
        // Warpgroup 0: thread = channel c = accumulator lane of Z^T C; warpgroup 1: lane = coefficient column of C^T 1.
        auto rhs_at = [&](int k, int c) {
            return __bfloat162float(*reinterpret_cast<const bf16*>(rhs + (c >> 6) * (KT * 128) + k * 128 + (((((c & 63) >> 3) ^ (k & 7))) << 4) + (c & 7) * 2));
        };
        if (nt > 0) {
            mbar_wait(&bars[BFIN], 0);
            fence_after_sync();
            if (g == 0) {
                float dg = 0.f;
                for (int k0 = 0; k0 < KT; k0 += 32) {
                    float v[32];
                    tmem_ld32(tb + tl + TD_QT + k0, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int e = 0; e < 32; ++e) dg = fmaf(v[e], rhs_at(k0 + e, r), dg);
                }
                atomicAdd(a.dgamma + r, dg);
            } else {
                float v[32];
                tmem_ld32(tb + tl + TD_CS, v);
                tmem_wait_ld();
                csv[r] = v[0]; csv[128 + r] = v[16];
                asm volatile("bar.sync 2, 128;\n" ::: "memory");
                float db = 0.f;
                for (int k = 0; k < KT; ++k) db = fmaf(csv[k], rhs_at(k, r), db);
                atomicAdd(a.dbeta + r, db);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (a.trace && tid == 0) {
        long long t2; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t2));
        a.trace[3 * (blockIdx.y * gridDim.x + blockIdx.x) + 2] = t2;
    }
    if (warp == DX_EPI_WARPS) tmem_dealloc(tb, TD_COLS);
}

// [frames][N][128] bf16 as a rank-3 tensor map with box [1][128 tokens][64 channels], SWIZZLE_128B: the shared-memory image of a
// box is one K-major / MN-major UMMA operand block; rows past N are zero-filled on loads and clipped on stores
cudaError_t make_token_map(CUtensorMap* tm, const void* base, int frames, int N) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const cuuint64_t dims[3] = {128, (cuuint64_t)N, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {256, (cuuint64_t)N * 256};
    const cuuint32_t box[3] = {64, 128, 1}, estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
}  // namespace

cudaError_t savi_launch_dx_umma(const BwdArgs& a, const void* inputs, void* grad_inputs, bool overlap, cudaStream_t st) {
    const Dims& d = a.d;
    DxUArgs x;
    cudaError_t e = make_token_map(&x.tm_x, inputs, d.B * d.T, d.N);
    if (e != cudaSuccess) return e;
    e = make_token_map(&x.tm_dx, grad_inputs, d.B * d.T, d.N);
    if (e != cudaSuccess) return e;
    x.stats = reinterpret_cast<const float2*>(a.saved + a.sl.stats);
    x.coef = reinterpret_cast<const unsigned char*>(a.ws) + a.wl.coef;
    x.qk = reinterpret_cast<const float*>(a.saved + a.sl.fbase) + a.sl.qk;
    x.dux = a.ws + a.wl.duxs;
    x.gamma = a.packed + a.po.ln_in_w;
    x.dgamma = a.grad_params + a.po.ln_in_w; x.dbeta = a.grad_params + a.po.ln_in_b;
    x.B = d.B; x.T = d.T; x.N = d.N; x.K = d.K; x.I = d.I; x.NTILE = d.NTILE;
    // whole frames per CTA (one staging of the right-hand side) once there are >= 2 waves of frames; else >= 2 CTAs per frame
    x.tiles_per_cta = d.NTILE <= 8 ? ((d.B * d.T >= 2 * 148 || d.NTILE < 8) ? d.NTILE : 4) : 8;
    // (finer work items would shorten the tail left when the overlapped clip kernel ends, but measured slower: 1.724 vs 1.701 ms/step)
    if (savi_options().dx_tpc > 0) x.tiles_per_cta = savi_options().dx_tpc;                // development knob
    const int smem = dx_smem_total(d.I);
    e = cudaFuncSetAttribute(dx_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    x.flags = overlap ? reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(a.ws) + a.wl.flags) : nullptr;
    x.flag_target = d.CN;
    x.gate_last = savi_options().dx_gate_last;
    x.trace = (a.dbg && savi_options().dx_trace) ? a.dbg + 64 : nullptr;    // the debug buffer then holds 64 + 3 * grid + 2 entries
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((d.NTILE + x.tiles_per_cta - 1) / x.tiles_per_cta, d.B * d.T);
    cfg.blockDim = dim3(DX_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // may start while the preceding clip kernel still runs
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, dx_umma_kernel, x);
}
// (I <= 3 also bounds the TMEM plan: 256 + 64 I + 32 columns)
int savi_dx_umma_smem_bytes(int I) { return I <= 3 ? dx_smem_total(I) : (1 << 30); }
