// d_inputs of the tcgen05 path, token-parallel (outside the sequential recurrence):
//   dxhat[n,:] = sum_i  dL_i[n,:] . qk_i  +  W_i[n,:] . dUx_i            (folded form)
//   d_inputs   = LayerNorm backward of dxhat through norm_inputs (steve.py:60), plus d gamma / d beta.
// The clip kernel stages the per-token coefficients as ready-made K-major SWIZZLE_128B operand blocks
// (one 16 KB block per iteration: [128 tokens][32 dL | 32 W]), so a 128-token tile is fetched with one bulk copy and
// multiplied by the [64 I x 128] right-hand side (qk_i, dUx_i rows, bf16) with 4 I tcgen05.mma (M = N = 128).
// Epilogue: thread = token (accumulator row = TMEM lane), so the LayerNorm-backward row sums are thread-local; two
// warpgroups drain alternate tiles (two TMEM accumulators), each thread fetching its whole input row (16 x 16 B) before
// it waits for the accumulator; the d gamma / d beta column sums over tokens are fp32 recursive-halving warp shuffles
// accumulated in registers over all tiles of the CTA.  HBM traffic = coefficients + inputs + d_inputs, each once.
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"
#include <cstdlib>

using namespace umma;
typedef __nv_bfloat16 bf16;

namespace {
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 16-byte global load the compiler may not merge with an earlier load of the same address (the row is deliberately fetched twice)
__device__ __forceinline__ uint4 ldg_v4_again(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
constexpr int DX_EPI_WARPS = 8;                 // two epilogue warpgroups (thread = token); tile j is drained by warpgroup j & 1
constexpr int DX_THREADS = (DX_EPI_WARPS + 1) * 32;   // + warp 8: loader + MMA issuer
constexpr int D128 = 128;
constexpr uint32_t IDESC_A_K_B_MN_128 = idesc_bf16(128, 128, false, true);
enum { TD_ACC0 = 0, TD_ACC1 = 128, TD_COLS = 256 };

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
    const int* flags;           // [B*T] frame-staged counters of the clip kernel (nullptr: not overlapped)
    int flag_target;            // CTAs per clip
    int gate_last;              // development: every CTA waits for frame 0 of its clip (spins, but adds no traffic while the clip kernel runs)
    long long* trace;           // development (SAVI_DX_TRACE): per CTA globaltimer at start, after the flag wait, at the end
};

// shared memory plan (bytes): rhs [2 blocks][64 I rows][128 B] | coef tiles x2 [I][16 KB] | column-sum scratch [8 warps][256] fp32 | gamma, bars
__host__ __device__ inline int dx_rhs_bytes(int I) { return 2 * 64 * I * 128; }
__host__ __device__ inline int dx_smem_total(int I) { return dx_rhs_bytes(I) + 2 * I * 16384 + DX_EPI_WARPS * 256 * 4 + 1024; }

// v[j] of lane l = element (row l, column j) of a 32 x 32 block; returns to lane l the sum of column l over the 32 rows
// (recursive halving: 31 shuffles instead of 32 x 5)
template <int H>
__device__ __forceinline__ void colsum_step(float (&v)[32], int lane) {
    const bool up = (lane & H) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float send = up ? v[i] : v[i + H];
        const float keep = up ? v[i + H] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
    }
}
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
    colsum_step<16>(v, lane); colsum_step<8>(v, lane); colsum_step<4>(v, lane); colsum_step<2>(v, lane); colsum_step<1>(v, lane);
    return v[0];
}

__global__ void __launch_bounds__(DX_THREADS, 1) dx_umma_kernel(const __grid_constant__ DxUArgs a) {
    extern __shared__ __align__(1024) unsigned char sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.I, K = a.K, N = a.N, KT = 64 * I;
    // frames in the order the backward clip kernel finishes them: t descending, all clips of a frame together
    const int t = a.T - 1 - (int)blockIdx.y / a.B, b = (int)blockIdx.y % a.B, f = b * a.T + t;
    unsigned char* rhs = sm;
    unsigned char* ct = rhs + dx_rhs_bytes(I);                 // 2 stages of I blocks
    float* csum = reinterpret_cast<float*>(ct + 2 * I * 16384);   // [8 warps][2][128]
    float* gam = csum + DX_EPI_WARPS * 256;                    // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(gam + 128);   // full[2], accfull[2], accfree[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    enum { BF = 0, BAF = 2, BAE = 4 };
    const int tile_lo = blockIdx.x * a.tiles_per_cta, tile_hi = min(a.NTILE, tile_lo + a.tiles_per_cta), nt = tile_hi - tile_lo;
    const unsigned char* coef_f = a.coef + ((size_t)f * a.NTILE + tile_lo) * I * 16384;

    if (tid == 0) {
        mbar_init(&bars[BF], 1); mbar_init(&bars[BF + 1], 1);
        mbar_init(&bars[BAF], 1); mbar_init(&bars[BAF + 1], 1);
        mbar_init(&bars[BAE], 4); mbar_init(&bars[BAE + 1], 4);
        mbar_init_fence();
    }
    if (warp == DX_EPI_WARPS) tmem_alloc(tmem_slot, TD_COLS);
    for (int i = tid; i < 128; i += DX_THREADS) gam[i] = a.gamma[i];
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");      // the weight-gradient kernel may fill the SMs this grid's last wave leaves idle
    long long t_start = 0;
    if (a.trace && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_start));
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
    __syncthreads();                                           // barriers exist (and the frame is staged) before the first bulk copy is issued
    if (tid == DX_EPI_WARPS * 32 && nt > 0) {                  // the first coefficient tile streams in while the right-hand side is staged
        // the coefficient blocks were written by the (possibly still running) clip kernel with generic-proxy stores and are
        // read here through the async proxy (bulk copy): order the two proxies after the acquire above
        asm volatile("fence.proxy.async.global;\n" ::: "memory");
        mbar_expect_tx(&bars[BF], I * 16384); bulk_g2s(ct, coef_f, I * 16384, &bars[BF]);
    }
    // right-hand side rows: iteration i -> rows [64 i, 64 i + 64): qk_i (32 rows, zero beyond K) then dUx_i; MN-major, two 64-column blocks
#pragma unroll 4
    for (int idx = tid; idx < KT * 16; idx += DX_THREADS) {
        const int k = idx >> 4, c8 = (idx & 15) * 8;            // row, first of 8 columns
        const int i = k >> 6, which = (k >> 5) & 1, s = k & 31;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (s < K) {
            const float* src = (which == 0 ? a.qk : a.dux) + ((((size_t)t * I + i) * a.B + b) * K + s) * D128 + c8;
            const float4 p = ld4(src), q = ld4(src + 4);
            v.x = pack2(p.x, p.y); v.y = pack2(p.z, p.w); v.z = pack2(q.x, q.y); v.w = pack2(q.z, q.w);
        }
        *reinterpret_cast<uint4*>(rhs + (c8 >> 6) * (KT * 128) + k * 128 + ((((c8 & 63) >> 3) ^ (k & 7)) << 4)) = v;
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp == DX_EPI_WARPS) {
        // ---- loader + issuer (warp-uniform, one elected lane issues) ----
        const bool el = (lane == 0);                           // lane 0 issued the first copy: keep one issuing thread
        const uint32_t rhs0 = dlo_mn(smem_u32(rhs), KT * 128);
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
        }
    } else {
        // ---- epilogue: thread = token; warpgroup g drains the tiles with j & 1 == g ----
        const int g = warp >> 2;
        const uint32_t tl = (uint32_t)((warp & 3) * 32) << 16;
        const int r = tid & 127;                                          // row inside the tile
        float cs1[4] = {0.f, 0.f, 0.f, 0.f}, cs2[4] = {0.f, 0.f, 0.f, 0.f};   // column sums (d beta, d gamma) of columns 32 c + lane over this warp's rows
        for (int j = g; j < nt; j += 2) {
            const int n = (tile_lo + j) * 128 + r;
            const bool valid = n < N;
            const float2 ms = valid ? a.stats[(size_t)f * N + n] : make_float2(0.f, 0.f);
            const bf16* xrow = a.x + ((size_t)f * N + (valid ? n : 0)) * D128;
            bf16* orow = a.dx + ((size_t)f * N + (valid ? n : 0)) * D128;
            uint4 xq[16];                                                 // the whole input row, in flight before the accumulator is awaited
#pragma unroll
            for (int q = 0; q < 16; ++q) xq[q] = valid ? ldg_v4_again(xrow + q * 8) : make_uint4(0u, 0u, 0u, 0u);
            mbar_wait(&bars[BAF + g], (j >> 1) & 1u);
            fence_after_sync();
            const uint32_t acc = tb + tl + (g ? TD_ACC1 : TD_ACC0);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                 // pass 1: row sums + column sums
                float v[32], w[32];
                tmem_ld32(acc + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xq[c * 4 + q]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 xf = __bfloat1622float2(xp[e]);
                        const int col = c * 32 + q * 8 + e * 2, u = q * 8 + e * 2;
                        const float z0 = (xf.x - ms.x) * ms.y, z1 = (xf.y - ms.x) * ms.y;
                        const float d0 = valid ? v[u] : 0.f, d1 = valid ? v[u + 1] : 0.f;
                        const float dz0 = d0 * gam[col], dz1 = d1 * gam[col + 1];
                        s1 += dz0 + dz1; s2 = fmaf(dz0, z0, fmaf(dz1, z1, s2));
                        v[u] = d0; v[u + 1] = d1; w[u] = d0 * z0; w[u + 1] = d1 * z1;
                    }
                }
                cs2[c] += warp_colsum32(w, lane);
                cs1[c] += warp_colsum32(v, lane);
            }
            s1 *= (1.0f / D128); s2 *= (1.0f / D128);
            // the row again (L1 / L2 hit): keeping it in registers across the column sums would spill
#pragma unroll
            for (int q = 0; q < 16; ++q) xq[q] = valid ? ldg_v4_again(xrow + q * 8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                 // pass 2: d_inputs row
                float v[32];
                tmem_ld32(acc + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xq[c * 4 + q]);
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
            if (lane == 0) mbar_arrive(&bars[BAE + g]);
        }
        // d gamma / d beta: combine the 8 warps' column sums, one atomic per column and CTA
#pragma unroll
        for (int c = 0; c < 4; ++c) { csum[warp * 256 + c * 32 + lane] = cs1[c]; csum[warp * 256 + 128 + c * 32 + lane] = cs2[c]; }
        asm volatile("bar.sync 1, %0;\n" :: "r"(DX_EPI_WARPS * 32) : "memory");
        if (nt > 0) {
            float tot = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < DX_EPI_WARPS; ++w8) tot += csum[w8 * 256 + tid];
            atomicAdd((tid < 128 ? a.dbeta : a.dgamma - 128) + tid, tot);
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
}  // namespace

cudaError_t savi_launch_dx_umma(const BwdArgs& a, const void* inputs, void* grad_inputs, bool overlap, cudaStream_t st) {
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
    // whole frames per CTA (one staging of the right-hand side) once there are >= 2 waves of frames; else >= 2 CTAs per frame
    x.tiles_per_cta = d.NTILE <= 8 ? ((d.B * d.T >= 2 * 148 || d.NTILE < 8) ? d.NTILE : 4) : 8;
    // (finer work items would shorten the tail left when the overlapped clip kernel ends, but measured slower: 1.724 vs 1.701 ms/step)
    if (savi_options().dx_tpc > 0) x.tiles_per_cta = savi_options().dx_tpc;                // development knob
    const int smem = dx_smem_total(d.I);
    cudaError_t e = cudaFuncSetAttribute(dx_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
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
int savi_dx_umma_smem_bytes(int I) { return dx_smem_total(I); }
