// d_inputs for the bf16 mode, token-parallel (outside the sequential recurrence):
//   dxhat[n,:] = sum_i  dL_i[n,:] . qk_i  +  W_i[n,:] . dUx_i          (folded form)
//   d_inputs   = LayerNorm backward of dxhat through norm_inputs (steve.py:60), plus d gamma / d beta.
// The recurrence kernel only stages the [tokens x slots] coefficients dL_i, W_i (bf16, token-contiguous)
// and the tiny per-step right-hand sides; this kernel does the dense contraction
// [128 tokens x (I*2*KC)] . [(I*2*KC) x D] on tensor cores and fuses the LayerNorm backward as its
// epilogue, so dxhat never touches HBM.
#pragma once
#include "savi_token_mma.cuh"

struct DxArgs {
    const bf16* x;            // inputs [B*T, N, D]
    const float2* stats;      // [B*T*N] mean, rstd
    const bf16* coef;         // [B*T][I][2][KC][N]
    const float* qk;          // saved field array: rows ((t*I+i)*B + b)*K + k, width D
    const float* dux;         // bwd staging field array, same row order
    const float* gamma;       // norm_inputs.weight
    bf16* dx;                 // d_inputs [B*T, N, D]
    float* dgamma; float* dbeta;
    int B, T, N, D, K, I, KC, tiles_per_cta;
    int IG;                   // iterations staged in shared memory at a time (== I unless the operands of all I do not fit)
};

// shared memory with IG iterations' right-hand side rows and coefficient rows staged at a time
__host__ __device__ __forceinline__ size_t dx_smem_bytes(int IG, int KC, int D) {
    const size_t rows = (size_t)IG * 2 * KC;
    return rows * tmma_xs(D) + rows * TMMA_ATS + (size_t)TMMA_TN * tmma_xs(D) + 2 * (size_t)D * 4 + 16;
}
// the largest group of iterations whose operands fit `max_bytes` (0: not even one iteration fits)
__host__ __device__ __forceinline__ int dx_pick_ig(int I, int KC, int D, size_t max_bytes) {
    int ig = I;
    while (ig > 0 && dx_smem_bytes(ig, KC, D) > max_bytes) --ig;
    return ig;
}

constexpr int DX_NT = 256;      // 8 warps x 16 tokens = one 128-token tile

template <int ND>   // n-tiles of 8 features held per warp: D <= 8*ND
__global__ void __launch_bounds__(DX_NT, 1) dx_finalize_kernel(const __grid_constant__ DxArgs a) {
    extern __shared__ float4 smem4[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(smem4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3, mi = lane >> 3, rr = lane & 7;
    const int D = a.D, N = a.N, K = a.K, KC = a.KC, I = a.I, xs = tmma_xs(D);
    const int IG = a.IG, rows = IG * 2 * KC, nd = D >> 3;          // rows staged at a time
    unsigned char* rs = smem;                                       // R  [rows][D+8] bf16
    unsigned char* cs = rs + (size_t)rows * xs;                     // coef tile [rows][TN+8] bf16
    unsigned char* xo = cs + (size_t)rows * TMMA_ATS;               // x tile, later the d_inputs tile [TN][D+8]
    float* red = reinterpret_cast<float*>(xo + (size_t)TMMA_TN * xs);   // [2][D] d gamma, d beta
    const int f = blockIdx.y, b = f / a.T, t = f - b * a.T;
    for (int i = tid; i < 2 * D; i += DX_NT) red[i] = 0.f;
    // right-hand side rows of iterations [i0, i0 + ig): for each iteration qk_i (KC rows, zero beyond K) then dUx_i
    auto stage_rhs = [&](int i0, int ig) {
        const int d4 = D >> 2;
        for (int idx = tid; idx < ig * 2 * KC * d4; idx += DX_NT) {
            const int r = idx / d4, c = (idx - r * d4) * 4;
            const int i = r / (2 * KC), rem = r - i * 2 * KC, which = rem / KC, k = rem - which * KC;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K) {
                const float* src = (which == 0 ? a.qk : a.dux) + ((((size_t)t * I + i0 + i) * a.B + b) * K + k) * D + c;
                v = ld4(src);
            }
            uint2 p; p.x = pack_bf16(v.x, v.y); p.y = pack_bf16(v.z, v.w);
            *reinterpret_cast<uint2*>(rs + (size_t)r * xs + c * 2) = p;
        }
    };
    if (IG == I) stage_rhs(0, I);                                   // everything fits: staged once per CTA
    const bf16* coef_f = a.coef + (size_t)f * I * 2 * KC * N;
    const int tile0 = blockIdx.x * a.tiles_per_cta;
    for (int tt = 0; tt < a.tiles_per_cta; ++tt) {
        const int n0 = (tile0 + tt) * TMMA_TN;
        if (n0 >= N) break;
        const int tn = min(TMMA_TN, N - n0);
        float acc[ND][4];
#pragma unroll
        for (int n = 0; n < ND; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
        for (int i0 = 0; i0 < I; i0 += IG) {                        // one pass per group of iterations (a single pass when IG == I)
            const int ig = min(IG, I - i0), rows_g = ig * 2 * KC;
            __syncthreads();                                        // the previous pass / tile no longer reads the staging buffers
            if (IG != I) stage_rhs(i0, ig);
            {   // coefficient rows (contiguous 256 B per row) and, in the first pass, the x tile
                const int cpr = TMMA_TN / 8;
                const bf16* coef_g = coef_f + (size_t)i0 * 2 * KC * N;
                for (int idx = tid; idx < rows_g * cpr; idx += DX_NT) {
                    const int r = idx / cpr, c = idx - r * cpr;
                    unsigned char* dst = cs + (size_t)r * TMMA_ATS + c * 16;
                    if (c * 8 < tn) cp_async16(dst, coef_g + (size_t)r * N + n0 + c * 8);     // N % 8 == 0: whole chunks valid
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
                }
                if (i0 == 0) tmma_issue_tile<DX_NT>(xo, a.x + (size_t)f * N * D, n0, tn, D);
                cp_async_wait_all();
            }
            __syncthreads();
            for (int k0 = 0; k0 < rows_g; k0 += 16) {
                uint32_t af[4];
                ldsm_x4_t(af, cs + (size_t)(k0 + (mi >> 1) * 8 + rr) * TMMA_ATS + (warp * 16 + (mi & 1) * 8) * 2);
#pragma unroll
                for (int n2 = 0; n2 < ND / 2; ++n2) {
                    if (n2 * 2 < nd) {
                        uint32_t bf[4];
                        ldsm_x4_t(bf, rs + (size_t)(k0 + (mi & 1) * 8 + rr) * xs + (n2 * 16 + (mi >> 1) * 8) * 2);
                        mma16816(acc[n2 * 2], af[0], af[1], af[2], af[3], bf[0], bf[1]);
                        mma16816(acc[n2 * 2 + 1], af[0], af[1], af[2], af[3], bf[2], bf[3]);
                    }
                }
            }
        }
        // ---- LayerNorm backward epilogue on this warp's 16 tokens ----
        const int r0 = warp * 16 + g, r1 = r0 + 8;
        const bool v0 = r0 < tn, v1 = r1 < tn;
        const float2 st0 = v0 ? a.stats[(size_t)f * N + n0 + r0] : make_float2(0.f, 0.f);
        const float2 st1 = v1 ? a.stats[(size_t)f * N + n0 + r1] : make_float2(0.f, 0.f);
        float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
        for (int n = 0; n < ND; ++n) {
            if (n < nd) {
                const int c = n * 8 + q * 2;
                const float2 gm = __ldg(reinterpret_cast<const float2*>(a.gamma + c));
                const float2 x0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xo + (size_t)r0 * xs + c * 2));
                const float2 x1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(xo + (size_t)r1 * xs + c * 2));
                const float z00 = (x0.x - st0.x) * st0.y, z01 = (x0.y - st0.x) * st0.y;
                const float z10 = (x1.x - st1.x) * st1.y, z11 = (x1.y - st1.x) * st1.y;
                // d gamma / d beta partials over this thread's two rows, reduced over the 8 row-lanes below
                float dg0 = (v0 ? acc[n][0] * z00 : 0.f) + (v1 ? acc[n][2] * z10 : 0.f);
                float dg1 = (v0 ? acc[n][1] * z01 : 0.f) + (v1 ? acc[n][3] * z11 : 0.f);
                float db0 = (v0 ? acc[n][0] : 0.f) + (v1 ? acc[n][2] : 0.f);
                float db1 = (v0 ? acc[n][1] : 0.f) + (v1 ? acc[n][3] : 0.f);
                dg0 = colsum(dg0); dg1 = colsum(dg1); db0 = colsum(db0); db1 = colsum(db1);
                if (g == 0) {
                    atomicAdd(red + c, dg0); atomicAdd(red + c + 1, dg1);
                    atomicAdd(red + D + c, db0); atomicAdd(red + D + c + 1, db1);
                }
                // dz = dxhat * gamma (kept in acc), row sums
                acc[n][0] *= gm.x; acc[n][1] *= gm.y; acc[n][2] *= gm.x; acc[n][3] *= gm.y;
                s1a += acc[n][0] + acc[n][1]; s2a = fmaf(acc[n][0], z00, fmaf(acc[n][1], z01, s2a));
                s1b += acc[n][2] + acc[n][3]; s2b = fmaf(acc[n][2], z10, fmaf(acc[n][3], z11, s2b));
            }
        }
        s1a += __shfl_xor_sync(0xffffffffu, s1a, 1); s1a += __shfl_xor_sync(0xffffffffu, s1a, 2);
        s2a += __shfl_xor_sync(0xffffffffu, s2a, 1); s2a += __shfl_xor_sync(0xffffffffu, s2a, 2);
        s1b += __shfl_xor_sync(0xffffffffu, s1b, 1); s1b += __shfl_xor_sync(0xffffffffu, s1b, 2);
        s2b += __shfl_xor_sync(0xffffffffu, s2b, 1); s2b += __shfl_xor_sync(0xffffffffu, s2b, 2);
        const float invD = 1.0f / (float)D;
        s1a *= invD; s2a *= invD; s1b *= invD; s2b *= invD;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < ND; ++n) {
            if (n < nd) {
                const int c = n * 8 + q * 2;
                unsigned char* p0 = xo + (size_t)r0 * xs + c * 2;
                unsigned char* p1 = xo + (size_t)r1 * xs + c * 2;
                const float2 x0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p0));
                const float2 x1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p1));
                const float z00 = (x0.x - st0.x) * st0.y, z01 = (x0.y - st0.x) * st0.y;
                const float z10 = (x1.x - st1.x) * st1.y, z11 = (x1.y - st1.x) * st1.y;
                // each thread overwrites exactly the x elements it read: the tile becomes the d_inputs tile
                *reinterpret_cast<uint32_t*>(p0) = pack_bf16(st0.y * (acc[n][0] - s1a - z00 * s2a), st0.y * (acc[n][1] - s1a - z01 * s2a));
                *reinterpret_cast<uint32_t*>(p1) = pack_bf16(st1.y * (acc[n][2] - s1b - z10 * s2b), st1.y * (acc[n][3] - s1b - z11 * s2b));
            }
        }
        __syncthreads();
        {   // coalesced 16-byte stores of the d_inputs tile
            const int chunks = D >> 3;
            unsigned char* gb = reinterpret_cast<unsigned char*>(a.dx + ((size_t)f * N + n0) * D);
            for (int idx = tid; idx < tn * chunks; idx += DX_NT) {
                const int r = idx / chunks, c = idx - r * chunks;
                *reinterpret_cast<uint4*>(gb + (size_t)r * D * 2 + c * 16) = *reinterpret_cast<const uint4*>(xo + (size_t)r * xs + c * 16);
            }
        }
    }
    __syncthreads();
    for (int c = tid; c < D; c += DX_NT) { atomicAdd(a.dgamma + c, red[c]); atomicAdd(a.dbeta + c, red[D + c]); }
}
