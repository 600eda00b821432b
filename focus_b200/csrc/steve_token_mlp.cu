// N1: the token-encoder tail that produces the slot-attention module's `inputs`,
//     emb_set = mlp(layer_norm(emb.permute(0,2,3,1).flatten(1,2)))        reference slowfast/models/STEVE/steve.py:307-309, 342-344
// (channels-first CNN map -> channels-last tokens -> LayerNorm -> Linear + ReLU + Linear), as ONE tcgen05 kernel, forward only
// (STEVE.encode / evaluation; C ABI: include/focus_steve.h).  The reference runs a permute copy, a LayerNorm pass, two cuBLAS
// GEMMs with a ReLU pass between them: ~10 passes over [B*T*N, C]; here the map is read once and the tokens written once.
//
// One CTA owns a tile of 128 consecutive tokens at a time (thread = token = TMEM lane):
//   load + LayerNorm (two coalesced passes over the tile's C channel rows; the second one hits L1/L2)
//     -> A operand [128 tokens][C] bf16, K-major SWIZZLE_128B blocks of 64 channels
//   tcgen05.mma  H = A . W1^T   (M = 128 tokens, N = C, K = C; W1 resident in shared memory as K-major B blocks)
//   epilogue 1   TMEM -> + b1, ReLU -> bf16 -> the same A buffer
//   tcgen05.mma  Y = H . W2^T
//   epilogue 2   TMEM -> + b2 -> staged [128][C] tile -> ONE bulk copy (TMA engine) to the contiguous token rows of `out`
// bf16 operands, fp32 accumulation: the accuracy class of the reference under bf16 autocast (its cuBLAS GEMMs round the same
// operands to bf16); the output feeds the slot-attention kernels, which store their token stream in bf16 anyway.
// HBM roofline: algorithmic bytes = C*4 (map, fp32) + C*out_bytes per token.
#include <cuda_fp16.h>
#include "savi_umma.cuh"
#include "focus_savi.h"
#include "focus_steve.h"

int savi_set_error(int code, const char* fmt, ...);

namespace tokmlp {
using namespace umma;

constexpr int TOK = 128;               // tokens per tile
constexpr int NTHR = 128;

struct Args {
    const float* emb;                  // [BT][C][HW]
    const float* ln_w; const float* ln_b; const float* b1; const float* b2;
    const unsigned char* wimg;         // W1 image then W2 image: [C/64 blocks][C rows][64] bf16, SWIZZLE_128B
    void* out;                         // [BT*HW][C]
    long long ntok; int HW; int C; float eps;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// fp32 [C][C] (out x in, nn.Linear layout) -> K-major SWIZZLE_128B bf16 blocks: block cb holds columns [64 cb, 64 cb + 64) of all C rows
__global__ void pack_w_kernel(const float* __restrict__ w1, const float* __restrict__ w2, unsigned char* __restrict__ img, int C) {
    const int n8 = C * (C >> 3);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n8; i += gridDim.x * blockDim.x) {
        const int m = i / n8, j = i - m * n8;
        const int r = j / (C >> 3), c0 = (j - r * (C >> 3)) * 8;
        const float* src = (m ? w2 : w1) + (size_t)r * C + c0;
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        unsigned char* blk = img + (size_t)m * C * C * 2 + (size_t)(c0 >> 6) * C * 128;
        *reinterpret_cast<uint4*>(blk + sw128_off(r, c0 & 63)) = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
    }
}

template <int C, typename OutT>
__global__ void __launch_bounds__(NTHR) token_mlp_kernel(const __grid_constant__ Args a) {
    constexpr int NCB = C / 64;
    constexpr int WB = C * C * 2;                            // bytes of one weight image
    constexpr int AB = TOK * C * 2;                          // bytes of the A operand tile
    constexpr int OB = TOK * C * (int)sizeof(OutT);          // bytes of the staged output tile
    constexpr uint32_t TCOLS = (2 * C <= 128) ? 128u : (2 * C <= 256) ? 256u : 512u;
    constexpr uint32_t IDESC = idesc_bf16(128, C, false, false);   // A K-major (tokens x channels), B K-major (outputs x channels)
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* sW = sm;                                  // W1 | W2 images
    unsigned char* sA = sm + 2 * WB;                         // A tile (LayerNorm'd tokens, then the hidden activations)
    // staged output rows: a bf16 tile is exactly as large as the A tile and is written only after the second product has
    // consumed it, so it aliases the A buffer (the next tile waits for the bulk store's reads before it writes A again)
    constexpr bool ALIAS_O = sizeof(OutT) == 2;
    unsigned char* sO = ALIAS_O ? sA : sA + AB;
    float* sP = reinterpret_cast<float*>(sA + AB + (ALIAS_O ? 0 : OB));   // ln_w | ln_b | b1 | b2
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * C);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    for (int i = tid; i < C; i += NTHR) { sP[i] = a.ln_w[i]; sP[C + i] = a.ln_b[i]; sP[2 * C + i] = a.b1[i]; sP[3 * C + i] = a.b2[i]; }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tslot;
    if (tid == 0) {                                          // both weight images: resident for the life of the CTA
        mbar_expect_tx(&bars[0], 2 * WB);
        for (int off = 0; off < 2 * WB; off += 32768) bulk_g2s(sW + off, a.wimg + off, (2 * WB - off) < 32768 ? (2 * WB - off) : 32768, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    uint32_t ph = 0;
    const long long ntile = (a.ntok + TOK - 1) / TOK;
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long g = tile * TOK + tid;                // this thread's token
        const bool valid = g < a.ntok;
        const long long bt = valid ? g / a.HW : 0;
        const int pix = valid ? (int)(g - bt * a.HW) : 0;
        const float* src = a.emb + (size_t)bt * C * a.HW + pix;      // channel c at src[c * HW]: a warp reads 32 consecutive pixels per channel
        // ---- LayerNorm (two-pass like torch: mean, then the centred second moment) -> A operand ----
        if constexpr (C <= 128) {
            // the token's whole channel vector lives in registers: ONE read of the map, C independent loads in flight per thread
            float x[C];
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = valid ? __ldg(src + (size_t)c * a.HW) : 0.f;
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) s += x[c];
            const float mean = s * (1.0f / C);
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) { const float d = x[c] - mean; v = fmaf(d, d, v); }
            const float rstd = rsqrtf(v * (1.0f / C) + a.eps);
            bulk_wait_read();                                // (sO of the previous tile has been read by its bulk store; sA is free: its MMA completed)
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 8) {
                float y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = (x[c0 + e] - mean) * rstd * sP[c0 + e] + sP[C + c0 + e];
                *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 6) * (TOK * 128) + sw128_off(tid, c0 & 63)) =
                    make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
            }
        } else {
            // wider maps: three coalesced passes over the tile's channel rows (the second and third hit L1 / L2)
            float s = 0.f;
#pragma unroll 32
            for (int c = 0; c < C; ++c) s += valid ? __ldg(src + (size_t)c * a.HW) : 0.f;
            const float mean = s * (1.0f / C);
            float v = 0.f;
#pragma unroll 32
            for (int c = 0; c < C; ++c) { const float d = (valid ? __ldg(src + (size_t)c * a.HW) : 0.f) - mean; v = fmaf(d, d, v); }
            const float rstd = rsqrtf(v * (1.0f / C) + a.eps);
            bulk_wait_read();
#pragma unroll 4
            for (int c0 = 0; c0 < C; c0 += 8) {
                float y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float raw = valid ? __ldg(src + (size_t)(c0 + e) * a.HW) : mean;
                    y[e] = (raw - mean) * rstd * sP[c0 + e] + sP[C + c0 + e];
                }
                *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 6) * (TOK * 128) + sw128_off(tid, c0 & 63)) =
                    make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        // ---- H = A . W1^T ----
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        mma_ss(tb, desc_kmajor(smem_u32(sA) + cb * (TOK * 128) + k4 * 32), desc_kmajor(smem_u32(sW) + cb * (C * 128) + k4 * 32),
                               IDESC, (cb | k4) != 0);
                mma_commit(&bars[1]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[1], ph); ph ^= 1u;
        fence_after_sync();
        // ---- epilogue 1: + b1, ReLU -> bf16 hidden activations into the A buffer ----
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float h[32];
            tmem_ld32(taddr + c0, h);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c0 + q * 8 + e * 2;
                    w[e] = pack2(fmaxf(h[q * 8 + e * 2] + sP[2 * C + c], 0.f), fmaxf(h[q * 8 + e * 2 + 1] + sP[2 * C + c + 1], 0.f));
                }
                const int c = c0 + q * 8;
                *reinterpret_cast<uint4*>(sA + (size_t)(c >> 6) * (TOK * 128) + sw128_off(tid, c & 63)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        // ---- Y = H . W2^T ----
        if (warp == 0) {
            fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        mma_ss(tb + C, desc_kmajor(smem_u32(sA) + cb * (TOK * 128) + k4 * 32), desc_kmajor(smem_u32(sW) + WB + cb * (C * 128) + k4 * 32),
                               IDESC, (cb | k4) != 0);
                mma_commit(&bars[1]);
            }
            __syncwarp();
        }
        mbar_wait(&bars[1], ph); ph ^= 1u;
        fence_after_sync();
        // ---- epilogue 2: + b2 -> staged row -> one bulk store of the tile's contiguous token rows ----
        OutT* orow = reinterpret_cast<OutT*>(sO) + (size_t)tid * C;
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float y[32];
            tmem_ld32(taddr + C + c0, y);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) y[e] += sP[3 * C + c0 + e];
            // rotate the 16-byte chunk order by the row index: the 128 rows are C*sizeof(OutT) apart (a multiple of 128 B), so
            // an un-rotated store would put every lane of a warp on the same banks
            if constexpr (sizeof(OutT) == 2) {
                uint4 ch[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    ch[q] = make_uint4(pack2(y[q * 8], y[q * 8 + 1]), pack2(y[q * 8 + 2], y[q * 8 + 3]), pack2(y[q * 8 + 4], y[q * 8 + 5]), pack2(y[q * 8 + 6], y[q * 8 + 7]));
                const int r = tid & 3;                       // (register arrays cannot be indexed dynamically: select the rotated chunk)
#pragma unroll
                for (int q0 = 0; q0 < 4; ++q0) {
                    const int q = (q0 + r) & 3;
                    uint4 v;
                    v.x = q == 0 ? ch[0].x : q == 1 ? ch[1].x : q == 2 ? ch[2].x : ch[3].x;
                    v.y = q == 0 ? ch[0].y : q == 1 ? ch[1].y : q == 2 ? ch[2].y : ch[3].y;
                    v.z = q == 0 ? ch[0].z : q == 1 ? ch[1].z : q == 2 ? ch[2].z : ch[3].z;
                    v.w = q == 0 ? ch[0].w : q == 1 ? ch[1].w : q == 2 ? ch[2].w : ch[3].w;
                    *reinterpret_cast<uint4*>(orow + c0 + q * 8) = v;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + c0 + q * 4) = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
            }
        }
        fence_async_smem();
        fence_before_sync();                                 // the accumulators have been read: the next tile's MMAs may overwrite them
        __syncthreads();
        if (tid == 0) {
            const long long rows = (a.ntok - tile * TOK) < TOK ? (a.ntok - tile * TOK) : TOK;
            bulk_s2g(reinterpret_cast<OutT*>(a.out) + (size_t)tile * TOK * C, sO, (uint32_t)(rows * C * sizeof(OutT)));
        }
    }
    bulk_wait_read();
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, TCOLS);
}

template <int C, typename OutT> static size_t smem_bytes() {
    return (size_t)2 * C * C * 2 + (size_t)TOK * C * 2 + (sizeof(OutT) == 2 ? 0 : (size_t)TOK * C * sizeof(OutT)) + 4 * C * sizeof(float) + 64;
}
template <int C, typename OutT> static cudaError_t launch(const Args& a, cudaStream_t st) {
    const size_t smem = smem_bytes<C, OutT>();
    cudaError_t e = cudaFuncSetAttribute(token_mlp_kernel<C, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long ntile = (a.ntok + TOK - 1) / TOK;
    const int per_sm = smem <= 113 * 1024 ? 2 : 1;           // two CTAs per SM cover each other's load / MMA / epilogue phases
    long long grid = 148LL * per_sm;
    if (grid > ntile) grid = ntile;
    token_mlp_kernel<C, OutT><<<(unsigned)grid, NTHR, smem, st>>>(a);
    return cudaGetLastError();
}
}  // namespace tokmlp

extern "C" int64_t steve_token_mlp_ws_bytes(int C) { return (int64_t)2 * C * C * 2; }

extern "C" int steve_token_mlp(const float* emb, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                               const float* w2, const float* b2, void* out, int out_dtype, int64_t BT, int HW, int C,
                               float ln_eps, void* ws, void* stream) {
    if (!emb || !ln_w || !ln_b || !w1 || !b1 || !w2 || !b2 || !out || !ws) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: null pointer");
    if (BT < 1 || HW < 1) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: BT, HW must be >= 1");
    if (C != 64 && C != 128 && C != 192) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: d_model C=%d not in {64, 128, 192}", C);
    if (out_dtype != SAVI_DTYPE_F32 && out_dtype != SAVI_DTYPE_BF16) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: unknown output dtype %d", out_dtype);
    if (C == 192 && out_dtype == SAVI_DTYPE_F32) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: C=192 supports bf16 output only (shared memory)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    tokmlp::pack_w_kernel<<<64, 256, 0, st>>>(w1, w2, reinterpret_cast<unsigned char*>(ws), C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "pack_w_kernel: %s", cudaGetErrorString(e));
    tokmlp::Args a;
    a.emb = emb; a.ln_w = ln_w; a.ln_b = ln_b; a.b1 = b1; a.b2 = b2; a.wimg = reinterpret_cast<const unsigned char*>(ws);
    a.out = out; a.ntok = BT * (long long)HW; a.HW = HW; a.C = C; a.eps = ln_eps;
    const bool f32 = out_dtype == SAVI_DTYPE_F32;
    if (C == 64) e = f32 ? tokmlp::launch<64, float>(a, st) : tokmlp::launch<64, __nv_bfloat16>(a, st);
    else if (C == 128) e = f32 ? tokmlp::launch<128, float>(a, st) : tokmlp::launch<128, __nv_bfloat16>(a, st);
    else e = tokmlp::launch<192, __nv_bfloat16>(a, st);
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "token_mlp_kernel: %s", cudaGetErrorString(e));
    return SAVI_OK;
}
