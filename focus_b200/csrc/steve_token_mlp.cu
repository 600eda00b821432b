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
    const float* b1f; const float* b2; // b1' = b1 + W1 ln_b (from the workspace), b2
    const unsigned char* wimg;         // W1' image then W2 image: [C/64 blocks][C rows][64] bf16, SWIZZLE_128B
    void* out;                         // [BT*HW][C]
    long long ntok; int HW; int C; float eps;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {   // bf16x2(max(a, 0), max(b, 0)), one instruction
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void prefetch_line_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" :: "l"(p)); }
__device__ __forceinline__ void bar_sync_n(int id, int nthreads) { asm volatile("bar.sync %0, %1;\n" :: "r"(id), "r"(nthreads) : "memory"); }
template <int C, int G> __host__ __device__ constexpr uint32_t tmem_cols() { return G * C <= 32 ? 32u : G * C <= 64 ? 64u : G * C <= 128 ? 128u : G * C <= 256 ? 256u : 512u; }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// fp32 [C][C] (out x in, nn.Linear layout) -> K-major SWIZZLE_128B bf16 blocks: block cb holds columns [64 cb, 64 cb + 64) of all C rows.
// The LayerNorm affine is folded into the first Linear (W1' = W1 diag(ln_w), b1' = b1 + W1 ln_b), so the kernel's A operand is the
// plain normalised token; b1' (fp32) follows the two images in the workspace.
__global__ void pack_w_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ ln_w,
                              const float* __restrict__ ln_b, const float* __restrict__ b1, unsigned char* __restrict__ img, int C) {
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // the token kernel's prologue and first-tile prefetch run beside this grid
    const int n8 = C * (C >> 3);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n8; i += gridDim.x * blockDim.x) {
        const int m = i / n8, j = i - m * n8;
        const int r = j / (C >> 3), c0 = (j - r * (C >> 3)) * 8;
        const float* src = (m ? w2 : w1) + (size_t)r * C + c0;
        float4 p = *reinterpret_cast<const float4*>(src), q = *reinterpret_cast<const float4*>(src + 4);
        if (m == 0) {
            const float4 g0 = *reinterpret_cast<const float4*>(ln_w + c0), g1 = *reinterpret_cast<const float4*>(ln_w + c0 + 4);
            p.x *= g0.x; p.y *= g0.y; p.z *= g0.z; p.w *= g0.w; q.x *= g1.x; q.y *= g1.y; q.z *= g1.z; q.w *= g1.w;
        }
        unsigned char* blk = img + (size_t)m * C * C * 2 + (size_t)(c0 >> 6) * C * 128;
        *reinterpret_cast<uint4*>(blk + sw128_off(r, c0 & 63)) = make_uint4(pack2(p.x, p.y), pack2(p.z, p.w), pack2(q.x, q.y), pack2(q.z, q.w));
    }
    float* b1f = reinterpret_cast<float*>(img + (size_t)2 * C * C * 2);
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < C; o += gridDim.x * blockDim.x) {
        float acc = b1[o];
        for (int c = 0; c < C; ++c) acc = fmaf(w1[(size_t)o * C + c], ln_b[c], acc);
        b1f[o] = acc;
    }
}

// G independent groups of 128 threads per CTA share the resident weight images; each group owns an A tile, C TMEM columns, an mbarrier
// and a named barrier, and walks its own sequence of token tiles, so the groups (and the CTAs of an SM) cover each other's load /
// LayerNorm / MMA / epilogue latencies.
// HWT = tokens per frame when known at compile time (0: read a.HW): the channel stride becomes an immediate offset of each load
// (a run-time stride costs ~6 address instructions per load, a third of the kernel's instruction count) and the frame index a shift.
template <int C, typename OutT, int G, int HWT>
__global__ void __launch_bounds__(G * NTHR) token_mlp_kernel(const __grid_constant__ Args a) {
    constexpr int NCB = C / 64;
    const int HW = HWT ? HWT : a.HW;
    constexpr int WB = C * C * 2;                            // bytes of one weight image
    constexpr int AB = TOK * C * 2;                          // bytes of one A operand tile
    constexpr int OB = TOK * C * (int)sizeof(OutT);          // bytes of one staged output tile
    constexpr uint32_t TCOLS = tmem_cols<C, G>();
    constexpr uint32_t IDESC = idesc_bf16(128, C, false, false);   // A K-major (tokens x channels), B K-major (outputs x channels)
    extern __shared__ __align__(1024) unsigned char sm[];
    const int tid = threadIdx.x, lt = tid & (NTHR - 1);
    // warp index through a lane-0 broadcast: ptxas then knows it (and the tile loop below) is warp-uniform and keeps the loads'
    // memory descriptor in uniform registers instead of re-materialising it (2 R2UR per load)
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp_u >> 2, lwarp = warp_u & 3;
    unsigned char* sW = sm;                                  // W1' | W2 images
    unsigned char* sA = sm + 2 * WB + grp * AB;              // this group's A tile (normalised tokens, then the hidden activations)
    // staged output rows: a bf16 tile is exactly as large as the A tile and is written only after the second product has
    // consumed it, so it aliases the A buffer (the next tile waits for the bulk store's reads before it writes A again)
    constexpr bool ALIAS_O = sizeof(OutT) == 2;
    unsigned char* sO = ALIAS_O ? sA : sm + 2 * WB + G * AB + grp * OB;
    float* sP = reinterpret_cast<float*>(sm + 2 * WB + G * AB + (ALIAS_O ? 0 : G * OB));   // b1' | b2
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * C);                              // [0] weights, [1 + grp] this group's MMAs
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 1 + G);

    if (tid == 0) { for (int i = 0; i <= G; ++i) mbar_init(&bars[i], 1); mbar_init_fence(); }
    if (tid < 32) tmem_alloc(tslot, TCOLS);
    const long long ntile = (a.ntok + TOK - 1) / TOK;
    const long long tstride = (long long)gridDim.x * G;      // neighbouring groups take neighbouring tiles
    const bool i32 = a.ntok <= 0x7fffffffLL;                 // (a 64-bit division costs ~100 instructions)
    // pull tile `t` of the map into L2: thread = channel row, one request per 128-byte line (token 127 covers the row's last line
    // when the tile is not line-aligned or straddles a frame)
    auto prefetch_tile = [&](long long t) {
        const long long g2 = t * TOK;
        if (g2 >= a.ntok) return;
        const long long bt2 = i32 ? (long long)((int)g2 / HW) : g2 / HW;
        const int pix2 = (int)(g2 - bt2 * HW);
        const int last = (a.ntok - 1 - g2) < TOK - 1 ? (int)(a.ntok - 1 - g2) : TOK - 1;
        for (int c = lt; c < C; c += NTHR)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int off = (j < 4 && j * 32 < last) ? j * 32 : last;
                long long b = bt2; int p = pix2 + off;
                while (p >= HW) { p -= HW; ++b; }
                prefetch_line_l2(a.emb + ((size_t)b * C + c) * HW + p);
            }
    };
    prefetch_tile((long long)blockIdx.x * G + grp);          // the map does not depend on the packing grid: fetch while it finishes
    asm volatile("griddepcontrol.wait;\n" ::: "memory");     // (programmatic launch: the weight images and b1' are complete past here)
    for (int i = tid; i < C; i += G * NTHR) { sP[i] = a.b1f[i]; sP[C + i] = a.b2[i]; }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tslot + (uint32_t)(grp * C);        // this group's accumulator columns (H, then Y over it)
    if (tid == 0) {                                          // both weight images: resident for the life of the CTA
        mbar_expect_tx(&bars[0], 2 * WB);
        for (int off = 0; off < 2 * WB; off += 32768) bulk_g2s(sW + off, a.wimg + off, (2 * WB - off) < 32768 ? (2 * WB - off) : 32768, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    uint64_t* mbar = &bars[1 + grp];
    uint32_t ph = 0;
    const uint32_t taddr = tb + ((uint32_t)(lwarp * 32) << 16);
    for (long long tile = (long long)blockIdx.x * G + grp; tile < ntile; tile += tstride) {
        // this thread's token; rows past the end (last tile only) recompute the last token and are not stored
        const long long g = (tile * TOK + lt) < a.ntok ? tile * TOK + lt : a.ntok - 1;
        const long long bt = i32 ? (long long)((int)g / HW) : g / HW;
        const int pix = (int)(g - bt * HW);
        const float* src = a.emb + (size_t)bt * C * HW + pix;      // channel c at src[c * HW]: a warp reads 32 consecutive pixels per channel
        // All groups run their load / LayerNorm / MMA / epilogue phases nearly in lock-step (same start, same durations), so without
        // help HBM idles while they compute: behind this tile's loads, pull the group's NEXT tile into L2.
        auto prefetch_next = [&]() { prefetch_tile(tile + tstride); };
        // ---- LayerNorm (two-pass like torch: mean, then the centred second moment) -> A operand ----
        if constexpr (C <= 128) {
            // the token's whole channel vector lives in registers: ONE read of the map, C independent loads in flight per thread
            float x[C];
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * HW);
            prefetch_next();
            float s[4] = {0.f, 0.f, 0.f, 0.f};               // four partial sums: the dependent-add chain is C/4 long
#pragma unroll
            for (int c = 0; c < C; ++c) s[c & 3] += x[c];
            const float mean = ((s[0] + s[1]) + (s[2] + s[3])) * (1.0f / C);
            float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < C; ++c) { x[c] -= mean; v[c & 3] = fmaf(x[c], x[c], v[c & 3]); }
            const float rstd = rsqrtf(((v[0] + v[1]) + (v[2] + v[3])) * (1.0f / C) + a.eps);
            if (lt == 0) bulk_wait_read();                   // the previous tile's bulk store (issued by this thread) has read sO ...
            bar_sync_n(1 + grp, NTHR);                       // ... so every thread of the group may overwrite the aliased A tile
#pragma unroll
            for (int c0 = 0; c0 < C; c0 += 8)
                *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 6) * (TOK * 128) + sw128_off(lt, c0 & 63)) =
                    make_uint4(pack2(x[c0] * rstd, x[c0 + 1] * rstd), pack2(x[c0 + 2] * rstd, x[c0 + 3] * rstd),
                               pack2(x[c0 + 4] * rstd, x[c0 + 5] * rstd), pack2(x[c0 + 6] * rstd, x[c0 + 7] * rstd));
        } else {
            // wider maps: three coalesced passes over the tile's channel rows (the second and third hit L1 / L2)
            float s = 0.f;
#pragma unroll 32
            for (int c = 0; c < C; ++c) s += __ldg(src + (size_t)c * HW);
            prefetch_next();
            const float mean = s * (1.0f / C);
            float v = 0.f;
#pragma unroll 32
            for (int c = 0; c < C; ++c) { const float d = __ldg(src + (size_t)c * HW) - mean; v = fmaf(d, d, v); }
            const float rstd = rsqrtf(v * (1.0f / C) + a.eps);
            if (lt == 0) bulk_wait_read();
            bar_sync_n(1 + grp, NTHR);
#pragma unroll 4
            for (int c0 = 0; c0 < C; c0 += 8) {
                float y[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = (__ldg(src + (size_t)(c0 + e) * HW) - mean) * rstd;
                *reinterpret_cast<uint4*>(sA + (size_t)(c0 >> 6) * (TOK * 128) + sw128_off(lt, c0 & 63)) =
                    make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
            }
        }
        fence_async_smem();
        fence_before_sync();
        bar_sync_n(1 + grp, NTHR);
        // ---- H = A . W1'^T ----
        if (lwarp == 0) {
            fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        mma_ss(tb, desc_kmajor(smem_u32(sA) + cb * (TOK * 128) + k4 * 32), desc_kmajor(smem_u32(sW) + cb * (C * 128) + k4 * 32),
                               IDESC, (cb | k4) != 0);
                mma_commit(mbar);
            }
            __syncwarp();
        }
        mbar_wait(mbar, ph); ph ^= 1u;
        fence_after_sync();
        // ---- epilogue 1: + b1', ReLU -> bf16 hidden activations into the A buffer ----
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float h[32];
            tmem_ld32(taddr + c0, h);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 b0 = *reinterpret_cast<const float4*>(sP + c0 + q * 8), b1 = *reinterpret_cast<const float4*>(sP + c0 + q * 8 + 4);
                const float* hq = h + q * 8;
                const int c = c0 + q * 8;
                *reinterpret_cast<uint4*>(sA + (size_t)(c >> 6) * (TOK * 128) + sw128_off(lt, c & 63)) =
                    make_uint4(pack2_relu(hq[0] + b0.x, hq[1] + b0.y), pack2_relu(hq[2] + b0.z, hq[3] + b0.w),
                               pack2_relu(hq[4] + b1.x, hq[5] + b1.y), pack2_relu(hq[6] + b1.z, hq[7] + b1.w));
            }
        }
        fence_async_smem();
        fence_before_sync();                                 // H has been read: the second product may overwrite its columns
        bar_sync_n(1 + grp, NTHR);
        // ---- Y = H . W2^T ----
        if (lwarp == 0) {
            fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int cb = 0; cb < NCB; ++cb)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        mma_ss(tb, desc_kmajor(smem_u32(sA) + cb * (TOK * 128) + k4 * 32), desc_kmajor(smem_u32(sW) + WB + cb * (C * 128) + k4 * 32),
                               IDESC, (cb | k4) != 0);
                mma_commit(mbar);
            }
            __syncwarp();
        }
        mbar_wait(mbar, ph); ph ^= 1u;
        fence_after_sync();
        // ---- epilogue 2: + b2 -> staged row -> one bulk store of the tile's contiguous token rows ----
        OutT* orow = reinterpret_cast<OutT*>(sO) + (size_t)lt * C;
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
            float y[32];
            tmem_ld32(taddr + c0, y);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = *reinterpret_cast<const float4*>(sP + C + c0 + q * 4);
                y[q * 4] += b.x; y[q * 4 + 1] += b.y; y[q * 4 + 2] += b.z; y[q * 4 + 3] += b.w;
            }
            // rotate the 16-byte chunk order by the row index: the 128 rows are C*sizeof(OutT) apart (a multiple of 128 B), so
            // an un-rotated store would put every lane of a warp on the same banks
            if constexpr (sizeof(OutT) == 2) {
                uint4 ch[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    ch[q] = make_uint4(pack2(y[q * 8], y[q * 8 + 1]), pack2(y[q * 8 + 2], y[q * 8 + 3]), pack2(y[q * 8 + 4], y[q * 8 + 5]), pack2(y[q * 8 + 6], y[q * 8 + 7]));
                // rotate by r = lt & 3 with two conditional-swap stages (plain selects: register arrays cannot be indexed
                // dynamically, and a chain of ternaries compiles to branches)
                const bool r1 = lt & 1, r2 = lt & 2;
                uint4 s1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 p = ch[q], n = ch[(q + 1) & 3];
                    s1[q] = make_uint4(r1 ? n.x : p.x, r1 ? n.y : p.y, r1 ? n.z : p.z, r1 ? n.w : p.w);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 p = s1[q], n = s1[(q + 2) & 3];
                    *reinterpret_cast<uint4*>(orow + c0 + ((q + (lt & 3)) & 3) * 8) =
                        make_uint4(r2 ? n.x : p.x, r2 ? n.y : p.y, r2 ? n.z : p.z, r2 ? n.w : p.w);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + c0 + q * 4) = make_float4(y[q * 4], y[q * 4 + 1], y[q * 4 + 2], y[q * 4 + 3]);
            }
        }
        fence_async_smem();
        fence_before_sync();                                 // the accumulators have been read: the next tile's MMAs may overwrite them
        bar_sync_n(1 + grp, NTHR);
        if (lt == 0) {
            const long long rows = (a.ntok - tile * TOK) < TOK ? (a.ntok - tile * TOK) : TOK;
            bulk_s2g(reinterpret_cast<OutT*>(a.out) + (size_t)tile * TOK * C, sO, (uint32_t)(rows * C * sizeof(OutT)));
        }
    }
    if (lt == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tslot, TCOLS);
}

template <int C, typename OutT, int G> static size_t smem_bytes() {
    return (size_t)2 * C * C * 2 + (size_t)G * TOK * C * 2 + (sizeof(OutT) == 2 ? 0 : (size_t)G * TOK * C * sizeof(OutT)) + 2 * C * sizeof(float) + 64;
}
template <int C, typename OutT, int G, int HWT = 0> static cudaError_t launch(const Args& a, cudaStream_t st) {
    const size_t smem = smem_bytes<C, OutT, G>();
    cudaError_t e = cudaFuncSetAttribute(token_mlp_kernel<C, OutT, G, HWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long ntile = (a.ntok + TOK - 1) / TOK;
    int per_sm = (int)((227 * 1024) / (smem + 1024));        // resident CTAs per SM: shared memory, TMEM columns
    if (per_sm > (int)(512 / tmem_cols<C, G>())) per_sm = 512 / tmem_cols<C, G>();
    if (per_sm < 1) per_sm = 1;
    long long grid = 148LL * per_sm;
    if (grid > (ntile + G - 1) / G) grid = (ntile + G - 1) / G;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(G * NTHR); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // starts beside pack_w_kernel; griddepcontrol.wait orders the data
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, token_mlp_kernel<C, OutT, G, HWT>, a);
}
}  // namespace tokmlp

extern "C" int64_t steve_token_mlp_ws_bytes(int C) { return (int64_t)2 * C * C * 2 + (int64_t)C * 4; }

extern "C" int steve_token_mlp(const float* emb, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                               const float* w2, const float* b2, void* out, int out_dtype, int64_t BT, int HW, int C,
                               float ln_eps, void* ws, void* stream) {
    if (!emb || !ln_w || !ln_b || !w1 || !b1 || !w2 || !b2 || !out || !ws) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: null pointer");
    if (BT < 1 || HW < 1) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: BT, HW must be >= 1");
    if (C != 64 && C != 128 && C != 192) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: d_model C=%d not in {64, 128, 192}", C);
    if (out_dtype != SAVI_DTYPE_F32 && out_dtype != SAVI_DTYPE_BF16) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: unknown output dtype %d", out_dtype);
    if (C == 192 && out_dtype == SAVI_DTYPE_F32) return savi_set_error(SAVI_EINVAL, "steve_token_mlp: C=192 supports bf16 output only (shared memory)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    tokmlp::pack_w_kernel<<<64, 256, 0, st>>>(w1, w2, ln_w, ln_b, b1, reinterpret_cast<unsigned char*>(ws), C);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "pack_w_kernel: %s", cudaGetErrorString(e));
    tokmlp::Args a;
    a.emb = emb; a.b2 = b2; a.wimg = reinterpret_cast<const unsigned char*>(ws);
    a.b1f = reinterpret_cast<const float*>(a.wimg + (size_t)2 * C * C * 2);
    a.out = out; a.ntok = BT * (long long)HW; a.HW = HW; a.C = C; a.eps = ln_eps;
    const bool f32 = out_dtype == SAVI_DTYPE_F32;
    // groups per CTA: as many as shared memory (A / staging tile per group), TMEM (C columns per group) and registers allow
    if (C == 64) e = f32 ? tokmlp::launch<64, float, 2>(a, st) : tokmlp::launch<64, __nv_bfloat16, 2>(a, st);
    else if (C == 128) e = f32 ? tokmlp::launch<128, float, 1>(a, st)
                             : HW == 1024 ? tokmlp::launch<128, __nv_bfloat16, 3, 1024>(a, st)       // MOVi 128 x 128 video: 32 x 32 tokens
                             : HW == 4096 ? tokmlp::launch<128, __nv_bfloat16, 3, 4096>(a, st)       // 64 x 64 tokens
                                          : tokmlp::launch<128, __nv_bfloat16, 3>(a, st);
    else e = HW == 4096 ? tokmlp::launch<192, __nv_bfloat16, 1, 4096>(a, st) : tokmlp::launch<192, __nv_bfloat16, 1>(a, st);
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "token_mlp_kernel: %s", cudaGetErrorString(e));
    return SAVI_OK;
}
