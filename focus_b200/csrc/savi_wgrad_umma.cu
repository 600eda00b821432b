// Weight gradients of the tcgen05 path: dW[o][c] += alpha * sum_r dY[r][o] X[r][c] for every weight matrix in one
// launch (r runs over all steps and clips: the "one tall GEMM per weight" layout of savi_layout.h).
// Per CTA: one [128 x 128] output tile and a slice of the rows.  Eight loader warps stream 64-row chunks of dY and X
// (fp32, row-major), split them into bf16 hi / lo and write them as MN-major SWIZZLE_128B operand blocks (the row
// index is the contraction index, so no transposition is needed); one warp issues three tcgen05.mma per k-step
// (hi.hi + hi.lo + lo.hi, M = N = 128) into a TMEM accumulator; the epilogue adds the tile into the flat gradient buffer.
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"

using namespace umma;
typedef __nv_bfloat16 bf16;

namespace {
constexpr int WU_LOADERS = 256, WU_THREADS = 288;     // warps 0-7 load + convert, warp 8 issues
constexpr int WU_CH = 32;                             // rows per chunk (32: 32 KB stages, two CTAs per SM hide each other's load latency)
constexpr int WU_CB = WU_CH * 128;                    // bytes of one 64-column block of a chunk
constexpr int WU_SUB = 2 * WU_CB;                     // one operand: [WU_CH rows][2 x 64 columns] bf16
constexpr int WU_STAGE = 4 * WU_SUB;                  // A hi | A lo | B hi | B lo
constexpr int WU_LD = WU_CH * 16 / WU_LOADERS;        // 8-float groups per loader thread and operand
constexpr uint32_t IDESC_MN_MN_128 = idesc_bf16(128, 128, true, true);

struct WUArgs { WgradArgs wa; int chunks_per_cta;
                const int* done; int done_target; };      // overlapped launch: CTA counter of the backward clip kernel to wait for

__device__ __forceinline__ void split8(const float4& p, const float4& q, uint4& hi, uint4& lo) {
    const float v[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh); l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(WU_THREADS, 2) wgrad_umma_kernel(const __grid_constant__ WUArgs ua) {
    extern __shared__ __align__(1024) unsigned char sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    const WgradArgs& wa = ua.wa;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- locate this CTA's work item: (job, output tile, row slice) ----
    int job = -1, tile = 0, slice = 0;
    {
        int rem = blockIdx.x;
        for (int j = 0; j < wa.njobs; ++j) {
            const int tiles = (wa.job[j].O >> 7) * (wa.job[j].C >> 7);
            const int chunks = (wa.job[j].R + WU_CH - 1) / WU_CH;
            const int slices = (chunks + ua.chunks_per_cta - 1) / ua.chunks_per_cta;
            if (rem < tiles * slices) { job = j; tile = rem / slices; slice = rem - tile * slices; break; }
            rem -= tiles * slices;
        }
    }
    if (job < 0) return;
    if (ua.done && tid == 0) {          // launched early (programmatic dependent of d_inputs): the clip kernel's records must be complete
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(ua.done) : "memory");
            if (v < ua.done_target) __nanosleep(256);
        } while (v < ua.done_target);
    }
    const WgradJob& jb = wa.job[job];
    const int tc = jb.C >> 7, o0 = (tile / tc) * 128, c0 = (tile % tc) * 128;
    const int r_begin = slice * ua.chunks_per_cta * WU_CH, r_end = min(jb.R, r_begin + ua.chunks_per_cta * WU_CH);
    const int nch = (r_end - r_begin + WU_CH - 1) / WU_CH;

    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * WU_STAGE);      // full[2], empty[2], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    if (tid == 0) {
        mbar_init(&bars[0], 8); mbar_init(&bars[1], 8); mbar_init(&bars[2], 1); mbar_init(&bars[3], 1); mbar_init(&bars[4], 1);
        mbar_init_fence();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp < 8) {
        // ---- loaders: fp32 rows -> bf16 hi / lo MN-major blocks [64 rows][2 x 64 columns] ----
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1;
            unsigned char* base = sm + st * WU_STAGE;
            const int r0 = r_begin + ch * WU_CH;
            float4 pa[WU_LD][2], pb[WU_LD][2];
#pragma unroll
            for (int i = 0; i < WU_LD; ++i) {                             // all loads of the chunk are in flight before the stage wait
                const int u = tid + i * WU_LOADERS, r = u >> 4, cg = u & 15;
                const bool ok = r0 + r < r_end;
                const float* ya = jb.dY + (size_t)(r0 + r) * jb.ldy + o0 + cg * 8;
                const float* xa = jb.X + (size_t)(r0 + r) * jb.ldx + c0 + cg * 8;
                pa[i][0] = ok ? ld4(ya) : make_float4(0.f, 0.f, 0.f, 0.f); pa[i][1] = ok ? ld4(ya + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                pb[i][0] = ok ? ld4(xa) : make_float4(0.f, 0.f, 0.f, 0.f); pb[i][1] = ok ? ld4(xa + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&bars[2 + st], ((ch >> 1) & 1u) ^ 1u);               // the MMAs that read this stage two chunks ago are done
#pragma unroll
            for (int i = 0; i < WU_LD; ++i) {
                const int u = tid + i * WU_LOADERS, r = u >> 4, cg = u & 15;
                const uint32_t off = (uint32_t)(cg >> 3) * (uint32_t)WU_CB + (uint32_t)r * 128u + ((((uint32_t)cg & 7u) ^ ((uint32_t)r & 7u)) << 4);
                uint4 hi, lo;
                split8(pa[i][0], pa[i][1], hi, lo);
                *reinterpret_cast<uint4*>(base + off) = hi; *reinterpret_cast<uint4*>(base + WU_SUB + off) = lo;
                split8(pb[i][0], pb[i][1], hi, lo);
                *reinterpret_cast<uint4*>(base + 2 * WU_SUB + off) = hi; *reinterpret_cast<uint4*>(base + 3 * WU_SUB + off) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[st]);
        }
        // ---- epilogue (warps 0-3: thread = output row o) ----
        if (warp < 4 && nch > 0) {
            mbar_wait(&bars[4], 0);
            fence_after_sync();
            float* dst = jb.dW + (size_t)(o0 + tid) * jb.C + c0;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) atomicAdd(dst + c * 32 + e, jb.alpha * v[e]);
            }
        }
    } else {
        // ---- issuer ----
        const bool el = elect_one();
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1;
            mbar_wait(&bars[st], (ch >> 1) & 1u);
            fence_after_sync();
            const uint32_t ah = dlo_mn(smem_u32(sm) + st * WU_STAGE, WU_CB), al = ah + (WU_SUB >> 4), bh = ah + 2 * (WU_SUB >> 4), bl = ah + 3 * (WU_SUB >> 4);
            if (el) {
#pragma unroll
                for (int ks = 0; ks < WU_CH / 16; ++ks) {
                    mma_lo(tb, ah + ks * 128, bh + ks * 128, IDESC_MN_MN_128, (ch > 0 || ks > 0) ? 1u : 0u);
                    mma_lo(tb, ah + ks * 128, bl + ks * 128, IDESC_MN_MN_128, 1u);
                    mma_lo(tb, al + ks * 128, bh + ks * 128, IDESC_MN_MN_128, 1u);
                }
                mma_commit(&bars[2 + st]);
                if (ch == nch - 1) mma_commit(&bars[4]);
            }
            __syncwarp();
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tb, 128);
}
}  // namespace

cudaError_t savi_launch_wgrad_umma(const WgradArgs& wa, const int* done, int done_target, cudaStream_t st) {
    WUArgs ua;
    ua.wa = wa; ua.done = done; ua.done_target = done_target;
    int64_t total = 0;
    for (int j = 0; j < wa.njobs; ++j) {
        if ((wa.job[j].O & 127) || (wa.job[j].C & 127)) return cudaErrorInvalidValue;
        total += (int64_t)(wa.job[j].O >> 7) * (wa.job[j].C >> 7) * ((wa.job[j].R + WU_CH - 1) / WU_CH);
    }
    // about two CTAs per SM-slot worth of work, at least 8 chunks each
    int cpc = (int)((total + 2 * 148 - 1) / (2 * 148));
    if (cpc < 8) cpc = 8;
    ua.chunks_per_cta = cpc;
    int grid = 0;
    for (int j = 0; j < wa.njobs; ++j) {
        const int chunks = (wa.job[j].R + WU_CH - 1) / WU_CH;
        grid += (wa.job[j].O >> 7) * (wa.job[j].C >> 7) * ((chunks + cpc - 1) / cpc);
    }
    if (grid == 0) return cudaSuccess;
    const int smem = 2 * WU_STAGE + 128;
    cudaError_t e = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(WU_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = done ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, wgrad_umma_kernel, ua);
}
