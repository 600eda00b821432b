// Weight gradients of the tcgen05 path: dW[o][c] += alpha * sum_r dY[r][o] X[r][c] for every weight matrix in one
// launch (r runs over all steps and clips: the "one tall GEMM per weight" layout of savi_layout.h).
// Per CTA: one [128 x 128] output tile and a slice of the rows, in chunks of 32 rows.  One producer thread streams the chunk's
// dY and X row slices (fp32, [32 rows][128 columns] boxes of rank-2 tensor maps: two TMA copies per chunk; per-row bulk
// copies are issue-bound at ~40 ns each) into a 4-deep shared-memory ring, so 96 KB per SM are always in flight; eight
// converter warps split them into bf16 hi / lo and write them as MN-major SWIZZLE_128B operand blocks (the row
// index is the contraction index, so no transposition is needed); one warp issues three tcgen05.mma per k-step
// (hi.hi + hi.lo + lo.hi, M = N = 128) into a TMEM accumulator; the epilogue adds the tile into the flat gradient buffer.
// (The previous version loaded the rows into registers, one chunk and one DRAM round trip at a time: 82 us at 37 % of HBM peak.)
#include "savi_umma.cuh"
#include "savi_dev.cuh"
#include "savi_args.h"
#include <cuda.h>
#include <cudaTypedefs.h>

using namespace umma;
typedef __nv_bfloat16 bf16;

namespace {
constexpr int WU_LOADERS = 256, WU_THREADS = 320;     // warps 0-7 convert, warp 8 issues the MMAs, warp 9 streams the rows in
constexpr int WU_CH = 32;                             // rows per chunk
constexpr int WU_NS = 4;                              // ring depth
constexpr int WU_RING = 2 * WU_CH * 512;              // one ring stage: dY rows [32][128] fp32 | X rows [32][128] fp32
constexpr int WU_CB = WU_CH * 128;                    // bytes of one 64-column block of a chunk
constexpr int WU_SUB = 2 * WU_CB;                     // one operand: [WU_CH rows][2 x 64 columns] bf16
constexpr int WU_STAGE = 4 * WU_SUB;                  // A hi | A lo | B hi | B lo
constexpr int WU_LD = WU_CH * 16 / WU_LOADERS;        // 8-float groups per converter thread and operand
constexpr uint32_t IDESC_MN_MN_128 = idesc_bf16(128, 128, true, true);

struct alignas(64) WUArgs {
    CUtensorMap tm[2 * WGRAD_MAX_JOBS];                   // job j: dY rows at [2 j], X rows at [2 j + 1]; fp32 [R][ld], box [32][128]
    WgradArgs wa; int chunks_per_cta;
    const int* done; int done_target;                     // overlapped launch: CTA counter of the backward clip kernel to wait for
};
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 :: "r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

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

__global__ void __launch_bounds__(WU_THREADS, 1) wgrad_umma_kernel(const __grid_constant__ WUArgs ua) {
    extern __shared__ __align__(1024) unsigned char sm[];
    if ((smem_u32(sm) & 1023u) != 0u) __trap();
    const WgradArgs& wa = ua.wa;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // (provably warp-uniform role index)
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

    unsigned char* ring = sm + 2 * WU_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + WU_NS * WU_RING);
    // operands written[2], operands consumed[2], accumulator complete, ring stage landed[NS], ring stage read[NS]
    enum { BOF = 0, BOE = 2, BDONE = 4, BRF = 5, BRE = 5 + WU_NS };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * WU_NS);
    if (tid == 0) {
        mbar_init(&bars[BOF], 8); mbar_init(&bars[BOF + 1], 8); mbar_init(&bars[BOE], 1); mbar_init(&bars[BOE + 1], 1); mbar_init(&bars[BDONE], 1);
        for (int i = 0; i < WU_NS; ++i) { mbar_init(&bars[BRF + i], 1); mbar_init(&bars[BRE + i], 8); }
        mbar_init_fence();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 128);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = *tmem_slot;

    if (warp < 8) {
        // ---- converters: fp32 rows (ring) -> bf16 hi / lo MN-major blocks [32 rows][2 x 64 columns] ----
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1, rs = ch % WU_NS;
            unsigned char* base = sm + st * WU_STAGE;
            const unsigned char* src = ring + rs * WU_RING;
            float4 pa[WU_LD][2], pb[WU_LD][2];
            mbar_wait(&bars[BRF + rs], (ch / WU_NS) & 1u);                 // the chunk's rows landed
#pragma unroll
            for (int i = 0; i < WU_LD; ++i) {
                const int u = tid + i * WU_LOADERS, r = u >> 4, cg = u & 15;
                // (rows past R are zero-filled by the tensor copy; a slice ends inside a chunk only at R)
                const float4* ya = reinterpret_cast<const float4*>(src + r * 512 + cg * 32);
                const float4* xa = reinterpret_cast<const float4*>(src + WU_CH * 512 + r * 512 + cg * 32);
                // a thread's two 16-byte halves in the order that puts the 8 lanes of a shared-memory phase on 8 distinct slots
                // (rows are 32 bytes apart per lane: reading half 0 everywhere is a 2- to 8-way bank conflict)
                const int h = (cg >> 2) & 1;
                const float4 y0 = ya[h], y1 = ya[h ^ 1], x0 = xa[h], x1 = xa[h ^ 1];
                pa[i][0] = h ? y1 : y0; pa[i][1] = h ? y0 : y1;
                pb[i][0] = h ? x1 : x0; pb[i][1] = h ? x0 : x1;
            }
            mbar_wait(&bars[BOE + st], ((ch >> 1) & 1u) ^ 1u);             // the MMAs that read this operand stage two chunks ago are done
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
            if (lane == 0) {
                // The ring stage is released only HERE, behind the stores that consumed its values: an arrive that merely follows
                // the shared-memory loads in program order may be scheduled before their data has returned (it was, once the
                // role branches became provably uniform), and the next tensor copy then overwrites what is still being read.
                mbar_arrive(&bars[BRE + rs]);
                mbar_arrive(&bars[BOF + st]);
            }
        }
        // ---- epilogue (warps 0-3: thread = output row o) ----
        if (warp < 4 && nch > 0) {
            mbar_wait(&bars[BDONE], 0);
            fence_after_sync();
            float* dst = jb.dW + (size_t)(o0 + tid) * jb.C + c0;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                float v[32];
                tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; e += 4)                           // 16-byte vector reductions: a quarter of the L2 transactions
                    atomicAdd(reinterpret_cast<float4*>(dst + c * 32 + e), make_float4(jb.alpha * v[e], jb.alpha * v[e + 1], jb.alpha * v[e + 2], jb.alpha * v[e + 3]));
            }
        }
    } else if (warp == 8) {
        // ---- issuer ----
        const bool el = elect_one();
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1;
            mbar_wait(&bars[BOF + st], (ch >> 1) & 1u);
            fence_after_sync();
            const uint32_t ah = dlo_mn(smem_u32(sm) + st * WU_STAGE, WU_CB), al = ah + (WU_SUB >> 4), bh = ah + 2 * (WU_SUB >> 4), bl = ah + 3 * (WU_SUB >> 4);
            if (el) {
#pragma unroll
                for (int ks = 0; ks < WU_CH / 16; ++ks) {
                    mma_lo(tb, ah + ks * 128, bh + ks * 128, IDESC_MN_MN_128, (ch > 0 || ks > 0) ? 1u : 0u);
                    mma_lo(tb, ah + ks * 128, bl + ks * 128, IDESC_MN_MN_128, 1u);
                    mma_lo(tb, al + ks * 128, bh + ks * 128, IDESC_MN_MN_128, 1u);
                }
                mma_commit(&bars[BOE + st]);
                if (ch == nch - 1) mma_commit(&bars[BDONE]);
            }
            __syncwarp();
        }
    } else {
        // ---- producer (one thread): two tensor copies per chunk ----
        // the records were written by the clip kernel with generic-proxy stores and are read here through the async proxy
        asm volatile("fence.proxy.async.global;\n" ::: "memory");
        if (lane == 0) {
            for (int ch = 0; ch < nch; ++ch) {
                const int rs = ch % WU_NS;
                const int r0 = r_begin + ch * WU_CH;
                mbar_wait(&bars[BRE + rs], ((ch / WU_NS) & 1u) ^ 1u);      // the converters have read the stage's previous chunk
                mbar_expect_tx(&bars[BRF + rs], (uint32_t)WU_RING);
                tma_load_2d(ring + rs * WU_RING, &ua.tm[2 * job], o0, r0, &bars[BRF + rs]);
                tma_load_2d(ring + rs * WU_RING + WU_CH * 512, &ua.tm[2 * job + 1], c0, r0, &bars[BRF + rs]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tb, 128);
}
}  // namespace

// fp32 [R][ld] as a rank-2 tensor map with box [32 rows][128 columns], no swizzle (rows past R read as zeros)
static cudaError_t make_rows_map(CUtensorMap* tm, const float* base, int R, int ld) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)R};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {128, (cuuint32_t)WU_CH}, estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t savi_launch_wgrad_umma(const WgradArgs& wa, const int* done, int done_target, cudaStream_t st) {
    WUArgs ua;
    for (int j = 0; j < wa.njobs; ++j) {
        cudaError_t me = make_rows_map(&ua.tm[2 * j], wa.job[j].dY, wa.job[j].R, wa.job[j].ldy);
        if (me == cudaSuccess) me = make_rows_map(&ua.tm[2 * j + 1], wa.job[j].X, wa.job[j].R, wa.job[j].ldx);
        if (me != cudaSuccess) return me;
    }
    ua.wa = wa; ua.done = done; ua.done_target = done_target;
    int64_t total = 0;
    for (int j = 0; j < wa.njobs; ++j) {
        if ((wa.job[j].O & 127) || (wa.job[j].C & 127)) return cudaErrorInvalidValue;
        total += (int64_t)(wa.job[j].O >> 7) * (wa.job[j].C >> 7) * ((wa.job[j].R + WU_CH - 1) / WU_CH);
    }
    // two waves of one CTA per SM, at least 8 chunks each
    int cpc = (int)((total + 2 * 148 - 1) / (2 * 148));
    if (cpc < 8) cpc = 8;
    ua.chunks_per_cta = cpc;
    int grid = 0;
    for (int j = 0; j < wa.njobs; ++j) {
        const int chunks = (wa.job[j].R + WU_CH - 1) / WU_CH;
        grid += (wa.job[j].O >> 7) * (wa.job[j].C >> 7) * ((chunks + cpc - 1) / cpc);
    }
    if (grid == 0) return cudaSuccess;
    const int smem = 2 * WU_STAGE + WU_NS * WU_RING + 256;
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
