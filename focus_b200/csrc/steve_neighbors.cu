// Components next to the slot-attention path (C ABI: include/focus_steve.h).
//
//   N2  attention-map consumer: reference slowfast/models/STEVE/steve.py:314-319, 349-355.  HBM-bound: the overlay
//       [B,T,K,C,H,W] is K*C*(H*W/N) times larger than the attention maps it is made from (e.g. 24 * 3 * 4 at 128 px); the
//       kernel writes it exactly once with 16-byte stores and reads `attn` / `video` through L1/L2 (every attention value
//       is reused by C * (H/He) * (W/We) output pixels).  Algorithmic bytes = overlay + up + video + attn.
//   N4  FG-ARI contingency tables: reference slowfast/utils/metrics.py:40-83.  Integer work: per pixel one argmax over the
//       predicted segments and one shared-memory atomic per ground-truth segment the pixel belongs to; the reference builds
//       an [N0, N1, D] boolean tensor per sample on the CPU instead.  Algorithmic bytes = true_mask + pred_mask, read once.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include "focus_savi.h"
#include "focus_steve.h"

int savi_set_error(int code, const char* fmt, ...);      // savi_api.cu: fills the thread-local savi_last_error() message

// ---------------------------------------------------------------------------------------------------------------
// N2
// ---------------------------------------------------------------------------------------------------------------
template <typename AT> __device__ __forceinline__ float attn_ld(const AT* p);
template <> __device__ __forceinline__ float attn_ld<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float attn_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// thread = 4 consecutive pixels of one image row; loops over the K slots (outer) and C channels (inner)
template <typename AT, int CMAX>
__global__ void __launch_bounds__(256) overlay_kernel(const AT* __restrict__ attn, const float* __restrict__ video,
                                                      float* __restrict__ overlay, float* __restrict__ up,
                                                      int K, int C, int H, int W, int He, int We) {
    const int64_t bt = blockIdx.y;
    const int w4n = W >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * w4n) return;
    const int h = idx / w4n, w = (idx - h * w4n) * 4;
    const int rh = H / He, rw = W / We;
    const int64_t HW = (int64_t)H * W;
    // attention cell of each of the four pixels (identical when rw >= 4)
    int cell[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cell[j] = (h / rh) * We + (w + j) / rw;
    float4 v[CMAX];
    const float* vp = video + bt * C * HW + (int64_t)h * W + w;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) if (c < C) v[c] = *reinterpret_cast<const float4*>(vp + c * HW);
    const AT* ap = attn + bt * (int64_t)He * We * K;
    float* op = overlay ? overlay + bt * K * C * HW + (int64_t)h * W + w : nullptr;
    float* upp = up ? up + bt * K * HW + (int64_t)h * W + w : nullptr;
    for (int k = 0; k < K; ++k) {
        float4 a;
        a.x = attn_ld<AT>(ap + (int64_t)cell[0] * K + k);
        a.y = (cell[1] == cell[0]) ? a.x : attn_ld<AT>(ap + (int64_t)cell[1] * K + k);
        a.z = (cell[2] == cell[1]) ? a.y : attn_ld<AT>(ap + (int64_t)cell[2] * K + k);
        a.w = (cell[3] == cell[2]) ? a.z : attn_ld<AT>(ap + (int64_t)cell[3] * K + k);
        if (upp) __stcs(reinterpret_cast<float4*>(upp + k * HW), a);
        if (op) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                if (c < C) {
                    // video * attn + (1 - attn): the reference's operation order (mul, rsub, add), fp32
                    float4 o;
                    // (explicitly rounded product and sum: an FMA contraction would differ from the reference by one ulp)
                    o.x = __fadd_rn(__fmul_rn(v[c].x, a.x), 1.0f - a.x); o.y = __fadd_rn(__fmul_rn(v[c].y, a.y), 1.0f - a.y);
                    o.z = __fadd_rn(__fmul_rn(v[c].z, a.z), 1.0f - a.z); o.w = __fadd_rn(__fmul_rn(v[c].w, a.w), 1.0f - a.w);
                    __stcs(reinterpret_cast<float4*>(op + ((int64_t)k * C + c) * HW), o);      // streaming: written once, never re-read here
                }
            }
        }
    }
}

extern "C" int steve_attention_overlay(const void* attn, int attn_dtype, const float* video, float* overlay, float* up,
                                       int64_t BT, int K, int C, int H, int W, int He, int We, void* stream) {
    if (!attn || !video || (!overlay && !up)) return savi_set_error(SAVI_EINVAL, "steve_attention_overlay: null pointer");
    if (BT < 1 || BT > 65535) return savi_set_error(SAVI_EINVAL, "steve_attention_overlay: B*T = %lld outside [1, 65535]", (long long)BT);
    if (K < 1 || C < 1 || C > 4 || H < 1 || W < 4 || (W & 3)) return savi_set_error(SAVI_EINVAL, "steve_attention_overlay: needs 1 <= C <= 4 and W %% 4 == 0 (K=%d C=%d H=%d W=%d)", K, C, H, W);
    if (He < 1 || We < 1 || H % He || W % We) return savi_set_error(SAVI_EINVAL, "steve_attention_overlay: H, W must be multiples of the attention grid (%d x %d vs %d x %d)", H, W, He, We);
    if (attn_dtype != SAVI_DTYPE_F32 && attn_dtype != SAVI_DTYPE_BF16) return savi_set_error(SAVI_EINVAL, "steve_attention_overlay: unknown attention dtype %d", attn_dtype);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)((H * (W >> 2) + 255) / 256), (unsigned)BT);
    if (attn_dtype == SAVI_DTYPE_F32)
        overlay_kernel<float, 4><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(attn), video, overlay, up, K, C, H, W, He, We);
    else
        overlay_kernel<__nv_bfloat16, 4><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(attn), video, overlay, up, K, C, H, W, He, We);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "overlay_kernel: %s", cudaGetErrorString(e));
    return SAVI_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// N4
// ---------------------------------------------------------------------------------------------------------------
constexpr int ARI_THREADS = 256;
constexpr int ARI_PIX = 8;         // pixels per thread

__global__ void __launch_bounds__(ARI_THREADS) ari_table_kernel(const float* __restrict__ tm, const float* __restrict__ pm,
                                                                int* __restrict__ tables, int N0, int N1, int64_t D) {
    extern __shared__ int stab[];                      // [N0][N1]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < N0 * N1; i += ARI_THREADS) stab[i] = 0;
    __syncthreads();
    const float* t = tm + (int64_t)b * N0 * D;
    const float* p = pm + (int64_t)b * N1 * D;
    const int64_t d0 = (int64_t)blockIdx.x * ARI_THREADS * ARI_PIX + threadIdx.x;
#pragma unroll 2
    for (int q = 0; q < ARI_PIX; ++q) {
        const int64_t d = d0 + (int64_t)q * ARI_THREADS;      // consecutive threads -> consecutive pixels: coalesced in every plane
        if (d >= D) break;
        // argmax over the predicted segments, first maximum wins (torch.argmax; metrics.py:72)
        float best = __ldg(p + d);
        int arg = 0;
        for (int j = 1; j < N1; ++j) {
            const float x = __ldg(p + (int64_t)j * D + d);
            if (x > best) { best = x; arg = j; }
        }
        for (int i = 0; i < N0; ++i) {
            const float x = __ldg(t + (int64_t)i * D + d);
            // mask0.byte() & mask1.byte() (metrics.py:50-55): float -> uint8 truncation, then the low bit against the one-hot 1
            if (((unsigned)(int)x) & 1u) atomicAdd(&stab[i * N1 + arg], 1);
        }
    }
    __syncthreads();
    int* out = tables + (int64_t)b * N0 * N1;
    for (int i = threadIdx.x; i < N0 * N1; i += ARI_THREADS) if (stab[i]) atomicAdd(out + i, stab[i]);
}

extern "C" int steve_ari_tables(const float* true_mask, const float* pred_mask, int32_t* tables,
                                int B, int N0, int N1, int64_t D, void* stream) {
    if (!true_mask || !pred_mask || !tables) return savi_set_error(SAVI_EINVAL, "steve_ari_tables: null pointer");
    if (B < 1 || B > 65535 || N0 < 1 || N0 > 64 || N1 < 1 || N1 > 64 || D < 1)
        return savi_set_error(SAVI_EINVAL, "steve_ari_tables: needs 1 <= B <= 65535, 1 <= N0, N1 <= 64, D >= 1 (B=%d N0=%d N1=%d D=%lld)", B, N0, N1, (long long)D);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(tables, 0, (size_t)B * N0 * N1 * sizeof(int32_t), st);
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int64_t per = (int64_t)ARI_THREADS * ARI_PIX;
    const int64_t gx = (D + per - 1) / per;
    if (gx > 2147483647LL) return savi_set_error(SAVI_EINVAL, "steve_ari_tables: D too large");
    ari_table_kernel<<<dim3((unsigned)gx, (unsigned)B), ARI_THREADS, (size_t)N0 * N1 * sizeof(int), st>>>(true_mask, pred_mask, tables, N0, N1, D);
    e = cudaGetLastError();
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "ari_table_kernel: %s", cudaGetErrorString(e));
    return SAVI_OK;
}
