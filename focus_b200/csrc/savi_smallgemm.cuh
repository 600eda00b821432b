// A handful of small dense fp32 products C = alpha * A . B in ONE launch (strided operands, so transposes are free):
// the folded weights of the tcgen05 path (pack time) and the chain rule back through them (after the weight gradients).
// 32 x 32 output tile per CTA, 256 threads, shared-memory staging with coalesced loads along whichever index is
// contiguous.  M, N are multiples of 32 and Kd of 128 (the tcgen05 path has D = Ds = 128).
#pragma once
#include <cuda_runtime.h>

struct SmallGemm {
    const float* A; const float* B; float* C;
    int M, N, Kd;            // C [M][N] row-major (ldc = N), contraction length Kd
    int sai, sak;            // A(i, k) = A[i * sai + k * sak]
    int sbk, sbj;            // B(k, j) = B[k * sbk + j * sbj]
    float alpha;
    int accumulate;          // 0: C = alpha A B (plain store); 1: C += alpha A B (atomic: a split contraction into a zeroed C)
};
constexpr int SMALLGEMM_MAX = 6;
struct SmallGemmArgs { SmallGemm g[SMALLGEMM_MAX]; int count; };

constexpr int SG_KC = 128;       // contraction chunk: 32 independent loads in flight per thread, 1-3 chunks per product
static __global__ void __launch_bounds__(256) small_gemm_kernel(const __grid_constant__ SmallGemmArgs ga) {
    __shared__ float As[32][SG_KC + 1], Bs[SG_KC][33];
    int job = 0, tile = blockIdx.x;
    for (; job < ga.count; ++job) {
        const int tiles = (ga.g[job].M >> 5) * (ga.g[job].N >> 5);
        if (tile < tiles) break;
        tile -= tiles;
    }
    if (job >= ga.count) return;
    const SmallGemm& g = ga.g[job];
    const int tn = g.N >> 5, i0 = (tile / tn) * 32, j0 = (tile % tn) * 32;
    const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < g.Kd; k0 += SG_KC) {
        float ra[16], rb[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {                       // all loads of the chunk are issued before any is used
            if (g.sak == 1) ra[m] = g.A[(size_t)(i0 + (t >> 7) + 2 * m) * g.sai + (k0 + (t & 127))];
            else            ra[m] = g.A[(size_t)(i0 + tx) * g.sai + (size_t)(k0 + ty + 8 * m) * g.sak];
            if (g.sbj == 1) rb[m] = g.B[(size_t)(k0 + ty + 8 * m) * g.sbk + (j0 + tx)];
            else            rb[m] = g.B[(size_t)(k0 + (t & 127)) * g.sbk + (size_t)(j0 + (t >> 7) + 2 * m) * g.sbj];
        }
        __syncthreads();                                     // the previous chunk's tiles have been consumed
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (g.sak == 1) As[(t >> 7) + 2 * m][t & 127] = ra[m]; else As[tx][ty + 8 * m] = ra[m];
            if (g.sbj == 1) Bs[ty + 8 * m][tx] = rb[m]; else Bs[t & 127][(t >> 7) + 2 * m] = rb[m];
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < SG_KC; ++k) {
            const float b = Bs[k][tx];
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m] = fmaf(As[ty + 8 * m][k], b, acc[m]);
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float* dst = g.C + (size_t)(i0 + ty + 8 * m) * g.N + j0 + tx;
        if (g.accumulate) atomicAdd(dst, g.alpha * acc[m]); else *dst = g.alpha * acc[m];
    }
}

static inline cudaError_t launch_small_gemms(const SmallGemmArgs& ga, cudaStream_t st) {
    int tiles = 0;
    for (int j = 0; j < ga.count; ++j) {
        if (((ga.g[j].M | ga.g[j].N) & 31) || ga.g[j].Kd % SG_KC) return cudaErrorInvalidValue;
        tiles += (ga.g[j].M >> 5) * (ga.g[j].N >> 5);
    }
    small_gemm_kernel<<<tiles, 256, 0, st>>>(ga);
    return cudaGetLastError();
}
