// A handful of small dense fp32 products C = alpha * A . B in ONE launch (strided operands, so transposes are free):
// the folded weights of the tcgen05 path (pack time) and the chain rule back through them (after the weight gradients).
// 32 x 32 output tile per CTA, 256 threads, shared-memory staging with coalesced loads along whichever index is
// contiguous.  All extents are multiples of 32 (the tcgen05 path has D = Ds = 128).
#pragma once
#include <cuda_runtime.h>

struct SmallGemm {
    const float* A; const float* B; float* C;
    int M, N, Kd;            // C [M][N] row-major (ldc = N), contraction length Kd
    int sai, sak;            // A(i, k) = A[i * sai + k * sak]
    int sbk, sbj;            // B(k, j) = B[k * sbk + j * sbj]
    float alpha;
};
constexpr int SMALLGEMM_MAX = 4;
struct SmallGemmArgs { SmallGemm g[SMALLGEMM_MAX]; int count; };

static __global__ void __launch_bounds__(256) small_gemm_kernel(const __grid_constant__ SmallGemmArgs ga) {
    __shared__ float As[32][33], Bs[32][33];
    int job = 0, tile = blockIdx.x;
    for (; job < ga.count; ++job) {
        const int tiles = (ga.g[job].M >> 5) * (ga.g[job].N >> 5);
        if (tile < tiles) break;
        tile -= tiles;
    }
    if (job >= ga.count) return;
    const SmallGemm& g = ga.g[job];
    const int tn = g.N >> 5, i0 = (tile / tn) * 32, j0 = (tile % tn) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < g.Kd; k0 += 32) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int r = ty + 8 * m;
            if (g.sak == 1) As[r][tx] = g.A[(size_t)(i0 + r) * g.sai + (k0 + tx)];
            else            As[tx][r] = g.A[(size_t)(i0 + tx) * g.sai + (size_t)(k0 + r) * g.sak];
            if (g.sbj == 1) Bs[r][tx] = g.B[(size_t)(k0 + r) * g.sbk + (j0 + tx)];
            else            Bs[tx][r] = g.B[(size_t)(k0 + tx) * g.sbk + (size_t)(j0 + r) * g.sbj];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float b = Bs[k][tx];
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m] = fmaf(As[ty + 8 * m][k], b, acc[m]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) g.C[(size_t)(i0 + ty + 8 * m) * g.N + j0 + tx] = g.alpha * acc[m];
}

static inline cudaError_t launch_small_gemms(const SmallGemmArgs& ga, cudaStream_t st) {
    int tiles = 0;
    for (int j = 0; j < ga.count; ++j) {
        if ((ga.g[j].M | ga.g[j].N | ga.g[j].Kd) & 31) return cudaErrorInvalidValue;
        tiles += (ga.g[j].M >> 5) * (ga.g[j].N >> 5);
    }
    small_gemm_kernel<<<tiles, 256, 0, st>>>(ga);
    return cudaGetLastError();
}
