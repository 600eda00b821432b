// How fast can the N1 tile gather ([BT][C][HW] fp32 map, 128 consecutive pixels x C channel rows per tile) be read at all?
// Times (a) a linear float4 stream over the same bytes, (b) the gather with one token per thread and C loads in flight,
// (c) the same with the next tile prefetched into L2.   nvcc -arch=sm_100a -O3 -o tools/build/n1_loadpattern tools/n1_loadpattern.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int C = 128, HW = 1024, BT = 384;
__global__ void stream_k(const float4* __restrict__ p, float* out, size_t n4) {
    float s = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldg(p + i); s += v.x + v.y + v.z + v.w; }
    if (s == 1234.5f) out[0] = s;
}
template <bool PF>
__global__ void gather_k(const float* __restrict__ emb, float* out, long long ntile) {
    const int lt = threadIdx.x & 127, grp = threadIdx.x >> 7, G = blockDim.x >> 7;
    const long long stride = (long long)gridDim.x * G;
    float acc = 0.f;
    for (long long tile = (long long)blockIdx.x * G + grp; tile < ntile; tile += stride) {
        const long long g = tile * 128 + lt, bt = g / HW; const int pix = (int)(g - bt * HW);
        const float* src = emb + (size_t)bt * C * HW + pix;
        float x[C];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)c * HW);
        if (PF && tile + stride < ntile) {
            const long long g2 = (tile + stride) * 128, bt2 = g2 / HW; const int pix2 = (int)(g2 - bt2 * HW);
            const float* s2 = emb + ((size_t)bt2 * C + lt) * HW + pix2;
#pragma unroll
            for (int j = 0; j < 4; ++j) asm volatile("prefetch.global.L2 [%0];" :: "l"(s2 + j * 32));
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc += x[c];
    }
    if (acc == 1234.5f) out[0] = acc;
}
int main() {
    const size_t n = (size_t)BT * C * HW;
    float *emb, *out, *flush; cudaMalloc(&emb, n * 4); cudaMalloc(&out, 4096); cudaMalloc(&flush, 256 << 20);
    cudaMemset(emb, 0, n * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const long long ntile = (long long)BT * HW / 128;
    auto run = [&](const char* name, auto fn) {
        float best = 1e9f;
        for (int i = 0; i < 8; ++i) { cudaMemsetAsync(flush, i, 256 << 20); cudaEventRecord(a); fn(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 3 && ms < best) best = ms; }
        printf("%-44s %7.1f us  %7.0f GB/s  (%s)\n", name, best * 1e3, n * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    run("linear float4 stream, 148x8 x 256", [&] { stream_k<<<148 * 8, 256>>>((const float4*)emb, out, n / 4); });
    run("gather, 296 x 128", [&] { gather_k<false><<<296, 128>>>(emb, out, ntile); });
    run("gather, 148 x 384", [&] { gather_k<false><<<148, 384>>>(emb, out, ntile); });
    run("gather + L2 prefetch, 148 x 384", [&] { gather_k<true><<<148, 384>>>(emb, out, ntile); });
    run("gather, 592 x 128", [&] { gather_k<false><<<592, 128>>>(emb, out, ntile); });
    run("gather, 3072 x 128 (one tile per CTA)", [&] { gather_k<false><<<3072, 128>>>(emb, out, ntile); });
    return 0;
}
