// Bring-up probe for the tcgen05 primitives in focus_b200/csrc/savi_umma.cuh (run on a B200):
//   T1  K-major A [128x64] . K-major B [32x64]^T  (SWIZZLE_128B, 4 k-steps)       -> TMEM -> tcgen05.ld
//   T2  MN-major A (token tile [128 tok][128 D] used as D x tok) . K-major B [32 slots x 128 tok]
//   T3  operands delivered by cp.async.bulk from a pre-swizzled global image (mbarrier complete_tx)
//   T4  cluster of 2: DSMEM stores + barrier.cluster, multicast bulk copy
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I focus_b200/csrc tools/umma_probe.cu -o gpurun_out/umma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "savi_umma.cuh"

using namespace umma;
typedef __nv_bfloat16 bf16;

__host__ __device__ inline float aval(int r, int c) { return (float)((r * 3 + c) % 7 - 3); }
__host__ __device__ inline float bval(int r, int c) { return (float)((r + 2 * c) % 5 - 2); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// mode 0: T1, mode 1: T2, mode 2: T3 (T1 with bulk-copied operands from gA / gB images)
// mode 3: T5 K-major A . MN-major B [64 k][64 n] (N = 64); mode 4: T6 same operands, N = 32 (left half of B's rows)
// mode 5: T7 MN-major A (token tile) . MN-major B [128 tok][64] (N = 64)
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, float* out, const unsigned char* gA, const unsigned char* gB) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                 // up to 32 KB
    unsigned char* sB = smem + 32768;         // up to 16 KB
    __shared__ uint64_t bar_mma, bar_tx;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_tx, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_base_s;

    if (mode == 0) {
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<bf16*>(sA + sw128_off(r, c)) = __float2bfloat16(aval(r, c)); }
        for (int i = tid; i < 32 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<bf16*>(sB + sw128_off(r, c)) = __float2bfloat16(bval(r, c)); }
        fence_async_smem();
        __syncthreads();
    } else if (mode == 1) {
        // X[t][d], t < 128 tokens, d < 128: block (d / 64) of [128 rows = tokens][64]; value aval(d, t)
        for (int i = tid; i < 128 * 128; i += 128) { int t = i / 128, d = i % 128; *reinterpret_cast<bf16*>(sA + (d >> 6) * 16384 + sw128_off(t, d & 63)) = __float2bfloat16(aval(d, t)); }
        // Wt[s][t], s < 32, t < 128: block (t / 64) of [32 rows][64]; value bval(s, t)
        for (int i = tid; i < 32 * 128; i += 128) { int s = i / 128, t = i % 128; *reinterpret_cast<bf16*>(sB + (t >> 6) * 4096 + sw128_off(s, t & 63)) = __float2bfloat16(bval(s, t)); }
        fence_async_smem();
        __syncthreads();
    } else if (mode == 3 || mode == 4) {
        for (int i = tid; i < 128 * 64; i += 128) { int r = i / 64, c = i % 64; *reinterpret_cast<bf16*>(sA + sw128_off(r, c)) = __float2bfloat16(aval(r, c)); }
        // B[k][n], k < 64 rows, n < 64: one [64 rows][64] block, value bval(n, k)
        for (int i = tid; i < 64 * 64; i += 128) { int k = i / 64, n = i % 64; *reinterpret_cast<bf16*>(sB + sw128_off(k, n)) = __float2bfloat16(bval(n, k)); }
        fence_async_smem();
        __syncthreads();
    } else if (mode == 5) {
        for (int i = tid; i < 128 * 128; i += 128) { int t = i / 128, d = i % 128; *reinterpret_cast<bf16*>(sA + (d >> 6) * 16384 + sw128_off(t, d & 63)) = __float2bfloat16(aval(d, t)); }
        // B[t][n], t < 128 token rows, n < 64: value bval(n, t)
        for (int i = tid; i < 128 * 64; i += 128) { int t = i / 64, n = i % 64; *reinterpret_cast<bf16*>(sB + sw128_off(t, n)) = __float2bfloat16(bval(n, t)); }
        fence_async_smem();
        __syncthreads();
    } else {
        if (tid == 0) {
            mbar_expect_tx(&bar_tx, 16384 + 4096);
            bulk_g2s(sA, gA, 16384, &bar_tx);
            bulk_g2s(sB, gB, 4096, &bar_tx);
        }
        mbar_wait(&bar_tx, 0);
    }

    if (tid == 0) {
        fence_after_sync();
        if (mode == 1) {
            const uint32_t id = idesc_bf16(128, 32, true, false);
            for (int kt = 0; kt < 8; ++kt)       // 16 tokens per step
                mma_ss(tb, desc_mnmajor(smem_u32(sA) + kt * 2048, 16384), desc_kmajor(smem_u32(sB) + (kt >> 2) * 4096 + (kt & 3) * 32), id, kt > 0);
        } else if (mode == 3 || mode == 4) {
            const uint32_t id = idesc_bf16(128, mode == 3 ? 64 : 32, false, true);
            for (int k = 0; k < 4; ++k)
                mma_ss(tb, desc_kmajor(smem_u32(sA) + k * 32), desc_mnmajor(smem_u32(sB) + k * 2048, 8192), id, k > 0);
        } else if (mode == 5) {
            const uint32_t id = idesc_bf16(128, 64, true, true);
            for (int kt = 0; kt < 8; ++kt)
                mma_ss(tb, desc_mnmajor(smem_u32(sA) + kt * 2048, 16384), desc_mnmajor(smem_u32(sB) + kt * 2048, 16384), id, kt > 0);
        } else {
            const uint32_t id = idesc_bf16(128, 32, false, false);
            for (int k = 0; k < 4; ++k)
                mma_ss(tb, desc_kmajor(smem_u32(sA) + k * 32), desc_kmajor(smem_u32(sB) + k * 32), id, k > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    fence_after_sync();
    float v[32];
    tmem_ld32(tmem_addr(tb, warp, 0), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * 64 + j] = v[j];
    tmem_ld32(tmem_addr(tb, warp, 32), v);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * 64 + 32 + j] = v[j];
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 64);
}

// T4: cluster of 2.  Each CTA writes its rank-tagged values into the peer's buffer (DSMEM), and rank 0 multicasts a bulk copy.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) cluster_kernel(float* out, const float* src) {
    __shared__ __align__(16) float inbox[2][128];
    __shared__ __align__(128) float mc[256];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const uint32_t rank = cluster_rank();
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&bar, 1024);
    cluster_arrive(); cluster_wait();            // both barriers armed before the multicast is issued
    if (rank == 0 && tid == 0) bulk_g2s_mcast(mc, src, 1024, &bar, (uint16_t)3);
    for (uint32_t r = 0; r < 2; ++r) st_cluster_f1(map_to_rank(&inbox[rank][tid], r), 1000.f * rank + tid);
    cluster_arrive(); cluster_wait();
    mbar_wait(&bar, 0);
    out[(rank * 4 + 0) * 128 + tid] = inbox[0][tid];
    out[(rank * 4 + 1) * 128 + tid] = inbox[1][tid];
    out[(rank * 4 + 2) * 128 + tid] = mc[tid];
    out[(rank * 4 + 3) * 128 + tid] = mc[128 + tid];
    cluster_arrive(); cluster_wait();
}

int main() {
    float* d_out; CK(cudaMalloc(&d_out, 128 * 64 * 4));
    std::vector<float> h(128 * 64);
    std::vector<unsigned char> imgA(16384), imgB(4096);
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 64; ++c) { bf16 v = __float2bfloat16(aval(r, c)); memcpy(&imgA[sw128_off(r, c)], &v, 2); }
    for (int r = 0; r < 32; ++r) for (int c = 0; c < 64; ++c) { bf16 v = __float2bfloat16(bval(r, c)); memcpy(&imgB[sw128_off(r, c)], &v, 2); }
    unsigned char *gA, *gB; CK(cudaMalloc(&gA, 16384)); CK(cudaMalloc(&gB, 4096));
    CK(cudaMemcpy(gA, imgA.data(), 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(gB, imgB.data(), 4096, cudaMemcpyHostToDevice));
    const int smem = 32768 + 16384 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int mode = 0; mode < 6; ++mode) {
        CK(cudaMemset(d_out, 0, 128 * 64 * 4));
        probe_kernel<<<1, 128, smem>>>(mode, d_out, gA, gB);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        double err = 0; int bad = 0;
        const int ncol = (mode == 3 || mode == 5) ? 64 : 32;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < ncol; ++n) {
            double ref = 0;
            if (mode == 1 || mode == 5) { for (int t = 0; t < 128; ++t) ref += (double)aval(m, t) * bval(n, t); }
            else { for (int k = 0; k < 64; ++k) ref += (double)aval(m, k) * bval(n, k); }
            double e = fabs(ref - h[m * 64 + n]); if (e > err) err = e; if (e > 1e-3) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, h[m * 64 + n], ref); ++bad; }
        }
        const char* names[6] = {"T1 K.K", "T2 MN.K", "T3 bulk", "T5 K.MN N=64", "T6 K.MN N=32", "T7 MN.MN N=64"};
        printf("%s max abs err %.3g  mismatches %d  -> %s\n", names[mode], err, bad, bad ? "FAIL" : "PASS");
    }
    // T4
    float *d_src, *d_o4; CK(cudaMalloc(&d_src, 1024)); CK(cudaMalloc(&d_o4, 2 * 4 * 128 * 4));
    std::vector<float> hs(256); for (int i = 0; i < 256; ++i) hs[i] = 7.f + i;
    CK(cudaMemcpy(d_src, hs.data(), 1024, cudaMemcpyHostToDevice));
    cluster_kernel<<<2, 128>>>(d_o4, d_src);
    CK(cudaDeviceSynchronize());
    std::vector<float> h4(2 * 4 * 128); CK(cudaMemcpy(h4.data(), d_o4, h4.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r = 0; r < 2; ++r) for (int t = 0; t < 128; ++t) {
        if (h4[(r * 4 + 0) * 128 + t] != (float)t) ++bad;
        if (h4[(r * 4 + 1) * 128 + t] != 1000.f + t) ++bad;
        if (h4[(r * 4 + 2) * 128 + t] != 7.f + t) ++bad;
        if (h4[(r * 4 + 3) * 128 + t] != 7.f + 128 + t) ++bad;
    }
    printf("T4 cluster DSMEM + multicast bulk: mismatches %d -> %s\n", bad, bad ? "FAIL" : "PASS");
    return 0;
}
