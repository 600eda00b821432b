// Microbenchmark (run on a B200): how long does ONE issuing thread take per tcgen05.mma of the shapes the clip kernels use?
//   M = 128, K = 16 (bf16), N in {32, 64, 128, 256}; A K-major, B MN-major, SWIZZLE_128B; 512 MMAs back to back into
//   (a) ONE accumulator (dependent chain, as the per-tile products are) and (b) four alternating accumulators; timed with
//   clock64 from the first issue to the arrival of the final tcgen05.commit.  One CTA, so no contention.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I focus_b200/csrc tools/umma_rate.cu -o tools/build/umma_rate
#include <cstdio>
#include <cstdlib>
#include "savi_umma.cuh"
using namespace umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int N, int M>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int nacc, int reps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid * 16; i < 98304; i += 128 * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tslot;
    constexpr uint32_t IDESC = idesc_bf16(M, N, false, true);
    if (warp == 1) {
        const bool el = elect_one();
        const uint32_t a = dlo_k(smem_u32(sm)), b = dlo_mn(smem_u32(sm) + 16384, 16384);
        long long t0 = clock64();
        if (el) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) mma_lo(tb + (uint32_t)((r % nacc) * (N > 128 ? 256 : 128)), a + k4 * 2, b + k4 * 128, IDESC, 1u);
            }
            mma_commit(&bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (el) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int M = 128> void run(long long* d, int nacc) {
    const int reps = 128;
    CK(cudaFuncSetAttribute(rate_kernel<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 99328));
    long long h[2];
    for (int it = 0; it < 2; ++it) {
        rate_kernel<N, M><<<1, 128, 99328>>>(d, nacc, reps);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("M=%3d N=%3d accumulators=%d: %d MMAs  issue %6.1f cyc/MMA   complete %6.1f cyc/MMA   (tensor floor 128*N/256 = %d)\n", M, N, nacc, reps * 4,
           (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), 128 * N / 256);
}
int main() {
    long long* d; CK(cudaMalloc(&d, 64));
    for (int nacc = 1; nacc <= 2; ++nacc) { run<32>(d, nacc); run<64>(d, nacc); run<128>(d, nacc); run<256>(d, nacc); }
    run<32, 64>(d, 1); run<64, 64>(d, 1); run<128, 64>(d, 1);
    return 0;
}
