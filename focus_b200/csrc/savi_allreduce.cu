// The data-parallel exchange step of the path as the library's own kernel: a one-shot all-reduce of the flat
// parameter-gradient buffer over NVLink / NVSwitch peer memory (the reference gets it from DDP: slowfast/models/build.py:79-83).
// The buffer is only ~1.6 MB, so the exchange is pure latency (NCCL: ~33 us per call at 8 GPUs, plus its host-side launch cost);
// here every rank reads all peers' buffers directly (or ONE multimem.ld_reduce through the switch where the allocation has a
// multicast mapping) and writes the scaled sum into its own output: ~11 MB of peer reads per rank at 8 GPUs.
//
// Every rank's gradient buffer lives in symmetric memory (torch.distributed._symmetric_memory: same offset on every rank,
// peer-mapped, with a zero-initialised "signal pad" per rank).  CTA b of a rank only ever talks to CTA b of its peers:
//   barrier (channel 0): peer buffers are complete — each rank's kernel runs behind its own backward in stream order, so a peer
//                        CTA that has arrived implies that peer's buffer is final
//   out[chunk b] = scale * sum_r peer_r[chunk b]        (chunk b: the float4 elements CTA b's threads stride over)
//   barrier (channel 1): every peer has finished reading this rank's chunk b; when the kernel ends no rank reads the buffer any
//                        more and the next step may overwrite it
// A barrier is the signal-pad handshake torch's own symmetric-memory kernels use: put = CAS 0 -> 1 (release, system scope) on the
// peer's pad slot [channel][block][my rank], wait = CAS 1 -> 0 (acquire) on my pad's slot [channel][block][peer]: self-resetting.
#include <cuda_runtime.h>
#include <cstdint>
#include "focus_savi.h"

int savi_set_error(int code, const char* fmt, ...);

namespace {
constexpr int AR_MAX_WORLD = 16;
constexpr int AR_THREADS = 512;

struct ArArgs {
    const float* peer[AR_MAX_WORLD];     // this rank's view of every rank's gradient buffer (peer[rank] = its own)
    uint32_t* pad[AR_MAX_WORLD];         // ... and of every rank's signal pad
    const float* mc;                     // multicast address of the buffer (nullptr: read the peers one by one)
    float* out;
    long long n4;                        // float4 elements
    float scale;
    int rank, world, nblk;
};

__device__ __forceinline__ void put_signal(uint32_t* addr) {
    uint32_t old;
    do {
        asm volatile("atom.release.sys.global.cas.b32 %0, [%1], 0, 1;\n" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 0u);
}
__device__ __forceinline__ void wait_signal(uint32_t* addr) {
    uint32_t old;
    do {
        asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], 1, 0;\n" : "=r"(old) : "l"(addr) : "memory");
    } while (old != 1u);
}
// all ranks' CTA `blk`: thread p handles peer p
__device__ __forceinline__ void peer_barrier(const ArArgs& a, int chan, int blk) {
    __syncthreads();
    const int p = threadIdx.x;
    if (p < a.world && p != a.rank) {
        const size_t slot = ((size_t)chan * a.nblk + blk) * a.world;
        put_signal(a.pad[p] + slot + a.rank);
        wait_signal(a.pad[a.rank] + slot + p);
    }
    __syncthreads();
}
__device__ __forceinline__ float4 ld_peer(const float* p) {      // peer memory changes between steps: never through the read-only path
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_reduce_mc(const float* p) {  // the NVSwitch adds the value at this offset of every rank's buffer
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_peers_kernel(const __grid_constant__ ArArgs a) {
    const int blk = blockIdx.x;
    peer_barrier(a, 0, blk);
    // chunk b = float4 elements {b, b + nblk, ...} x 512 threads; the multicast path keeps four loads per thread in flight, the peer
    // path one element (= `world` loads) at a time
    const long long stride = (long long)a.nblk * AR_THREADS;
    if (a.mc) {
        for (long long i0 = (long long)blk * AR_THREADS + threadIdx.x; i0 < a.n4; i0 += 4 * stride) {
            float4 s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * stride < a.n4) s[u] = ld_reduce_mc(a.mc + 4 * (i0 + u * stride));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * stride < a.n4)
                    reinterpret_cast<float4*>(a.out)[i0 + u * stride] = make_float4(s[u].x * a.scale, s[u].y * a.scale, s[u].z * a.scale, s[u].w * a.scale);
        }
    } else {
        for (long long i = (long long)blk * AR_THREADS + threadIdx.x; i < a.n4; i += stride) {
            float4 v[AR_MAX_WORLD];
#pragma unroll
            for (int r = 0; r < AR_MAX_WORLD; ++r)               // all peers' loads in flight together; summed in rank order on every rank
                if (r < a.world) v[r] = ld_peer(a.peer[r] + 4 * i);
            float4 s = v[0];
#pragma unroll
            for (int r = 1; r < AR_MAX_WORLD; ++r)
                if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
            reinterpret_cast<float4*>(a.out)[i] = make_float4(s.x * a.scale, s.y * a.scale, s.z * a.scale, s.w * a.scale);
        }
    }
    peer_barrier(a, 1, blk);
}
}  // namespace

extern "C" int savi_allreduce_peers(const void* const* peer_bufs, void* const* signal_pads, const void* multicast, int rank, int world,
                                    void* out, int64_t n_floats, float scale, int64_t signal_pad_bytes, void* stream) {
    if (!peer_bufs || !signal_pads || !out) return savi_set_error(SAVI_EINVAL, "savi_allreduce_peers: null pointer");
    if (world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world)
        return savi_set_error(SAVI_EINVAL, "savi_allreduce_peers: rank %d / world %d (world <= %d)", rank, world, AR_MAX_WORLD);
    if (n_floats < 0 || (n_floats & 3)) return savi_set_error(SAVI_EINVAL, "savi_allreduce_peers: n_floats must be a multiple of 4");
    if (n_floats == 0) return SAVI_OK;
    ArArgs a;
    for (int r = 0; r < AR_MAX_WORLD; ++r) {
        a.peer[r] = r < world ? reinterpret_cast<const float*>(peer_bufs[r]) : nullptr;
        a.pad[r] = r < world ? reinterpret_cast<uint32_t*>(signal_pads[r]) : nullptr;
        if (r < world && (!a.peer[r] || !a.pad[r])) return savi_set_error(SAVI_EINVAL, "savi_allreduce_peers: null peer pointer for rank %d", r);
    }
    a.mc = reinterpret_cast<const float*>(multicast);
    a.out = reinterpret_cast<float*>(out);
    a.n4 = n_floats / 4; a.scale = scale; a.rank = rank; a.world = world;
    // one CTA per 2 K float4 elements (4 per thread), at most what the signal pad has slots for (2 channels x nblk x world words) and 128
    long long nblk = (a.n4 + 4 * AR_THREADS - 1) / (4 * AR_THREADS);
    const long long cap = signal_pad_bytes / 4 / (2 * world);
    if (nblk > cap) nblk = cap;
    if (nblk > 128) nblk = 128;
    if (nblk < 1) return savi_set_error(SAVI_EINVAL, "savi_allreduce_peers: signal pad of %lld bytes is too small", (long long)signal_pad_bytes);
    a.nblk = (int)nblk;
    allreduce_peers_kernel<<<(unsigned)nblk, AR_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return savi_set_error(SAVI_ECUDA, "allreduce_peers_kernel: %s", cudaGetErrorString(e));
    return SAVI_OK;
}
