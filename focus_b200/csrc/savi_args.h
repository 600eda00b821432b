// Kernel argument blocks (passed by value as __grid_constant__) and the
// launcher prototypes shared by savi_api.cu / savi_fwd.cu / savi_bwd.cu.
#pragma once
#include <cuda_runtime.h>
#include "savi_layout.h"

struct FwdArgs {
    Dims d;
    ParamOff po;
    SavedLayout sl;
    FwdWsLayout wl;
    const float* packed;
    const float* noise;
    float* slots_out;
    void* attn_out;
    unsigned char* saved;
    float* ws;
    int stages;           // token-tile pipeline depth of the tensor-core path
    int op_bytes;         // bytes of ONE shared-memory operand buffer (two are carved before the arena; 0 = none)
    int op_width;         // the common contraction width those buffers serve (max(D, Ds)); 0 = pre-staging off
    long long* dbg;       // optional per-phase cycle counters (CTA 0), see savi_debug_set_phase_buffer
    const float* drop;    // predictor dropout masks (training mode, p > 0), nullptr otherwise: savi_dropout_layout()
    int TN;               // tokens per shared-memory tile
    int arena_floats;     // floats of dynamic shared memory usable as the linear-layer staging arena
    int smem_bytes;
};

struct BwdArgs {
    Dims d;
    ParamOff po;
    SavedLayout sl;
    BwdWsLayout wl;
    const float* packed;
    const float* noise;
    const unsigned char* saved;
    const float* grad_slots;
    const void* grad_attn;      // nullable
    float* grad_params;         // flat, zeroed before the clip kernel
    float* grad_noise;          // nullable
    float* ws;
    long long* dbg;
    const float* drop;          // predictor dropout masks of the forward (nullable)
    int stages;
    int op_bytes;
    int op_width;
    int TN;
    int arena_floats;
    int smem_bytes;
};

// Predictor dropout masks (reference transformer.py:12-13, 44, 48, 68): ONE flat fp32 buffer holding, for every predictor block
// evaluation f = j * (T - 1) + t and clip b, the multiplicative masks (0 or 1 / (1 - p)) of the three dropout sites:
//   att [Sp][B][H][K][K]   on the attention probabilities (attn_dropout)
//   out [Sp][B][K][Ds]     on the proj_o output (output_dropout)
//   ffn [Sp][B][K][Ds]     on the ffn.2 output (the nn.Dropout closing the FFN)
struct DropLayout { int64_t att, out, ffn, total; };     // float offsets
static SAVI_HD DropLayout savi_dropout_layout(const Dims& d) {
    DropLayout L;
    const int64_t n = (int64_t)d.Sp * d.B;
    L.att = 0; L.out = n * d.heads * d.K * d.K; L.ffn = L.out + n * d.K * d.Ds; L.total = L.ffn + n * d.K * d.Ds;
    return L;
}

// phase-timing macro shared by the clip kernels (active only when a.dbg != nullptr; CTA 0 reports)
#define SAVI_PH(id) do { if (a.dbg && blockIdx.x == 0) { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); \
    atomicAdd(reinterpret_cast<unsigned long long*>(a.dbg + (id)), (unsigned long long)(t_ - ph_last)); ph_last = t_; } } } while (0)

struct WgradJob {               // dW[o][c] (+)= alpha * sum_r dY[r][o] * X[r][c]
    const float* dY; const float* X; float* dW;
    int R, O, C, ldy, ldx;
    float alpha;
};
constexpr int WGRAD_MAX_JOBS = 7 + 6 * SAVI_MAX_BLOCKS;
struct WgradArgs { WgradJob job[WGRAD_MAX_JOBS]; int njobs; int rows_per_split; };

// Process-wide development options (include/focus_savi.h: savi_set_option).  Read-only on the forward / backward path; their
// initial values come from the SAVI_* environment variables, read ONCE when the library is loaded.
struct SaviOptions {
    int disable_umma;      // "disable_umma"  (SAVI_DISABLE_UMMA): never dispatch to the tcgen05 clip kernels
    int disable_mma;       // "disable_mma"   (SAVI_DISABLE_MMA):  never dispatch to any tensor-core family
    int no_opstage;        // "no_opstage"    (SAVI_NO_OPSTAGE):   mma.sync family: no operand hand-over buffers
    int no_overlap;        // "no_overlap"    (SAVI_NO_OVERLAP):   d_inputs / weight-gradient kernels after the clip kernel, not overlapped
    int dx_trace;          // "dx_trace"      (SAVI_DX_TRACE):     globaltimer trace of the overlapped kernels (tools/dx_trace.py)
    int dx_tpc;            // "dx_tpc"        (SAVI_DX_TPC):       tiles per d_inputs CTA (0 = automatic)
    int dx_gate_last;      // "dx_gate_last"  (SAVI_DX_GATE_LAST)
};
const SaviOptions& savi_options();

// per-kernel timing hooks (savi_api.cu); no-ops unless savi_profile_enable(1)
void savi_prof_begin(int slot, cudaStream_t st);
void savi_prof_end(int slot, cudaStream_t st);

cudaError_t savi_launch_forward(const FwdArgs& a, const void* inputs, cudaStream_t st, int* launches);
cudaError_t savi_launch_backward(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st, int* launches);
cudaError_t savi_launch_forward_f32(const FwdArgs& a, const void* inputs, cudaStream_t st, int* launches);
cudaError_t savi_launch_forward_bf16(const FwdArgs& a, const void* inputs, cudaStream_t st, int* launches);
cudaError_t savi_launch_bwd_clip_f32(const BwdArgs& a, cudaStream_t st);
cudaError_t savi_launch_bwd_clip_bf16(const BwdArgs& a, cudaStream_t st);
cudaError_t savi_launch_ln_bwd_f32(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st);
cudaError_t savi_launch_ln_bwd_bf16(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st);
cudaError_t savi_launch_dx_mma(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st);
size_t savi_dx_smem_bytes(const Dims& d);
static inline int savi_fwd_kmax(int K) { return K <= 8 ? 8 : K <= 16 ? 16 : K <= 24 ? 24 : K <= 32 ? 32 : 64; }
cudaError_t savi_launch_fwd_umma(const FwdArgs& a, const unsigned char* wimg, const WImg& wi, cudaStream_t st);
int savi_fwd_umma_smem_bytes(const Dims& d);
int savi_umma_mha_fits(int K, int heads);        // predictor attention core of the tcgen05 clip kernels fits its shared-memory windows
cudaError_t savi_launch_bwd_umma(const BwdArgs& a, const unsigned char* wimg, const WImg& wi, cudaStream_t st);
int savi_bwd_umma_smem_bytes(const Dims& d);
cudaError_t savi_launch_dx_umma(const BwdArgs& a, const void* inputs, void* grad_inputs, bool overlap, cudaStream_t st);
bool savi_prof_enabled();
int savi_dx_umma_smem_bytes(int I);
cudaError_t savi_launch_wgrad_umma(const WgradArgs& wa, const int* done, int done_target, cudaStream_t st);
size_t savi_fwd_smem_bytes(const Dims& d, int TN);
size_t savi_bwd_smem_bytes(const Dims& d, int TN);
