/*
 * focus_savi.h — C ABI of the B200-native video slot-attention encoder.
 *
 * Drop-in boundary.  The reference (srv902/FOCUS) has no FFI:
 * its boundary is the Python nn.Module
 *     slowfast/models/STEVE/steve.py:11-105   class SlotAttentionVideo
 * constructed at steve.py:229-232 and called at steve.py:311 (STEVE.forward)
 * and steve.py:346 (STEVE.encode).  focus_b200.SlotAttentionVideo keeps that
 * constructor / state_dict / forward contract and calls the entry points below
 * through ctypes from a torch.autograd.Function.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns all memory (outputs, saved-for-backward, workspaces); the
 *    library allocates no device memory.  Its only mutable global state is opt-in
 *    and process-wide: the development options of savi_set_option(), the event
 *    pairs of savi_profile_enable() and the savi_debug_set_phase_buffer() pointer.
 *    With none of them touched (the default) savi_query / savi_pack_params /
 *    savi_forward / savi_backward read nothing but their arguments;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host
 *    synchronisation inside; the launch sequence is shape-static, so a stream
 *    capture of pack + forward + backward replays as a CUDA graph
 *    (tests/test_cabi_gpu.py) and one process per GPU needs no coordination;
 *  - every entry point returns 0 or a negative SAVI_E* code; savi_last_error()
 *    returns a thread-local message.  Nothing aborts, throws or falls back to a
 *    CPU / library path.
 */
#ifndef FOCUS_SAVI_H
#define FOCUS_SAVI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAVI_VERSION 2   /* 2: savi_forward / savi_backward take the predictor dropout masks */
#define SAVI_MAX_BLOCKS 4   /* predictor blocks (reference default 4: slowfast/config/defaults.py:41-62) */
#define SAVI_MAX_SLOTS 64

#define SAVI_OK       0
#define SAVI_EINVAL  -1     /* unsupported shape / alignment / null pointer      */
#define SAVI_EARCH   -2     /* device is not sm_100                              */
#define SAVI_ECUDA   -3     /* a CUDA runtime call failed (see savi_last_error)  */

#define SAVI_DTYPE_F32  0   /* tokens (inputs, xhat, attention map) in fp32       */
#define SAVI_DTYPE_BF16 1   /* tokens in bf16; slot state and accumulation fp32   */

/* Problem description.  Mirrors the constructor arguments of the reference
 * module (steve.py:13-18) plus the call-time sizes of `inputs` [B,T,N,D]. */
/* Kernel family chosen for a shape (SaviSizes.path); every family is hand-written sm_100a CUDA, there is no fallback. */
#define SAVI_PATH_SIMT     0   /* fp32 tokens, or shapes outside the tensor-core tiles: fp32 FMA clip kernels      */
#define SAVI_PATH_MMA_SYNC 1   /* bf16 tokens, general D/Ds/M/K: warp-level mma.sync clip kernels                  */
#define SAVI_PATH_TCGEN05  2   /* bf16 tokens, D = Ds = M = 128, K <= 24: tcgen05 / TMEM clip kernels                */

typedef struct {
    int32_t B;        /* clips in this call (per GPU)                               */
    int32_t T;        /* frames per clip                                            */
    int32_t N;        /* tokens per frame  (num_inputs,  steve.py:53)               */
    int32_t D;        /* input_size                                                 */
    int32_t Ds;       /* slot_size                                                  */
    int32_t M;        /* mlp_hidden_size                                            */
    int32_t K;        /* num_slots                                                  */
    int32_t I;        /* num_iterations                                             */
    int32_t blocks;   /* num_predictor_blocks (0..SAVI_MAX_BLOCKS)                  */
    int32_t heads;    /* num_predictor_heads                                        */
    int32_t dtype;    /* SAVI_DTYPE_*                                               */
    int32_t cluster;  /* CTAs cooperating on one clip (1,2,4,8); 0 = choose         */
    float   eps;      /* epsilon added after the slot softmax (steve.py:18, 81)     */
    float   ln_eps;   /* LayerNorm eps (torch default 1e-5)                         */
} SaviShape;

/* Sizes the caller must allocate (bytes unless noted). */
typedef struct {
    int64_t n_params;        /* number of parameter tensors = 21 + 12*blocks          */
    int64_t param_floats;    /* floats in the flat parameter / gradient buffer        */
    int64_t packed_bytes;    /* savi_pack_params output (flat params + transposes)    */
    int64_t saved_bytes;     /* written by forward, read by backward                  */
    int64_t fwd_ws_bytes;    /* forward scratch (contents dead after forward)         */
    int64_t bwd_ws_bytes;    /* backward scratch                                      */
    int32_t cluster;         /* cluster size that will be used                        */
    int32_t path;            /* SAVI_PATH_*: which kernel family this shape runs on   */
    int64_t dropout_floats;  /* floats of the predictor dropout-mask buffer (below)   */
} SaviSizes;

int savi_version(void);
const char* savi_last_error(void);

/* Validates the shape and fills `sizes`.  No GPU needed. */
int savi_query(const SaviShape* shape, SaviSizes* sizes);

/* Offsets (in floats) of each parameter tensor inside the flat parameter /
 * gradient buffer, in reference state_dict order (steve.py:28-50):
 * slot_mu, slot_log_sigma, norm_inputs.{weight,bias}, norm_slots.*, norm_mlp.*,
 * project_q/k/v.weight, gru.{weight_ih,weight_hh,bias_ih,bias_hh},
 * mlp.0.{weight,bias}, mlp.2.{weight,bias}, then per predictor block
 * attn_layer_norm.{w,b}, attn.proj_{q,k,v,o}.weight, ffn_layer_norm.{w,b},
 * ffn.0.{w,b}, ffn.2.{w,b}, and finally predictor.layer_norm.{w,b}.
 * `offsets_host` and `numels_host` are HOST arrays with sizes.n_params entries. */
int savi_param_layout(const SaviShape* shape, int64_t* offsets_host, int64_t* numels_host);

/* Gathers the module's fp32 parameter tensors (HOST array of n_params DEVICE
 * pointers, state_dict order) into `packed` (flat copy + the transposed weight
 * copies the kernels stream).  One call per forward. */
int savi_pack_params(const SaviShape* shape, const void* const* param_ptrs_host, void* packed, void* stream);

/* Forward of SlotAttentionVideo.forward (steve.py:52-105) for one batch of clips.
 *   inputs    [B,T,N,D]   token dtype, contiguous
 *   noise     [B,K,Ds]    fp32, the N(0,1) draw of steve.py:56
 *   slots_out [B,T,K,Ds]  fp32
 *   attn_out  [B,T,N,K]   token dtype (softmax over slots, pre-epsilon; steve.py:77,96)
 *   dropout_masks  NULL (evaluation, or dropout = 0), or the predictor's dropout masks for a training-mode forward
 *             (transformer.py:12-13,44,48,68): fp32 [sizes.dropout_floats] = att [Sp,B,heads,K,K] | out [Sp,B,K,Ds] |
 *             ffn [Sp,B,K,Ds] with Sp = (T-1)*blocks, block evaluation index f = j*(T-1)+t; every entry 0 or 1/(1-p).
 *             The caller draws them (the Python wrapper with the reference's own RNG consumption order). */
int savi_forward(const SaviShape* shape, const void* packed, const void* inputs, const void* noise,
                 void* slots_out, void* attn_out, void* saved, void* fwd_ws, const void* dropout_masks, void* stream);

/* Backward (the reference relies on autograd; steve.py has no explicit backward).
 *   grad_slots  [B,T,K,Ds] fp32
 *   grad_attn   [B,T,N,K]  token dtype, or NULL (the trainer's case: SURVEY.md §3.1)
 *   grad_inputs [B,T,N,D]  token dtype
 *   grad_params flat fp32 buffer of sizes.param_floats (fully overwritten)
 *   grad_noise  [B,K,Ds]   fp32 or NULL
 *   dropout_masks  the buffer the forward was given (NULL if none) */
int savi_backward(const SaviShape* shape, const void* packed, const void* inputs, const void* noise,
                  const void* saved, const void* grad_slots, const void* grad_attn,
                  void* grad_inputs, void* grad_params, void* grad_noise, void* bwd_ws, const void* dropout_masks, void* stream);

/* Number of kernel launches the last savi_forward / savi_backward call on this
 * thread enqueued (bench.py reports it as gpu_launches). */
int savi_last_launch_count(void);

/* Optional per-kernel timing (bench.py's roofline leg).  When enabled, every kernel
 * the library launches is bracketed by cudaEventRecord on the launch stream (no host
 * sync).  savi_profile_read synchronises on the recorded events and returns the
 * duration in milliseconds of each kernel of the LAST forward and LAST backward call:
 *   [0] pack_params  [1] ln_tokens_fwd  [2] savi_fwd (clip recurrence)
 *   [3] savi_bwd (clip BPTT)  [4] wgrad  [5] ln_tokens_bwd
 * Entries that did not run are -1.  n <= SAVI_PROFILE_SLOTS. */
#define SAVI_PROFILE_SLOTS 6
int savi_profile_enable(int on);
int savi_profile_read(float* ms_host, int n);

/* Process-wide development options (A/B measurements, tests).  Initial values come from the SAVI_<NAME> environment
 * variables, read once when the library is loaded; nothing on the forward / backward path reads the environment.
 *   "disable_umma"  never dispatch to the tcgen05 clip kernels      "disable_mma"  never use a tensor-core family
 *   "no_overlap"    run d_inputs / weight gradients after the backward clip kernel instead of overlapped with it
 *   "no_opstage", "dx_trace", "dx_tpc", "dx_gate_last"  kernel-tuning knobs (focus_b200/csrc/savi_args.h)
 * Returns SAVI_EINVAL for an unknown name.  Not thread-safe against concurrent forward / backward calls. */
int savi_set_option(const char* name, int value);

/* Data-parallel exchange step (replaces the gradient all-reduce DistributedDataParallel performs for the
 * reference, slowfast/models/build.py:79-83): out = scale * sum over ranks of the flat gradient buffer, as ONE kernel over
 * NVLink / NVSwitch peer memory.  Every rank calls it once per step on its own stream with
 *   peer_bufs_host[r]    this process's mapping of rank r's gradient buffer (symmetric memory: same size on every rank;
 *                        entry [rank] is the local buffer savi_backward wrote grad_params_flat into),
 *   signal_pads_host[r]  this process's mapping of rank r's zero-initialised signal pad (signal_pad_bytes each; the kernel
 *                        uses 2 * blocks * world 32-bit slots of it and leaves them zero),
 *   multicast            multicast (NVLS) address of the buffer, or NULL to read the peers one by one,
 *   out                  local result, n_floats fp32 (n_floats % 4 == 0); must not alias the gradient buffer.
 * Both arrays are HOST arrays of `world` device pointers (world <= 16).  The kernel waits for all ranks to arrive (so each
 * peer's backward, earlier on that peer's stream, is complete) and does not finish before every peer has read this rank's
 * buffer, so the buffer may be overwritten by the next step.  No NCCL, no host synchronisation; capturable in a CUDA graph. */
int savi_allreduce_peers(const void* const* peer_bufs_host, void* const* signal_pads_host, const void* multicast,
                         int rank, int world, void* out, int64_t n_floats, float scale, int64_t signal_pad_bytes, void* stream);

/* Development aid: when set to a device buffer of 64 int64 counters, CTA 0 of the clip kernels
 * accumulates the SM cycles spent in each phase of the recurrence (tools/phase_times.py).
 * Pass NULL to disable (default).  Adds barriers; never enable while benchmarking.  The tcgen05 clip kernels carry their
 * probes only in a library built with -DSAVI_PHASE_PROFILE (the tool builds one); in the production build they are compiled out. */
int savi_debug_set_phase_buffer(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* FOCUS_SAVI_H */
