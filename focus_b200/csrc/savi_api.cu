// C-ABI entry points (include/focus_savi.h): validation, sizing, parameter packing,
// forward / backward launch sequencing.  No allocation, no host synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cuda_fp16.h>
#include "savi_dx_mma.cuh"
#include "savi_args.h"
#include "savi_smallgemm.cuh"

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
// same, for the other translation units of the library (steve_neighbors.cu)
int savi_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    return fail(SAVI_ECUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

extern "C" int savi_version(void) { return SAVI_VERSION; }
extern "C" const char* savi_last_error(void) { return g_err; }
extern "C" int savi_last_launch_count(void) { return g_launches; }

static const int kMaxSmem = 227 * 1024;

// ---- process-wide options ---------------------------------------------------------------------
static SaviOptions options_from_env() {
    SaviOptions o;
    o.disable_umma = getenv("SAVI_DISABLE_UMMA") ? 1 : 0;
    o.disable_mma = getenv("SAVI_DISABLE_MMA") ? 1 : 0;
    o.no_opstage = getenv("SAVI_NO_OPSTAGE") ? 1 : 0;
    o.no_overlap = getenv("SAVI_NO_OVERLAP") ? 1 : 0;
    o.dx_trace = getenv("SAVI_DX_TRACE") ? 1 : 0;
    o.dx_tpc = getenv("SAVI_DX_TPC") ? atoi(getenv("SAVI_DX_TPC")) : 0;
    o.dx_gate_last = getenv("SAVI_DX_GATE_LAST") ? 1 : 0;
    return o;
}
static SaviOptions g_opt = options_from_env();          // environment read once, at library load
const SaviOptions& savi_options() { return g_opt; }
extern "C" int savi_set_option(const char* name, int value) {
    if (!name) return fail(SAVI_EINVAL, "null option name");
    struct { const char* n; int* v; } tab[] = {{"disable_umma", &g_opt.disable_umma}, {"disable_mma", &g_opt.disable_mma},
        {"no_opstage", &g_opt.no_opstage}, {"no_overlap", &g_opt.no_overlap}, {"dx_trace", &g_opt.dx_trace},
        {"dx_tpc", &g_opt.dx_tpc}, {"dx_gate_last", &g_opt.dx_gate_last}};
    for (auto& t : tab) if (!strcmp(name, t.n)) { *t.v = value; return SAVI_OK; }
    return fail(SAVI_EINVAL, "unknown option '%s'", name);
}
static long long* g_dbg = nullptr;     // device buffer of 64 phase counters (development aid)
extern "C" int savi_debug_set_phase_buffer(void* dev_ptr) { g_dbg = reinterpret_cast<long long*>(dev_ptr); return SAVI_OK; }

// ---- optional per-kernel timing -------------------------------------------------
// process-wide (autograd runs backward on its own thread); meant for single-stream benchmarking only
static bool g_prof = false;
static cudaEvent_t g_ev[SAVI_PROFILE_SLOTS][2];
static bool g_ev_made = false;
static bool g_ev_used[SAVI_PROFILE_SLOTS];

extern "C" int savi_profile_enable(int on) {
    if (on && !g_ev_made) {
        for (int i = 0; i < SAVI_PROFILE_SLOTS; ++i)
            for (int j = 0; j < 2; ++j) {
                cudaError_t e = cudaEventCreate(&g_ev[i][j]);
                if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
            }
        g_ev_made = true;
        for (int i = 0; i < SAVI_PROFILE_SLOTS; ++i) g_ev_used[i] = false;
    }
    g_prof = on != 0;
    return SAVI_OK;
}
void savi_prof_begin(int slot, cudaStream_t st) { if (g_prof) { cudaEventRecord(g_ev[slot][0], st); } }
void savi_prof_end(int slot, cudaStream_t st) { if (g_prof) { cudaEventRecord(g_ev[slot][1], st); g_ev_used[slot] = true; } }
extern "C" int savi_profile_read(float* ms_host, int n) {
    if (!ms_host || n < 0 || n > SAVI_PROFILE_SLOTS) return fail(SAVI_EINVAL, "bad profile buffer");
    for (int i = 0; i < n; ++i) {
        ms_host[i] = -1.f;
        if (g_ev_made && g_ev_used[i]) {
            cudaError_t e = cudaEventSynchronize(g_ev[i][1]);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize");
            e = cudaEventElapsedTime(&ms_host[i], g_ev[i][0], g_ev[i][1]);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventElapsedTime");
        }
    }
    return SAVI_OK;
}

static int choose_cluster(const SaviShape& s) {
    if (s.cluster == 1 || s.cluster == 2 || s.cluster == 4 || s.cluster == 8) return s.cluster;
    int cn = 1;
    // as many CTAs per clip as fit one wave of 148 SMs, while every CTA keeps >= 64 tokens
    while (cn < 8 && (int64_t)s.B * cn * 2 <= 148 && s.N / (cn * 2) >= 64) cn *= 2;
    return cn;
}

static int validate(const SaviShape* s, Dims& d) {
    if (!s) return fail(SAVI_EINVAL, "null shape");
    if (s->B < 1 || s->T < 1 || s->N < 1 || s->I < 1) return fail(SAVI_EINVAL, "B,T,N,I must be >= 1");
    if (s->K < 1 || s->K > SAVI_MAX_SLOTS) return fail(SAVI_EINVAL, "num_slots K=%d outside [1,%d]", s->K, SAVI_MAX_SLOTS);
    if (s->D < 8 || s->D > 512 || s->D % 8) return fail(SAVI_EINVAL, "input_size D=%d must be a multiple of 8 in [8,512]", s->D);
    if (s->Ds < 4 || s->Ds > 512 || s->Ds % 4) return fail(SAVI_EINVAL, "slot_size Ds=%d must be a multiple of 4 in [4,512]", s->Ds);
    if (s->M < 4 || s->M % 4 || s->M > 4096) return fail(SAVI_EINVAL, "mlp_hidden_size M=%d must be a multiple of 4 in [4,4096]", s->M);
    if (s->blocks < 0 || s->blocks > SAVI_MAX_BLOCKS) return fail(SAVI_EINVAL, "num_predictor_blocks=%d outside [0,%d]", s->blocks, SAVI_MAX_BLOCKS);
    // (the reference only builds the multi-head attention when there are predictor blocks: transformer.py:94-101)
    if (s->blocks > 0 && (s->heads < 1 || s->Ds % s->heads)) return fail(SAVI_EINVAL, "slot_size %d not divisible by num_predictor_heads %d", s->Ds, s->heads);
    if ((int64_t)s->B * s->T > 65535) return fail(SAVI_EINVAL, "B*T = %lld frames per call exceed the 65535 limit of the token-parallel kernels' grid (split the batch)", (long long)s->B * s->T);
    if (s->dtype != SAVI_DTYPE_F32 && s->dtype != SAVI_DTYPE_BF16) return fail(SAVI_EINVAL, "unknown dtype %d", s->dtype);
    if (s->cluster != 0 && s->cluster != 1 && s->cluster != 2 && s->cluster != 4 && s->cluster != 8)
        return fail(SAVI_EINVAL, "cluster must be 0,1,2,4 or 8");
    d.B = s->B; d.T = s->T; d.N = s->N; d.D = s->D; d.Ds = s->Ds; d.M = s->M; d.K = s->K; d.I = s->I;
    d.blocks = s->blocks; d.heads = s->blocks > 0 ? s->heads : 1;
    d.KP = (s->K + 3) & ~3;
    d.CN = choose_cluster(*s);
    d.S = s->T * s->I; d.Sm = s->T * (s->I - 1); d.Sp = (s->T - 1) * s->blocks;
    d.tok_bytes = s->dtype == SAVI_DTYPE_F32 ? 4 : 2;
    d.eps = s->eps; d.ln_eps = s->ln_eps; d.qscale = 1.0f / sqrtf((float)s->Ds);
    d.KC = ((s->K + 15) / 16) * 16;
    // tensor-core path: bf16 token stream and shapes the mma.sync tiles cover; everything else takes the SIMT path
    const int MT = d.KC / 16 == 3 ? 4 : d.KC / 16;
    d.mma = (s->dtype == SAVI_DTYPE_BF16 && s->D % 16 == 0 && s->D <= 256 && s->Ds % 8 == 0 && s->M % 8 == 0 &&
             s->N % 8 == 0 && !g_opt.disable_mma &&
             dx_pick_ig(s->I, d.KC, s->D, (size_t)kMaxSmem) >= 1 &&
             tmma_smem_bytes(MT, s->D, s->K, 1, true) <= (size_t)kMaxSmem) ? 1 : 0;
    if (d.mma) d.KC = MT * 16;
    // tcgen05 clip kernels: bf16 tokens, 128-wide features (one M tile per product), slots and scratch that fit shared memory
    d.NTILE = (s->N + 127) / 128;
    d.umma = 0;
    // (any N: ragged and odd token counts are handled by the 128-token tiles; the mma.sync path needs N % 8 == 0)
    if (s->dtype == SAVI_DTYPE_BF16 && s->D == 128 && s->Ds == 128 && s->M == 128 && s->K <= 24 && (s->cluster == 0 || s->cluster <= 2) &&
        savi_umma_mha_fits(s->K, d.heads) && savi_dx_umma_smem_bytes(s->I) <= kMaxSmem && !g_opt.disable_umma && !g_opt.disable_mma) {
        d.umma = 1;
        d.mma = 1;                       // shares the tensor-core workspace layout (staged d(Ux) rows, no fp32 d xhat accumulator)
        d.CN = s->cluster ? s->cluster : ((int64_t)s->B * 2 <= 148 ? 2 : 1);
        if (d.NTILE < d.CN) d.CN = 1;
    }
    return SAVI_OK;
}

size_t savi_fwd_smem_bytes(const Dims& d, int TN) {
    const int KMAX = savi_fwd_kmax(d.K);
    return (size_t)KMAX * d.D * 4 + (size_t)d.KP * d.D * 4 + d.KP * 4 + (size_t)TN * coef_stride(d.KP) * 4 +
           (size_t)TN * tile_stride_bytes(d.D, d.tok_bytes);
}
size_t savi_bwd_smem_bytes(const Dims& d, int TN) {
    const int KMAX = savi_fwd_kmax(d.K);
    return (size_t)2 * KMAX * d.D * 4 + (size_t)d.KP * d.D * 4 + 2 * d.KP * 4 + (size_t)2 * TN * coef_stride(d.KP) * 4 +
           (size_t)TN * tile_stride_bytes(d.D, d.tok_bytes);
}

// tokens per tile: the largest of 256..16 that fits; the same dynamic smem doubles as the linear-layer arena
static int plan_smem(const Dims& d, bool bwd, int* TN, int* arena_floats, int* smem_bytes, int* stages, int* op_bytes, int* op_width) {
    *stages = 1; *op_bytes = 0; *op_width = 0;
    if (d.mma) {
        const int MT = d.KC / 16;
        int cmax = 4 * d.Ds; if (d.M > cmax) cmax = d.M; if (d.D > cmax) cmax = d.D;
        size_t tok = tmma_smem_bytes(MT, d.D, d.K, 2, bwd);
        *stages = 2;
        if (tok > (size_t)kMaxSmem) { tok = tmma_smem_bytes(MT, d.D, d.K, 1, bwd); *stages = 1; }
        size_t want = lin_mma_smem(MT, cmax);
        if (want > (size_t)kMaxSmem) want = (size_t)160 * 1024;     // such layers fall back to the SIMT linear inside the kernel
        size_t bytes = tok > want ? tok : want;
        bytes = (bytes + 15) / 16 * 16;
        // two operand buffers in front of the arena let consecutive linears hand their A operand over in shared memory
        const int cop = d.D > d.Ds ? d.D : d.Ds;
        size_t opb = (lin_mma_smem(MT, cop) + 15) / 16 * 16;
        if (cop % 32 != 0 || bytes + 2 * opb > (size_t)kMaxSmem || bwd || g_opt.no_opstage) opb = 0;   // backward does not use them (yet)
        if (opb == 0 && *stages == 2 && cop % 32 == 0 && !bwd && !g_opt.no_opstage) {      // prefer the hand-over buffers to the second token tile
            const size_t tok1 = tmma_smem_bytes(MT, d.D, d.K, 1, bwd);
            size_t b1 = tok1 > want ? tok1 : want; b1 = (b1 + 15) / 16 * 16;
            const size_t o1 = (lin_mma_smem(MT, cop) + 15) / 16 * 16;
            if (b1 + 2 * o1 <= (size_t)kMaxSmem) { bytes = b1; opb = o1; *stages = 1; }
        }
        *op_bytes = (int)opb; *op_width = opb ? cop : 0;
        *TN = TMMA_TN; *smem_bytes = (int)(bytes + 2 * opb); *arena_floats = (int)(bytes / 4);
        return SAVI_OK;
    }
    int per = ((d.N + d.CN - 1) / d.CN + 3) & ~3;
    int tn = 256;
    while (tn > 16 && (tn / 2 >= per || (bwd ? savi_bwd_smem_bytes(d, tn) : savi_fwd_smem_bytes(d, tn)) > (size_t)kMaxSmem - 1024)) tn /= 2;
    size_t tok = bwd ? savi_bwd_smem_bytes(d, tn) : savi_fwd_smem_bytes(d, tn);
    if (tok > (size_t)kMaxSmem) return fail(SAVI_EINVAL, "shape needs %zu B of shared memory per CTA (> %d)", tok, kMaxSmem);
    int cmax = 4 * d.Ds; if (d.M > cmax) cmax = d.M; if (d.D > cmax) cmax = d.D;
    size_t want = (size_t)d.K * cmax * 4;                // whole [K, Cmax] operand resident -> single pass per linear
    if (want > 160 * 1024) want = 160 * 1024;
    if (want < (size_t)8 * cmax * 4) want = (size_t)8 * cmax * 4;
    size_t bytes = tok > want ? tok : want;
    bytes = (bytes + 15) / 16 * 16;
    *TN = tn; *smem_bytes = (int)bytes; *arena_floats = (int)(bytes / 4);
    return SAVI_OK;
}

extern "C" int savi_query(const SaviShape* shape, SaviSizes* sizes) {
    Dims d;
    int rc = validate(shape, d);
    if (rc) return rc;
    if (!sizes) return fail(SAVI_EINVAL, "null sizes");
    ParamOff po; savi_param_offsets(*shape, po);
    SavedLayout sl; savi_saved_layout(d, sl);
    FwdWsLayout fl; savi_fwd_ws_layout(d, fl);
    BwdWsLayout bl; savi_bwd_ws_layout(d, bl);
    int tn, af, sb;
    int stg, opb, opw;
    if ((rc = plan_smem(d, false, &tn, &af, &sb, &stg, &opb, &opw))) return rc;
    if ((rc = plan_smem(d, true, &tn, &af, &sb, &stg, &opb, &opw))) return rc;
    sizes->n_params = 21 + 12 * shape->blocks;
    sizes->param_floats = po.total;
    sizes->packed_bytes = (int64_t)po.packed_total * 8;      // fp32 + bf16 hi + bf16 lo images
    if (d.umma) {                                            // + the blocked SWIZZLE_128B operand images of the tcgen05 path
        WImg wi; savi_wimg_layout(d.D, d.Ds, d.M, d.blocks, wi);
        sizes->packed_bytes = savi_wimg_base(po.packed_total) + wi.total_bytes;
    }
    sizes->saved_bytes = sl.total_bytes;
    sizes->fwd_ws_bytes = fl.total_bytes;
    sizes->bwd_ws_bytes = bl.total_bytes;
    sizes->cluster = d.CN;
    sizes->path = d.umma ? SAVI_PATH_TCGEN05 : d.mma ? SAVI_PATH_MMA_SYNC : SAVI_PATH_SIMT;
    sizes->dropout_floats = savi_dropout_layout(d).total;
    return SAVI_OK;
}

// parameter tensors in state_dict order: (offset, numel) pairs
static int param_table(const SaviShape& s, const ParamOff& o, int64_t* off, int64_t* num) {
    const int D = s.D, Ds = s.Ds, M = s.M;
    int n = 0;
    auto put = [&](int of, int ne) { off[n] = of; num[n] = ne; ++n; };
    put(o.slot_mu, Ds); put(o.slot_log_sigma, Ds);
    put(o.ln_in_w, D); put(o.ln_in_b, D); put(o.ln_s_w, Ds); put(o.ln_s_b, Ds); put(o.ln_m_w, Ds); put(o.ln_m_b, Ds);
    put(o.wq, Ds * Ds); put(o.wk, Ds * D); put(o.wv, Ds * D);
    put(o.wih, 3 * Ds * Ds); put(o.whh, 3 * Ds * Ds); put(o.bih, 3 * Ds); put(o.bhh, 3 * Ds);
    put(o.w1, M * Ds); put(o.b1, M); put(o.w2, Ds * M); put(o.b2, Ds);
    for (int j = 0; j < s.blocks; ++j) {
        const BlockOff& b = o.blk[j];
        put(b.ln1_w, Ds); put(b.ln1_b, Ds);
        put(b.pq, Ds * Ds); put(b.pk, Ds * Ds); put(b.pv, Ds * Ds); put(b.po, Ds * Ds);
        put(b.ln2_w, Ds); put(b.ln2_b, Ds);
        put(b.f1, 4 * Ds * Ds); put(b.f1b, 4 * Ds); put(b.f2, 4 * Ds * Ds); put(b.f2b, Ds);
    }
    put(o.lnf_w, Ds); put(o.lnf_b, Ds);
    return n;
}

extern "C" int savi_param_layout(const SaviShape* shape, int64_t* offsets_host, int64_t* numels_host) {
    Dims d;
    int rc = validate(shape, d);
    if (rc) return rc;
    if (!offsets_host || !numels_host) return fail(SAVI_EINVAL, "null output array");
    ParamOff po; savi_param_offsets(*shape, po);
    param_table(*shape, po, offsets_host, numels_host);
    return SAVI_OK;
}

// ---------------------------------------------------------------------------
// parameter packing
// ---------------------------------------------------------------------------
constexpr int MAXP = 21 + 12 * SAVI_MAX_BLOCKS;
struct PackArgs { const float* src[MAXP]; int off[MAXP]; int num[MAXP]; int count; };
struct TrJob { int src, dst, rows, cols; };
constexpr int MAXTR = 9 + 6 * SAVI_MAX_BLOCKS;
struct TrArgs { TrJob job[MAXTR]; int count; };

__global__ void pack_copy_kernel(const __grid_constant__ PackArgs pa, float* __restrict__ packed) {
    const int p = blockIdx.y;
    const float* __restrict__ src = pa.src[p];
    float* dst = packed + pa.off[p];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pa.num[p]; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
// dst[c][r] = src[r][c]
__global__ void pack_transpose_kernel(const __grid_constant__ TrArgs ta, float* __restrict__ packed) {
    const TrJob j = ta.job[blockIdx.y];
    const float* src = packed + j.src;
    float* dst = packed + j.dst;
    const int n = j.rows * j.cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / j.cols, c = i - r * j.cols;
        dst[(size_t)c * j.rows + r] = src[i];
    }
}

// packed fp32 -> bf16 hi / lo images (x = hi + lo up to 2^-17 |x|)
__global__ void pack_split_kernel(const float* __restrict__ packed, bf16* __restrict__ hi, bf16* __restrict__ lo, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        bf16 h, l;
        split_bf16(packed[i], h, l);
        hi[i] = h; lo[i] = l;
    }
}

// fp32 matrix A[R][C] -> blocked SWIZZLE_128B 16-bit image (savi_layout.h: WImg): fp16 forward orientation, bf16 backward.  tr = 0: A is row-major at packed + src (leading dimension
// ld); tr = 1: A is the TRANSPOSE of the row-major [C][R] matrix stored there (the backward-orientation images are read
// straight from the original weights: no transposed fp32 copy is needed on the tcgen05 path).
struct ImgJob { int src, R, C, ld, tr; long long dst; };
constexpr int MAXIMG = 2 * (7 + 6 * SAVI_MAX_BLOCKS);
struct ImgArgs { ImgJob job[MAXIMG]; int count; };
__global__ void pack_image_kernel(const __grid_constant__ ImgArgs ia, const float* __restrict__ packed, unsigned char* __restrict__ img) {
    const ImgJob j = ia.job[blockIdx.y];
    const int c8n = j.C >> 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.R * c8n; i += gridDim.x * blockDim.x) {
        int r, c0;
        float v[8];
        if (!j.tr) {
            r = i / c8n; c0 = (i - r * c8n) * 8;
            const float* src = packed + j.src + (size_t)r * j.ld + c0;
            const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
            v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
        } else {
            c0 = (i / j.R) * 8; r = i - (i / j.R) * j.R;            // consecutive threads read consecutive r: coalesced
            const float* src = packed + j.src + (size_t)c0 * j.ld + r;
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = src[(size_t)e * j.ld];
        }
        unsigned char* blk = img + j.dst + ((size_t)((r >> 7) * (j.C >> 6) + (c0 >> 6)) * WIMG_NB) * UMMA_BLK;
        const unsigned off = (unsigned)(r & 127) * 128u + (((unsigned)((c0 & 63) >> 3) ^ (unsigned)(r & 7)) << 4);
        uint4 out;
        if (j.tr) {                                  // backward orientation: bf16 (savi_layout.h)
            __nv_bfloat162 b2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) b2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            out = *reinterpret_cast<const uint4*>(b2);
        } else {                                     // forward orientation: fp16
            // saturating conversion: a weight beyond fp16's range (65504) stays finite instead of turning the product into inf / NaN
            unsigned h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2[e]) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
            out = make_uint4(h2[0], h2[1], h2[2], h2[3]);
        }
        *reinterpret_cast<uint4*>(blk + off) = out;
    }
}

static int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10) return fail(SAVI_EARCH, "device compute capability %d.x is not sm_100 (B200)", major);
    return SAVI_OK;
}

extern "C" int savi_pack_params(const SaviShape* shape, const void* const* param_ptrs_host, void* packed, void* stream) {
    Dims d;
    int rc = validate(shape, d);
    if (rc) return rc;
    if (!param_ptrs_host || !packed) return fail(SAVI_EINVAL, "null pointer");
    if ((rc = check_device())) return rc;
    ParamOff po; savi_param_offsets(*shape, po);
    int64_t off[MAXP], num[MAXP];
    PackArgs pa;
    pa.count = param_table(*shape, po, off, num);
    int maxn = 0;
    for (int i = 0; i < pa.count; ++i) {
        if (!param_ptrs_host[i]) return fail(SAVI_EINVAL, "parameter %d is null", i);
        pa.src[i] = reinterpret_cast<const float*>(param_ptrs_host[i]);
        pa.off[i] = (int)off[i]; pa.num[i] = (int)num[i];
        if (pa.num[i] > maxn) maxn = pa.num[i];
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int gx = (maxn + 255) / 256; if (gx > 64) gx = 64;
    savi_prof_begin(0, st);
    pack_copy_kernel<<<dim3(gx, pa.count), 256, 0, st>>>(pa, reinterpret_cast<float*>(packed));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "pack_copy_kernel");
    const int D = shape->D, Ds = shape->Ds, M = shape->M;
    g_launches = 1;                                          // pack_copy_kernel
    if (d.umma) {
        // folded weights of the tcgen05 path (ParamOff::wqk, wg), fp32, from the packed originals
        float* Pk = reinterpret_cast<float*>(packed);
        SmallGemmArgs ga; ga.count = 2;
        ga.g[0] = SmallGemm{Pk + po.wk, Pk + po.wq, Pk + po.wqk, D, Ds, Ds, 1, D, Ds, 1, d.qscale};       // wqk[d][c] = s sum_a Wk[a][d] Wq[a][c]
        ga.g[1] = SmallGemm{Pk + po.wih, Pk + po.wv, Pk + po.wg, 3 * Ds, D, Ds, Ds, 1, D, 1, 1.0f};        // wg[g][d] = sum_a Wih[g][a] Wv[a][d]
        e = launch_small_gemms(ga, st);
        if (e != cudaSuccess) return cuda_fail(e, "small_gemm_kernel (weight folds)");
        g_launches += 1;
    }
    if (!d.umma) {
        // transposed fp32 copies + bf16 hi / lo copies: operands of the SIMT and mma.sync clip kernels
        TrArgs ta; ta.count = 0;
        auto tr = [&](int src, int dst, int rows, int cols) { ta.job[ta.count++] = TrJob{src, dst, rows, cols}; };
        tr(po.wq, po.wq_t, Ds, Ds); tr(po.wk, po.wk_t, Ds, D); tr(po.wv, po.wv_t, Ds, D);
        tr(po.wih, po.wih_t, 3 * Ds, Ds); tr(po.whh, po.whh_t, 3 * Ds, Ds);
        tr(po.w1, po.w1_t, M, Ds); tr(po.w2, po.w2_t, Ds, M);
        for (int j = 0; j < shape->blocks; ++j) {
            const BlockOff& b = po.blk[j]; const BlockOffT& t = po.blkt[j];
            tr(b.pq, t.pq_t, Ds, Ds); tr(b.pk, t.pk_t, Ds, Ds); tr(b.pv, t.pv_t, Ds, Ds); tr(b.po, t.po_t, Ds, Ds);
            tr(b.f1, t.f1_t, 4 * Ds, Ds); tr(b.f2, t.f2_t, Ds, 4 * Ds);
        }
        pack_transpose_kernel<<<dim3(64, ta.count), 256, 0, st>>>(ta, reinterpret_cast<float*>(packed));
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "pack_transpose_kernel");
        bf16* hi = reinterpret_cast<bf16*>(reinterpret_cast<float*>(packed) + po.packed_total);
        pack_split_kernel<<<148, 512, 0, st>>>(reinterpret_cast<const float*>(packed), hi, hi + po.packed_total, po.packed_total);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "pack_split_kernel");
        g_launches += 2;
    } else {
        // tcgen05 path: every weight the clip kernels multiply by, as blocked SWIZZLE_128B bf16 hi / lo operand images
        WImg wi; savi_wimg_layout(D, Ds, M, shape->blocks, wi);
        ImgArgs ia; ia.count = 0;
        auto im = [&](int src, int R, int C, int64_t dst) { ia.job[ia.count++] = ImgJob{src, R, C, C, 0, (long long)dst}; };
        auto imT = [&](int src, int R, int C, int64_t dst) { ia.job[ia.count++] = ImgJob{src, R, C, R, 1, (long long)dst}; };   // image of src^T, src = [C][R]
        // forward orientation: rows = output feature of the product
        im(po.wqk, D, Ds, wi.wqk); im(po.wg, 3 * Ds, D, wi.wg);
        im(po.whh, 3 * Ds, Ds, wi.whh); im(po.w1, M, Ds, wi.w1); im(po.w2, Ds, M, wi.w2);
        // backward orientation: rows = input feature (dX = dY W)
        imT(po.wqk, Ds, D, wi.wqkT); imT(po.wg, D, 3 * Ds, wi.wgT);
        imT(po.whh, Ds, 3 * Ds, wi.whhT); imT(po.w1, Ds, M, wi.w1T); imT(po.w2, M, Ds, wi.w2T);
        for (int j = 0; j < shape->blocks; ++j) {
            const BlockOff& b = po.blk[j];
            const WImgBlock& f = wi.blk[j]; const WImgBlock& r = wi.blkT[j];
            im(b.pq, Ds, Ds, f.pq); im(b.pk, Ds, Ds, f.pk); im(b.pv, Ds, Ds, f.pv); im(b.po, Ds, Ds, f.po);
            im(b.f1, 4 * Ds, Ds, f.f1); im(b.f2, Ds, 4 * Ds, f.f2);
            imT(b.pq, Ds, Ds, r.pq); imT(b.pk, Ds, Ds, r.pk); imT(b.pv, Ds, Ds, r.pv); imT(b.po, Ds, Ds, r.po);
            imT(b.f1, Ds, 4 * Ds, r.f1); imT(b.f2, 4 * Ds, Ds, r.f2);
        }
        pack_image_kernel<<<dim3(16, ia.count), 256, 0, st>>>(ia, reinterpret_cast<const float*>(packed),
                                                              reinterpret_cast<unsigned char*>(packed) + savi_wimg_base(po.packed_total));
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e, "pack_image_kernel");
        g_launches += 1;
    }
    savi_prof_end(0, st);
    return SAVI_OK;
}

extern "C" int savi_forward(const SaviShape* shape, const void* packed, const void* inputs, const void* noise,
                            void* slots_out, void* attn_out, void* saved, void* fwd_ws, const void* dropout_masks, void* stream) {
    FwdArgs a;
    int rc = validate(shape, a.d);
    if (rc) return rc;
    if (!packed || !inputs || !noise || !slots_out || !attn_out || !saved || !fwd_ws) return fail(SAVI_EINVAL, "null pointer");
    if ((rc = check_device())) return rc;
    savi_param_offsets(*shape, a.po);
    savi_saved_layout(a.d, a.sl);
    savi_fwd_ws_layout(a.d, a.wl);
    if ((rc = plan_smem(a.d, false, &a.TN, &a.arena_floats, &a.smem_bytes, &a.stages, &a.op_bytes, &a.op_width))) return rc;
    a.packed = reinterpret_cast<const float*>(packed);
    a.noise = reinterpret_cast<const float*>(noise);
    a.slots_out = reinterpret_cast<float*>(slots_out);
    a.attn_out = attn_out;
    a.saved = reinterpret_cast<unsigned char*>(saved);
    a.ws = reinterpret_cast<float*>(fwd_ws);
    a.dbg = g_dbg;
    a.drop = reinterpret_cast<const float*>(dropout_masks);
    g_launches = 0;
    cudaError_t e = savi_launch_forward(a, inputs, reinterpret_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return cuda_fail(e, "savi_forward launch");
    return SAVI_OK;
}

extern "C" int savi_backward(const SaviShape* shape, const void* packed, const void* inputs, const void* noise,
                             const void* saved, const void* grad_slots, const void* grad_attn, void* grad_inputs,
                             void* grad_params, void* grad_noise, void* bwd_ws, const void* dropout_masks, void* stream) {
    BwdArgs a;
    int rc = validate(shape, a.d);
    if (rc) return rc;
    if (!packed || !inputs || !noise || !saved || !grad_slots || !grad_inputs || !grad_params || !bwd_ws)
        return fail(SAVI_EINVAL, "null pointer");
    if ((rc = check_device())) return rc;
    savi_param_offsets(*shape, a.po);
    savi_saved_layout(a.d, a.sl);
    savi_bwd_ws_layout(a.d, a.wl);
    if ((rc = plan_smem(a.d, true, &a.TN, &a.arena_floats, &a.smem_bytes, &a.stages, &a.op_bytes, &a.op_width))) return rc;
    a.packed = reinterpret_cast<const float*>(packed);
    a.noise = reinterpret_cast<const float*>(noise);
    a.saved = reinterpret_cast<const unsigned char*>(saved);
    a.grad_slots = reinterpret_cast<const float*>(grad_slots);
    a.grad_attn = grad_attn;
    a.grad_params = reinterpret_cast<float*>(grad_params);
    a.grad_noise = reinterpret_cast<float*>(grad_noise);
    a.ws = reinterpret_cast<float*>(bwd_ws);
    a.dbg = g_dbg;
    a.drop = reinterpret_cast<const float*>(dropout_masks);
    g_launches = 0;
    cudaError_t e = savi_launch_backward(a, inputs, grad_inputs, reinterpret_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return cuda_fail(e, "savi_backward launch");
    return SAVI_OK;
}
