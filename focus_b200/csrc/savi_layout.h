// Memory layout of the flat parameter buffer, the saved-for-backward buffer and
// the workspaces.  Shared by host (size queries, launches) and device code.
//
// Row-major everywhere.  "Field arrays" are 2-D [rows, width] fp32 arrays whose
// row index is (step * B + b) * K + k, so that every per-step slot-side tensor of
// every clip lines up as ONE tall matrix: the weight-gradient kernel then reduces
// dW = dY^T X over all steps and clips as a single GEMM per weight.
#pragma once
#include <stdint.h>
#include "focus_savi.h"

#if defined(__CUDACC__)
#define SAVI_HD __host__ __device__ __forceinline__
#else
#define SAVI_HD inline
#endif

struct BlockOff {           // offsets (floats) of one predictor block's tensors
    int ln1_w, ln1_b, pq, pk, pv, po, ln2_w, ln2_b, f1, f1b, f2, f2b;
};
struct BlockOffT {          // transposed copies ([in][out]) used by the forward linears
    int pq_t, pk_t, pv_t, po_t, f1_t, f2_t;
};

struct ParamOff {
    // flat parameter / gradient buffer, reference state_dict order (steve.py:28-50)
    int slot_mu, slot_log_sigma;
    int ln_in_w, ln_in_b, ln_s_w, ln_s_b, ln_m_w, ln_m_b;
    int wq, wk, wv, wih, whh, bih, bhh, w1, b1, w2, b2;
    BlockOff blk[SAVI_MAX_BLOCKS];
    int lnf_w, lnf_b;
    int total;               // floats in the flat buffer
    // packed-only region: transposed weights
    int wq_t, wk_t, wv_t, wih_t, whh_t, w1_t, w2_t;
    BlockOffT blkt[SAVI_MAX_BLOCKS];
    // packed-only region, tcgen05 path: products of weights that only ever act in sequence (built by pack_fold_kernel)
    //   wqk[d][c] = Ds^-1/2 sum_a Wk[a][d] Wq[a][c]   (qk = s~ . wqk^T :  project_q then the folded project_k, steve.py:61-63,75-76)
    //   wg[g][d]  = sum_a Wih[g][a] Wv[a][d]          (gi = Ux . wg^T  :  the folded project_v then gru.weight_ih, :83,87)
    int wqk, wg, wqk_t, wg_t;
    int packed_total;        // floats in the packed buffer
};

static SAVI_HD int savi_align(int64_t x, int a) { return (int)((x + a - 1) / a * a); }

// ---------------------------------------------------------------------------
// tcgen05 path: every weight the clip kernels multiply by is also kept as a "blocked image": the
// matrix A[R][C] (R = the 128-row M dimension of the MMA, C = contraction) is cut into [128 x 64]
// panels, each stored as ONE 16 KB SWIZZLE_128B K-major operand block, in the order the kernels consume
// them: for rt in R/128: for cb in C/64.  One 1-D bulk copy moves a block from L2 into the shared-memory
// ring, ready for tcgen05.mma.  Every CTA re-streams these blocks on every step of the recurrence, so the
// image is a single 16-bit value per weight (round 1 kept a bf16 hi + lo pair: twice the bytes, 3 MMAs):
//   * forward orientation (rows = output feature): fp16, 11 significand bits.  kind::f16 needs A and B in the
//     same format (a mixed fp16 x bf16 MMA is an illegal instruction on sm_100a), so the forward's slot-side
//     activation operands are fp16 hi | lo as well; forward activations are O(1..10^2), far inside fp16's range.
//     Rounding the weights to fp16 moves the slots by 2e-3 at C1 (oracle, weight_dtype="f16"); bf16 would move
//     them by 3e-2, over the 2e-2 bar, because the recurrence amplifies operand rounding.
//   * backward orientation (rows = input feature, dX = dY . W): bf16.  Gradient operands need bf16's exponent
//     range (their scale is the caller's loss scale), and the backward is far less sensitive: dX = dY . bf16(W)
//     moves d_inputs by 6e-3 and the parameter gradients by <= 7e-3 at C1 (oracle, bwd_weight_dtype="bf16").
// ---------------------------------------------------------------------------
constexpr int64_t UMMA_BLK = 16384;
constexpr int WIMG_NB = 1;                                   // 16 KB blocks per [128 x 64] weight panel
static SAVI_HD int64_t wimg_bytes(int R, int C) { return (int64_t)(R / 128) * (C / 64) * WIMG_NB * UMMA_BLK; }
struct WImgBlock { int64_t pq, pk, pv, po, f1, f2; };
struct WImg {                // byte offsets from the image base
    // forward orientation: rows = output feature (wqk, wg: the folded products of ParamOff)
    int64_t wqk, wg, whh, w1, w2;
    WImgBlock blk[SAVI_MAX_BLOCKS];
    // backward orientation: rows = input feature (dX = dY . W)
    int64_t wqkT, wgT, whhT, w1T, w2T;
    WImgBlock blkT[SAVI_MAX_BLOCKS];
    int64_t total_bytes;
};
static inline void savi_wimg_layout(int D, int Ds, int M, int blocks, WImg& w) {
    int64_t p = 0;
    auto take = [&](int R, int C) { int64_t r = p; p += wimg_bytes(R, C); return r; };
    w.wqk = take(D, Ds); w.wg = take(3 * Ds, D); w.whh = take(3 * Ds, Ds);
    w.w1 = take(M, Ds); w.w2 = take(Ds, M);
    for (int j = 0; j < SAVI_MAX_BLOCKS; ++j) {
        WImgBlock& b = w.blk[j];
        if (j < blocks) { b.pq = take(Ds, Ds); b.pk = take(Ds, Ds); b.pv = take(Ds, Ds); b.po = take(Ds, Ds); b.f1 = take(4 * Ds, Ds); b.f2 = take(Ds, 4 * Ds); }
        else b = WImgBlock{0, 0, 0, 0, 0, 0};
    }
    w.wqkT = take(Ds, D); w.wgT = take(D, 3 * Ds); w.whhT = take(Ds, 3 * Ds);
    w.w1T = take(Ds, M); w.w2T = take(M, Ds);
    for (int j = 0; j < SAVI_MAX_BLOCKS; ++j) {
        WImgBlock& b = w.blkT[j];
        if (j < blocks) { b.pq = take(Ds, Ds); b.pk = take(Ds, Ds); b.pv = take(Ds, Ds); b.po = take(Ds, Ds); b.f1 = take(Ds, 4 * Ds); b.f2 = take(4 * Ds, Ds); }
        else b = WImgBlock{0, 0, 0, 0, 0, 0};
    }
    w.total_bytes = p;
}
// the image region follows the fp32 / bf16-hi / bf16-lo copies in the packed buffer
static SAVI_HD int64_t savi_wimg_base(int packed_total) { return ((int64_t)packed_total * 8 + 1023) / 1024 * 1024; }

static inline void savi_param_offsets(const SaviShape& s, ParamOff& o) {
    const int D = s.D, Ds = s.Ds, M = s.M;
    int p = 0;
    auto take = [&](int n) { int r = p; p += n; return r; };   // tensors stay 4-float aligned because all dims are %4
    o.slot_mu = take(Ds); o.slot_log_sigma = take(Ds);
    o.ln_in_w = take(D);  o.ln_in_b = take(D);
    o.ln_s_w = take(Ds);  o.ln_s_b = take(Ds);
    o.ln_m_w = take(Ds);  o.ln_m_b = take(Ds);
    o.wq = take(Ds * Ds); o.wk = take(Ds * D); o.wv = take(Ds * D);
    o.wih = take(3 * Ds * Ds); o.whh = take(3 * Ds * Ds);
    o.bih = take(3 * Ds); o.bhh = take(3 * Ds);
    o.w1 = take(M * Ds); o.b1 = take(M); o.w2 = take(Ds * M); o.b2 = take(Ds);
    for (int j = 0; j < SAVI_MAX_BLOCKS; ++j) {
        BlockOff& b = o.blk[j];
        if (j < s.blocks) {
            b.ln1_w = take(Ds); b.ln1_b = take(Ds);
            b.pq = take(Ds * Ds); b.pk = take(Ds * Ds); b.pv = take(Ds * Ds); b.po = take(Ds * Ds);
            b.ln2_w = take(Ds); b.ln2_b = take(Ds);
            b.f1 = take(4 * Ds * Ds); b.f1b = take(4 * Ds);
            b.f2 = take(4 * Ds * Ds); b.f2b = take(Ds);
        } else {
            b = BlockOff{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        }
    }
    o.lnf_w = take(Ds); o.lnf_b = take(Ds);
    o.total = p;
    o.wq_t = take(Ds * Ds); o.wk_t = take(Ds * D); o.wv_t = take(Ds * D);
    o.wih_t = take(3 * Ds * Ds); o.whh_t = take(3 * Ds * Ds);
    o.w1_t = take(M * Ds); o.w2_t = take(Ds * M);
    for (int j = 0; j < SAVI_MAX_BLOCKS; ++j) {
        BlockOffT& b = o.blkt[j];
        if (j < s.blocks) {
            b.pq_t = take(Ds * Ds); b.pk_t = take(Ds * Ds); b.pv_t = take(Ds * Ds); b.po_t = take(Ds * Ds);
            b.f1_t = take(4 * Ds * Ds); b.f2_t = take(4 * Ds * Ds);
        } else {
            b = BlockOffT{0, 0, 0, 0, 0, 0};
        }
    }
    o.wqk = take(D * Ds); o.wg = take(3 * Ds * D); o.wqk_t = take(D * Ds); o.wg_t = take(3 * Ds * D);
    o.packed_total = p;
}

// ---------------------------------------------------------------------------
// saved-for-backward buffer.  Byte offsets of the token-side regions, float
// offsets (relative to `fbase`) of the slot-side field arrays.
// ---------------------------------------------------------------------------
struct Dims {
    int B, T, N, D, Ds, M, K, I, blocks, heads;
    int KP;          // K rounded up to a multiple of 4
    int CN;          // cluster size (CTAs per clip)
    int S;           // T*I           attention steps
    int Sm;          // T*(I-1)       steps that run the residual MLP
    int Sp;          // (T-1)*blocks  predictor block evaluations
    int tok_bytes;   // 4 (fp32) or 2 (bf16)
    int mma;         // 1: tensor-core path (bf16 tokens, D%16==0, N%8==0, ...), 0: SIMT path
    int KC;          // 16 * ceil(K/16): slot rows of the staged backward coefficients
    int umma;        // 1: tcgen05 clip kernels (bf16 tokens, D = Ds = M = 128, K <= 24, CN <= 2)
    int NTILE;       // ceil(N / 128): 128-token tiles per frame (tcgen05 path)
    float eps, ln_eps, qscale;
};

struct SavedLayout {
    int64_t xhat;        // bytes  [B,T,N,D] token dtype: LayerNorm'd tokens (steve.py:60)
    int64_t stats;       // bytes  [B,T,N,2] fp32: mean, rstd of each token
    int64_t fbase;       // bytes  start of the fp32 field arrays
    // per attention step (S steps), rows = S*B*K
    int64_t hp, q, qk, ux, u, r, z, n, ghn;      // widths Ds,Ds,D,D,Ds,Ds,Ds,Ds,Ds
    int64_t ssum;                                // [S*B, KP]  token sums of (P+eps)
    // per MLP step (Sm steps)
    int64_t hg, a;                               // widths Ds, M
    // per predictor block evaluation (Sp), rows = Sp*B*K
    int64_t px0;                                 // [(T-1)*B*K, Ds] predictor input of each frame (= slots_out[:, t])
    int64_t py, pq, pk, pv, po, px1, pl2, pf, px2;   // widths Ds x7, 4Ds, Ds
    int64_t patt;                                // [Sp*B, heads*K*K]
    int64_t ximg;        // bytes  [B*T][NTILE][D/64][128 rows][64] bf16, SWIZZLE_128B blocks of xhat (tcgen05 path only)
    // tcgen05 path only (float offsets like the fields above): LayerNorm outputs and statistics the backward reuses
    int64_t st, m;       // [S*B*K, Ds] s~ = LN_s(slots_prev);  [Sm*B*K, Ds] LN_m(h')
    int64_t lns, lnm;    // [S*B*K, 2], [Sm*B*K, 2]: (mean, rstd) of those LayerNorms
    int64_t total_bytes;
};

static inline void savi_saved_layout(const Dims& d, SavedLayout& L) {
    int64_t b = 0;
    // (the tcgen05 kernels read the LayerNorm'd tokens from the swizzled token image `ximg` only: no row-major copy is kept)
    L.xhat = b;  b += d.umma ? 0 : (int64_t)d.B * d.T * d.N * d.D * d.tok_bytes; b = (b + 255) / 256 * 256;
    L.stats = b; b += (int64_t)d.B * d.T * d.N * 2 * 4;             b = (b + 255) / 256 * 256;
    L.fbase = b;
    int64_t f = 0;
    auto take = [&](int64_t rows, int64_t w) { int64_t r = f; f += rows * w; f = (f + 3) / 4 * 4; return r; };
    const int64_t R = (int64_t)d.S * d.B * d.K, Rm = (int64_t)d.Sm * d.B * d.K, Rp = (int64_t)d.Sp * d.B * d.K;
    // (q and U are never formed on the tcgen05 path: project_q / project_k and project_v / gru.weight_ih are folded)
    L.hp = take(R, d.Ds); L.q = take(d.umma ? 0 : R, d.Ds); L.qk = take(R, d.D); L.ux = take(R, d.D); L.u = take(d.umma ? 0 : R, d.Ds);
    L.r = take(R, d.Ds); L.z = take(R, d.Ds); L.n = take(R, d.Ds); L.ghn = take(R, d.Ds);
    L.ssum = take((int64_t)d.S * d.B, d.KP);
    L.hg = take(Rm, d.Ds); L.a = take(Rm, d.M);
    L.px0 = take((int64_t)(d.T - 1) * d.B * d.K, d.Ds);
    L.py = take(Rp, d.Ds); L.pq = take(Rp, d.Ds); L.pk = take(Rp, d.Ds); L.pv = take(Rp, d.Ds); L.po = take(Rp, d.Ds);
    L.px1 = take(Rp, d.Ds); L.pl2 = take(Rp, d.Ds); L.pf = take(Rp, 4 * d.Ds); L.px2 = take(Rp, d.Ds);
    L.patt = take((int64_t)d.Sp * d.B, (int64_t)d.heads * d.K * d.K);
    L.st = L.m = L.lns = L.lnm = 0;
    if (d.umma) { L.st = take(R, d.Ds); L.m = take(Rm, d.Ds); L.lns = take(R, 2); L.lnm = take(Rm, 2); }
    int64_t e = L.fbase + f * 4;
    e = (e + 1023) / 1024 * 1024;
    L.ximg = e;
    if (d.umma) e += (int64_t)d.B * d.T * d.NTILE * (d.D / 64) * 16384;
    L.total_bytes = e;
}

// ---------------------------------------------------------------------------
// forward workspace: per-CTA scratch + the cross-CTA partial sums of one clip
// ---------------------------------------------------------------------------
struct FwdWsLayout {         // float offsets
    int64_t part;            // [B][2][CN][K*D + KP]   partial (sum A x, sum A) per cluster rank, double-buffered
    int64_t cta;             // [B*CN][cta_floats]     per-CTA scratch
    int64_t cta_floats;
    // inside one CTA's scratch:
    int64_t h, st, gi, gh, tmp;          // [K,Ds], [K,Ds], [K,3Ds], [K,3Ds], [K,max(D,Ds)]
    int64_t shadow;          // a shadow "step record" for cluster ranks > 0 (they recompute the slot state redundantly)
    int64_t sh_hp, sh_q, sh_qk, sh_ux, sh_u, sh_r, sh_z, sh_n, sh_ghn, sh_ssum, sh_hg, sh_a;
    int64_t sh_py, sh_pq, sh_pk, sh_pv, sh_po, sh_px1, sh_pl2, sh_pf, sh_px2, sh_patt;
    int64_t total_bytes;
};

static inline void savi_fwd_ws_layout(const Dims& d, FwdWsLayout& L) {
    int64_t f = 0;
    auto take = [&](int64_t n) { int64_t r = f; f += (n + 3) / 4 * 4; return r; };
    const int64_t KDs = (int64_t)d.K * d.Ds, KD = (int64_t)d.K * d.D;
    L.h = take(KDs); L.st = take(KDs); L.gi = take(3 * KDs); L.gh = take(3 * KDs);
    L.tmp = take((int64_t)d.K * (d.D > d.Ds ? d.D : d.Ds));
    L.shadow = f;
    L.sh_hp = take(KDs); L.sh_q = take(KDs); L.sh_qk = take(KD); L.sh_ux = take(KD); L.sh_u = take(KDs);
    L.sh_r = take(KDs); L.sh_z = take(KDs); L.sh_n = take(KDs); L.sh_ghn = take(KDs); L.sh_ssum = take(d.KP);
    L.sh_hg = take(KDs); L.sh_a = take((int64_t)d.K * d.M);
    L.sh_py = take(KDs); L.sh_pq = take(KDs); L.sh_pk = take(KDs); L.sh_pv = take(KDs); L.sh_po = take(KDs);
    L.sh_px1 = take(KDs); L.sh_pl2 = take(KDs); L.sh_pf = take(4 * KDs); L.sh_px2 = take(KDs);
    L.sh_patt = take((int64_t)d.heads * d.K * d.K);
    L.cta_floats = f;
    int64_t g = 0;
    L.part = g; g += (int64_t)d.B * 2 * d.CN * (KD + d.KP); g = (g + 3) / 4 * 4;
    L.cta = g;  g += (int64_t)d.B * d.CN * L.cta_floats;
    L.total_bytes = g * 4;
}

// ---------------------------------------------------------------------------
// backward workspace
// ---------------------------------------------------------------------------
struct BwdWsLayout {         // float offsets unless noted
    int64_t dxhat;           // [B,T,N,D] fp32 accumulator of d xhat over the I iterations of a frame
    // staged operands of the weight-gradient GEMMs (field arrays, same row order as SavedLayout)
    int64_t dq, st, dqk, du, dgi, dgh;           // S steps: widths Ds, Ds, D, Ds, 3Ds, 3Ds
    int64_t duxs;            // S steps, width D: d(Ux) of every step (right-hand side of the d_inputs kernel, mma mode)
    int64_t coef;            // BYTE offset: [B*T][I][2][KC][N] bf16 staged dL^T, W^T (mma mode)
    int64_t dwqk, dwg;       // tcgen05 path: gradients of the folded weights wqk [D,Ds], wg [3Ds,D] (contiguous; zeroed per call)
    int64_t flags;           // BYTE offset (tcgen05 path): int [B*T] "frame staged" counters, the clip kernel -> the overlapped d_inputs kernel
    int64_t dhm, da, m;                          // Sm steps: widths Ds, M, Ds
    int64_t pdq, pdk, pdv, pdx1, pdx2, pdf;      // Sp: widths Ds x5, 4Ds
    int64_t part;            // [B][2][CN][K*D]  partial d(qk) per cluster rank
    int64_t cta;             // [B*CN][cta_floats]
    int64_t cta_floats;
    // inside one CTA's scratch ([K, .] each)
    int64_t dh, dhg, t0, t1, t2, dux, cvec;      // Ds, Ds, max(4Ds,3Ds,M), same, same, D, KP
    int64_t sh;              // shadow staging for ranks > 0 (same fields as above, one step)
    int64_t sh_dq, sh_st, sh_dqk, sh_du, sh_dgi, sh_dgh, sh_dhm, sh_da, sh_m;
    int64_t sh_pdq, sh_pdk, sh_pdv, sh_pdx1, sh_pdx2, sh_pdf;
    int64_t total_bytes;
};

static inline void savi_bwd_ws_layout(const Dims& d, BwdWsLayout& L) {
    const int64_t KDs = (int64_t)d.K * d.Ds, KD = (int64_t)d.K * d.D;
    int wide = 4 * d.Ds; if (d.M > wide) wide = d.M; if (d.D > wide) wide = d.D; if (d.heads * d.K > wide) wide = d.heads * d.K;
    int64_t f = 0;
    auto take = [&](int64_t n) { int64_t r = f; f += (n + 3) / 4 * 4; return r; };
    L.dh = take(KDs); L.dhg = take(KDs);
    L.t0 = take((int64_t)d.K * wide); L.t1 = take((int64_t)d.K * wide); L.t2 = take((int64_t)d.K * wide);
    L.dux = take(KD); L.cvec = take(2 * d.KP);
    L.sh = f;
    L.sh_dq = take(KDs); L.sh_st = take(KDs); L.sh_dqk = take(KD); L.sh_du = take(KDs);
    L.sh_dgi = take(3 * KDs); L.sh_dgh = take(3 * KDs);
    L.sh_dhm = take(KDs); L.sh_da = take((int64_t)d.K * d.M); L.sh_m = take(KDs);
    L.sh_pdq = take(KDs); L.sh_pdk = take(KDs); L.sh_pdv = take(KDs); L.sh_pdx1 = take(KDs); L.sh_pdx2 = take(KDs);
    L.sh_pdf = take(4 * KDs);
    L.cta_floats = f;

    int64_t g = 0;
    auto gt = [&](int64_t rows, int64_t w) { int64_t r = g; g += rows * w; g = (g + 3) / 4 * 4; return r; };
    const int64_t R = (int64_t)d.S * d.B * d.K, Rm = (int64_t)d.Sm * d.B * d.K, Rp = (int64_t)d.Sp * d.B * d.K;
    L.dxhat = gt(d.mma ? 0 : (int64_t)d.B * d.T * d.N, d.D);
    L.dq = gt(R, d.Ds); L.st = gt(R, d.Ds); L.dqk = gt(R, d.D); L.du = gt(R, d.Ds);
    L.dgi = gt(R, 3 * d.Ds); L.dgh = gt(R, 3 * d.Ds);
    L.dhm = gt(Rm, d.Ds); L.da = gt(Rm, d.M); L.m = gt(Rm, d.Ds);
    L.pdq = gt(Rp, d.Ds); L.pdk = gt(Rp, d.Ds); L.pdv = gt(Rp, d.Ds); L.pdx1 = gt(Rp, d.Ds); L.pdx2 = gt(Rp, d.Ds);
    L.pdf = gt(Rp, 4 * d.Ds);
    L.duxs = gt(d.mma ? R : 0, d.D);
    L.part = gt((int64_t)d.B * 2 * d.CN, KD);
    L.dwqk = gt(d.umma ? d.D : 0, d.Ds); L.dwg = gt(d.umma ? 3 * d.Ds : 0, d.D);
    L.cta = g; g += (int64_t)d.B * d.CN * L.cta_floats;
    g = (g + 63) / 64 * 64;
    L.coef = g * 4;
    // tcgen05 path: one 16 KB K-major operand block [128 tokens][32 dL | 32 W] per (frame, tile, iteration)
    if (d.umma) L.total_bytes = (g * 4 + 1023) / 1024 * 1024 + (int64_t)d.B * d.T * d.NTILE * d.I * 16384;
    else L.total_bytes = g * 4 + (d.mma ? (int64_t)d.B * d.T * d.I * 2 * d.KC * d.N * 2 : 0);
    if (d.umma) L.coef = (g * 4 + 1023) / 1024 * 1024;
    L.flags = L.total_bytes;
    if (d.umma) L.total_bytes += (((int64_t)d.B * d.T + 1) * 4 + 255) / 256 * 256;     // + 1: "clip kernel finished" CTA counter
}
