// Backward kernels: per-clip BPTT kernel, weight-gradient GEMM, token LayerNorm backward.
//
// The reference has no explicit backward (autograd through steve.py:52-105); the
// closed form implemented here is SURVEY.md Appendix A.2 in the folded form and is
// restated and validated against the reference's autograd in oracle/savi_numpy.py.
#include "savi_dx_mma.cuh"
#include "savi_args.h"

// ---------------------------------------------------------------------------
// Attention-step backward over this CTA's tokens of frame (b,t).
//   recompute  L = xhat qk^T, P = softmax(L), W = (P+eps)/S
//   G  = xhat dUx^T ; dP = (G - c)/S (+ grad_attn) ; dL = P (dP - <P,dP>)
//   part (d qk) = sum_n dL[n,:]^T xhat[n,:]
//   dxhat[n,:] (+)= dL[n,:] qk + W[n,:] dUx
// ---------------------------------------------------------------------------
template <typename TokT, int KMAX>
__device__ __noinline__ void token_pass_bwd(const Dims& d, const TokT* __restrict__ xh, int n_lo, int n_hi, const float* qk_g,
                               const float* dux_g, const float* cvec_g, const TokT* __restrict__ gattn,
                               float* dxhat_g, bool accumulate, float* part_g, unsigned char* smem, int TN) {
    const int tid = threadIdx.x, K = d.K, KP = d.KP, D = d.D;
    const int cst = coef_stride(KP), xst = tile_stride_bytes(D, sizeof(TokT));
    float* qk_s = reinterpret_cast<float*>(smem);                  // [KMAX][D]
    float* dux_s = qk_s + (size_t)KMAX * D;                        // [KMAX][D]
    float* acc_s = dux_s + (size_t)KMAX * D;                       // [KP][D]
    float* cv = acc_s + (size_t)KP * D;                            // [2*KP]: c, 1/S
    float* dl = cv + 2 * KP;                                       // [TN][cst]
    float* wt = dl + (size_t)TN * cst;                             // [TN][cst]
    unsigned char* xs = reinterpret_cast<unsigned char*>(wt + (size_t)TN * cst);

    __syncthreads();
    for (int i = tid * 4; i < KMAX * D; i += NT * 4) {
        const bool in = i < K * D;
        st4(qk_s + i, in ? ld4(qk_g + i) : make_float4(0.f, 0.f, 0.f, 0.f));
        st4(dux_s + i, in ? ld4(dux_g + i) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    for (int i = tid; i < KP * D; i += NT) acc_s[i] = 0.f;
    for (int i = tid; i < 2 * KP; i += NT) cv[i] = cvec_g[i];

    for (int n0 = n_lo; n0 < n_hi; n0 += TN) {
        const int tn = min(TN, n_hi - n0);
        __syncthreads();
        load_token_tile<TokT>(xs, xst, xh, n0, tn, D);
        __syncthreads();
        for (int n = tid; n < tn; n += NT) {
            constexpr int VEC = Tok<TokT>::VEC;
            const TokT* xr = reinterpret_cast<const TokT*>(xs + (size_t)n * xst);
            float* wr = wt + (size_t)n * cst;
            float* lr = dl + (size_t)n * cst;
            float acc[KMAX];
            // pass 1: logits -> P (kept in the wt row)
#pragma unroll
            for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
            for (int c = 0; c < D; c += VEC) {
                float xv[VEC];
                Tok<TokT>::load(xr + c, xv);
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
#pragma unroll
                    for (int e = 0; e < VEC; e += 4) {
                        const float4 w = ld4(qk_s + (size_t)k * D + c + e);
                        acc[k] = fmaf(xv[e], w.x, acc[k]); acc[k] = fmaf(xv[e + 1], w.y, acc[k]);
                        acc[k] = fmaf(xv[e + 2], w.z, acc[k]); acc[k] = fmaf(xv[e + 3], w.w, acc[k]);
                    }
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) mx = fmaxf(mx, acc[k]);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) { acc[k] = (k < K) ? expf(acc[k] - mx) : 0.f; sum += acc[k]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k < KMAX; k += 4)
                if (k < KP) st4(wr + k, make_float4(acc[k] * inv, acc[k + 1] * inv, acc[k + 2] * inv, acc[k + 3] * inv));
            // pass 2: G = x . dUx
#pragma unroll
            for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
            for (int c = 0; c < D; c += VEC) {
                float xv[VEC];
                Tok<TokT>::load(xr + c, xv);
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
#pragma unroll
                    for (int e = 0; e < VEC; e += 4) {
                        const float4 w = ld4(dux_s + (size_t)k * D + c + e);
                        acc[k] = fmaf(xv[e], w.x, acc[k]); acc[k] = fmaf(xv[e + 1], w.y, acc[k]);
                        acc[k] = fmaf(xv[e + 2], w.z, acc[k]); acc[k] = fmaf(xv[e + 3], w.w, acc[k]);
                    }
                }
            }
            const TokT* gr = gattn ? gattn + (size_t)(n0 + n) * K : nullptr;
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    float dp = (acc[k] - cv[k]) * cv[KP + k];
                    if (gr) dp += Tok<TokT>::to_f(gr[k]);
                    acc[k] = dp;
                    dot = fmaf(wr[k], dp, dot);
                }
            }
#pragma unroll
            for (int k = 0; k < KMAX; k += 4) {
                if (k < KP) {
                    const float4 p = ld4(wr + k);
                    float4 l, w;
                    l.x = (k + 0 < K) ? p.x * (acc[k + 0] - dot) : 0.f;  w.x = (k + 0 < K) ? (p.x + d.eps) * cv[KP + k + 0] : 0.f;
                    l.y = (k + 1 < K) ? p.y * (acc[k + 1] - dot) : 0.f;  w.y = (k + 1 < K) ? (p.y + d.eps) * cv[KP + k + 1] : 0.f;
                    l.z = (k + 2 < K) ? p.z * (acc[k + 2] - dot) : 0.f;  w.z = (k + 2 < K) ? (p.z + d.eps) * cv[KP + k + 2] : 0.f;
                    l.w = (k + 3 < K) ? p.w * (acc[k + 3] - dot) : 0.f;  w.w = (k + 3 < K) ? (p.w + d.eps) * cv[KP + k + 3] : 0.f;
                    st4(lr + k, l);
                    st4(wr + k, w);
                }
            }
        }
        __syncthreads();
        // d qk partial
        tile_outer_accum<TokT>(acc_s, nullptr, dl, cst, xs, xst, tn, KP, D);
        // d xhat rows of this tile: 4 tokens x 4 features per thread, lanes along the feature dim (coalesced RMW)
        const int dgn = D >> 2, tgn = (tn + 3) >> 2;
        for (int item = tid; item < tgn * dgn; item += NT) {
            const int tg = item / dgn, dg = item - tg * dgn, nb = tg * 4;
            float4 o[4];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) o[tt] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k4 = 0; k4 < KP; k4 += 4) {
                float4 q[4], u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    q[j] = ld4(qk_s + (size_t)(k4 + j) * D + dg * 4);
                    u[j] = ld4(dux_s + (size_t)(k4 + j) * D + dg * 4);
                }
#pragma unroll
                for (int tt = 0; tt < 4; ++tt) {
                    const float4 l = ld4(dl + (size_t)(nb + tt) * cst + k4);
                    const float4 w = ld4(wt + (size_t)(nb + tt) * cst + k4);
                    const float lv[4] = {l.x, l.y, l.z, l.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[tt].x = fmaf(lv[j], q[j].x, o[tt].x); o[tt].y = fmaf(lv[j], q[j].y, o[tt].y);
                        o[tt].z = fmaf(lv[j], q[j].z, o[tt].z); o[tt].w = fmaf(lv[j], q[j].w, o[tt].w);
                        o[tt].x = fmaf(wv[j], u[j].x, o[tt].x); o[tt].y = fmaf(wv[j], u[j].y, o[tt].y);
                        o[tt].z = fmaf(wv[j], u[j].z, o[tt].z); o[tt].w = fmaf(wv[j], u[j].w, o[tt].w);
                    }
                }
            }
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
                if (nb + tt < tn) {
                    float* p = dxhat_g + (size_t)(n0 + nb + tt) * D + dg * 4;
                    float4 v = o[tt];
                    if (accumulate) { const float4 t = ld4(p); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
                    st4(p, v);
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < K * D; i += NT) part_g[i] = acc_s[i];
}

// Backward of the predictor's attention core.  dQ is w.r.t. the UNSCALED projection.
// m (nullable): the forward's dropout mask of the probabilities: O = (att . m) V, so dV uses att . m and d att = (dO V^T) . m.
static __device__ __noinline__ void mha_core_bwd(const float* dO, const float* Q, const float* Kk, const float* V, const float* att,
                             float* datt, float* dQ, float* dKk, float* dV, int K, int Ds, int H, float hscale, const float* m) {
    const int dh = Ds / H, tid = threadIdx.x;
    for (int idx = tid; idx < H * K * K; idx += NT) {
        const int j = idx % K, i = (idx / K) % K, h = idx / (K * K);
        const float* a = dO + (size_t)i * Ds + h * dh;
        const float* v = V + (size_t)j * Ds + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s = fmaf(a[c], v[c], s);
        datt[idx] = m ? s * m[idx] : s;
    }
    __syncthreads();
    for (int row = tid; row < H * K; row += NT) {
        const float* a = att + (size_t)row * K;
        float* da = datt + (size_t)row * K;
        float dot = 0.f;
        for (int j = 0; j < K; ++j) dot = fmaf(a[j], da[j], dot);
        for (int j = 0; j < K; ++j) da[j] = a[j] * (da[j] - dot);
    }
    __syncthreads();
    for (int idx = tid; idx < K * Ds; idx += NT) {
        const int c = idx % Ds, i = idx / Ds, h = c / dh;
        const float* dlg = datt + (size_t)h * K * K;
        const float* at = att + (size_t)h * K * K;
        float sq = 0.f, sk = 0.f, sv = 0.f;
        for (int j = 0; j < K; ++j) {
            sq = fmaf(dlg[(size_t)i * K + j], Kk[(size_t)j * Ds + c], sq);
            sk = fmaf(dlg[(size_t)j * K + i], Q[(size_t)j * Ds + c], sk);
            sv = fmaf(m ? at[(size_t)j * K + i] * m[(size_t)h * K * K + (size_t)j * K + i] : at[(size_t)j * K + i], dO[(size_t)j * Ds + c], sv);
        }
        dQ[idx] = sq * hscale; dKk[idx] = sk; dV[idx] = sv;
    }
    __syncthreads();
}

// Shared-memory version of mha_core_bwd (all five operands staged, odd row strides).
static __device__ __noinline__ void mha_core_bwd_smem(const float* dO, const float* Q, const float* Kk, const float* V, const float* att,
                                                      float* dQ, float* dKk, float* dV, int K, int Ds, int H, float hscale, float* arena, const float* m) {
    const int dh = Ds / H, tid = threadIdx.x, ld = Ds + 1, ka = K | 1;
    float* sO = arena; float* sQ = sO + (size_t)K * ld; float* sK = sQ + (size_t)K * ld; float* sV = sK + (size_t)K * ld;
    float* sA = sV + (size_t)K * ld; float* sD = sA + (size_t)H * K * ka;
    __syncthreads();
    for (int i = tid * 4; i < K * Ds; i += NT * 4) {
        const int r = i / Ds, c = i - r * Ds;
        const float4 o = ld4(dO + i), a = ld4(Q + i), b = ld4(Kk + i), v = ld4(V + i);
        float* p0 = sO + (size_t)r * ld + c; float* p1 = sQ + (size_t)r * ld + c; float* p2 = sK + (size_t)r * ld + c; float* p3 = sV + (size_t)r * ld + c;
        p0[0] = o.x; p0[1] = o.y; p0[2] = o.z; p0[3] = o.w;
        p1[0] = a.x; p1[1] = a.y; p1[2] = a.z; p1[3] = a.w;
        p2[0] = b.x; p2[1] = b.y; p2[2] = b.z; p2[3] = b.w;
        p3[0] = v.x; p3[1] = v.y; p3[2] = v.z; p3[3] = v.w;
    }
    for (int idx = tid; idx < H * K * K; idx += NT) sA[(size_t)(idx / K) * ka + idx % K] = att[idx];
    __syncthreads();
    for (int idx = tid; idx < H * K * K; idx += NT) {
        const int j = idx % K, i = (idx / K) % K, h = idx / (K * K);
        const float* a = sO + (size_t)i * ld + h * dh;
        const float* v = sV + (size_t)j * ld + h * dh;
        float s = 0.f;
        for (int c = 0; c < dh; ++c) s = fmaf(a[c], v[c], s);
        sD[(size_t)(h * K + i) * ka + j] = m ? s * m[idx] : s;
    }
    __syncthreads();
    for (int row = tid; row < H * K; row += NT) {
        const float* a = sA + (size_t)row * ka;
        float* da = sD + (size_t)row * ka;
        float dot = 0.f;
        for (int j = 0; j < K; ++j) dot = fmaf(a[j], da[j], dot);
        for (int j = 0; j < K; ++j) da[j] = a[j] * (da[j] - dot);
    }
    __syncthreads();
    if (m) {                                                  // d V = (att . m)^T dO: the probabilities are not needed undropped any more
        for (int idx = tid; idx < H * K * K; idx += NT) sA[(size_t)(idx / K) * ka + idx % K] *= m[idx];
        __syncthreads();
    }
    for (int idx = tid; idx < K * Ds; idx += NT) {
        const int c = idx % Ds, i = idx / Ds, h = c / dh;
        const float* dlg = sD + (size_t)h * K * ka;
        const float* at = sA + (size_t)h * K * ka;
        float sq = 0.f, sk = 0.f, sv = 0.f;
        for (int j = 0; j < K; ++j) {
            sq = fmaf(dlg[(size_t)i * ka + j], sK[(size_t)j * ld + c], sq);
            sk = fmaf(dlg[(size_t)j * ka + i], sQ[(size_t)j * ld + c], sk);
            sv = fmaf(at[(size_t)j * ka + i], sO[(size_t)j * ld + c], sv);
        }
        dQ[idx] = sq * hscale; dKk[idx] = sk; dV[idx] = sv;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// K3: BPTT through the whole recurrence of one clip (frames T-1..0, iterations I-1..0).
// Writes: dxhat accumulator, the staged (dY, X) operands of the weight-gradient
// GEMMs, vector-parameter gradients (atomics into the flat gradient buffer).
// ---------------------------------------------------------------------------
template <typename TokT, int KMAX, bool MMA>
__global__ void __launch_bounds__(NT, 1) savi_bwd_kernel(const __grid_constant__ BwdArgs a) {
    extern __shared__ float4 smem4[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(smem4);
    float* arena = reinterpret_cast<float*>(smem4);
    const Dims& d = a.d;
    const ParamOff& po = a.po;
    const int tid = threadIdx.x;
    const int CN = d.CN, b = blockIdx.x / CN, rank = blockIdx.x % CN;
    const int K = d.K, Ds = d.Ds, D = d.D, M = d.M, B = d.B, KP = d.KP;
    const int per = ((d.N + CN - 1) / CN + 15) & ~15;
    const int n_lo = min(d.N, rank * per), n_hi = min(d.N, n_lo + per);
    const float* P = a.packed;
    float* G = a.grad_params;
    float* fb = const_cast<float*>(reinterpret_cast<const float*>(a.saved + a.sl.fbase));   // read-only use
    const TokT* xhat = reinterpret_cast<const TokT*>(a.saved + a.sl.xhat);
    float* W = a.ws;
    float* cs = W + a.wl.cta + (size_t)blockIdx.x * a.wl.cta_floats;
    float* dh = cs + a.wl.dh;
    float* dhg = cs + a.wl.dhg;
    float* t0 = cs + a.wl.t0;
    float* t1 = cs + a.wl.t1;
    float* t2 = cs + a.wl.t2;
    float* dux = cs + a.wl.dux;     // mma mode: redirected per step into the staged field array
    float* cvec = cs + a.wl.cvec;
    const bool lead = (rank == 0);
    const int AF = a.arena_floats;
    const float hscale = 1.0f / sqrtf((float)(Ds / d.heads));
    constexpr int MT = (KMAX + 15) / 16;
    const bf16* Phi = reinterpret_cast<const bf16*>(P + po.packed_total);
    const bf16* Plo = Phi + po.packed_total;
    // backward linears dX = dY . W contract over the OUT index: SIMT uses the original [out][in] copy as its
    // "[C][O]" operand, the tensor-core path uses the transposed copy ([in][out] = "[O'][C']" with C' = out)
#define LIN(Y, ldy, X, ldx, W_io, W_oi, Res, ldr, Mask, ldm, R_, C_, O_, alpha) \
    lin<MMA, MT>(P, Phi, Plo, Y, ldy, X, ldx, W_io, W_oi, nullptr, Res, ldr, Mask, ldm, R_, C_, O_, alpha, 0, arena, AF)

    long long ph_last = clock64();
    for (int i = tid; i < K * Ds; i += NT) dh[i] = 0.f;
    __syncthreads();

    for (int t = d.T - 1; t >= 0; --t) {
        const TokT* xh_t = xhat + ((size_t)b * d.T + t) * d.N * D;
        float* dxh_t = W + a.wl.dxhat + ((size_t)b * d.T + t) * d.N * D;
        if (t < d.T - 1) {
            SAVI_PH(30);
            // ---- predictor backward (transformer.py:106-114, 70-86, 22-49) ----
            const float* px0 = fb + a.sl.px0 + ((size_t)t * B + b) * K * Ds;
            const float* x_last = d.blocks > 0 ? frow(fb, a.sl.px2, (int64_t)(d.blocks - 1) * (d.T - 1) + t, b, B, K, Ds) : px0;
            cta_ln_bwd(t0, Ds, nullptr, 0, dh, Ds, x_last, Ds, P + po.lnf_w, G + po.lnf_w, G + po.lnf_b, K, Ds, d.ln_eps, lead);
            for (int j = d.blocks - 1; j >= 0; --j) {
                const int64_t f = (int64_t)j * (d.T - 1) + t;
                const BlockOff& bo = po.blk[j];
                const BlockOffT& bt = po.blkt[j];
                const float* p_y = frow(fb, a.sl.py, f, b, B, K, Ds);
                const float* p_q = frow(fb, a.sl.pq, f, b, B, K, Ds);
                const float* p_k = frow(fb, a.sl.pk, f, b, B, K, Ds);
                const float* p_v = frow(fb, a.sl.pv, f, b, B, K, Ds);
                const float* p_x1 = frow(fb, a.sl.px1, f, b, B, K, Ds);
                const float* p_f = frow(fb, a.sl.pf, f, b, B, K, 4 * Ds);
                const float* p_att = fb + a.sl.patt + (f * B + b) * ((int64_t)d.heads * K * K);
                float* s_dq = lead ? frow(W, a.wl.pdq, f, b, B, K, Ds) : cs + a.wl.sh_pdq;
                float* s_dk = lead ? frow(W, a.wl.pdk, f, b, B, K, Ds) : cs + a.wl.sh_pdk;
                float* s_dv = lead ? frow(W, a.wl.pdv, f, b, B, K, Ds) : cs + a.wl.sh_pdv;
                float* s_dx1 = lead ? frow(W, a.wl.pdx1, f, b, B, K, Ds) : cs + a.wl.sh_pdx1;
                float* s_dx2 = lead ? frow(W, a.wl.pdx2, f, b, B, K, Ds) : cs + a.wl.sh_pdx2;
                float* s_df = lead ? frow(W, a.wl.pdf, f, b, B, K, 4 * Ds) : cs + a.wl.sh_pdf;
                (void)p_y;
                // training-mode dropout masks of this block evaluation (nullptr otherwise): savi_args.h, DropLayout
                const DropLayout dl = savi_dropout_layout(d);
                const float* m_att = a.drop ? a.drop + dl.att + (f * B + b) * ((int64_t)d.heads * K * K) : nullptr;
                const float* m_out = a.drop ? a.drop + dl.out + (f * B + b) * ((int64_t)K * Ds) : nullptr;
                const float* m_ffn = a.drop ? a.drop + dl.ffn + (f * B + b) * ((int64_t)K * Ds) : nullptr;
                // d(ffn.2 output) = d x2 . m_ffn: staged for d W2 / d b2 and pushed through ffn.2^T; the residual branch keeps d x2 (t0)
                if (m_ffn) { for (int i = tid; i < K * Ds; i += NT) s_dx2[i] = t0[i] * m_ffn[i]; __syncthreads(); }
                else cta_copy(s_dx2, t0, K * Ds);
                const float* dff = m_ffn ? s_dx2 : t0;       // (without dropout: the CTA-local copy, already synchronised)
                if (lead) cta_colsum_atomic(G + bo.f2b, dff, Ds, K, Ds);
                LIN(s_df, 4 * Ds, dff, Ds, bo.f2, bt.f2_t, nullptr, 0, p_f, 4 * Ds, K, Ds, 4 * Ds, 1.0f);      // d relu-out, masked
                if (lead) cta_colsum_atomic(G + bo.f1b, s_df, 4 * Ds, K, 4 * Ds);
                LIN(t1, Ds, s_df, 4 * Ds, bo.f1, bt.f1_t, nullptr, 0, nullptr, 0, K, 4 * Ds, Ds, 1.0f);
                cta_ln_bwd(s_dx1, Ds, t0, Ds, t1, Ds, p_x1, Ds, P + bo.ln2_w, G + bo.ln2_w, G + bo.ln2_b, K, Ds, d.ln_eps, lead);
                // s_dx1 = d x1.  With output dropout the proj_o branch sees d x1 . m_out: THAT is staged (d W_o) and pushed through
                // proj_o^T, while the residual branch below needs the undropped d x1: kept in the CTA scratch `dhg` (idle here)
                const float* dx1_res = s_dx1;
                if (m_out) {
                    __syncthreads();
                    for (int i = tid; i < K * Ds; i += NT) { const float v = s_dx1[i]; dhg[i] = v; s_dx1[i] = v * m_out[i]; }
                    __syncthreads();
                    dx1_res = dhg;
                }
                LIN(t1, Ds, s_dx1, Ds, bo.po, bt.po_t, nullptr, 0, nullptr, 0, K, Ds, Ds, 1.0f);                  // dO
                SAVI_PH(31);
                if (4 * K * (Ds + 1) + 2 * d.heads * K * (K | 1) <= AF) mha_core_bwd_smem(t1, p_q, p_k, p_v, p_att, s_dq, s_dk, s_dv, K, Ds, d.heads, hscale, arena, m_att);
                else mha_core_bwd(t1, p_q, p_k, p_v, p_att, t2, s_dq, s_dk, s_dv, K, Ds, d.heads, hscale, m_att);
                SAVI_PH(32);
                LIN(t1, Ds, s_dq, Ds, bo.pq, bt.pq_t, (j == 0) ? dx1_res : nullptr, Ds, nullptr, 0, K, Ds, Ds, 1.0f);
                LIN(t1, Ds, s_dk, Ds, bo.pk, bt.pk_t, t1, Ds, nullptr, 0, K, Ds, Ds, 1.0f);
                LIN(t1, Ds, s_dv, Ds, bo.pv, bt.pv_t, t1, Ds, nullptr, 0, K, Ds, Ds, 1.0f);                      // dy
                const float* x_in = (j == 0) ? px0 : frow(fb, a.sl.px2, (int64_t)(j - 1) * (d.T - 1) + t, b, B, K, Ds);
                cta_ln_bwd(t0, Ds, (j == 0) ? nullptr : dx1_res, Ds, t1, Ds, x_in, Ds, P + bo.ln1_w, G + bo.ln1_w, G + bo.ln1_b,
                           K, Ds, d.ln_eps, lead);
            }
            cta_copy(dh, t0, K * Ds);
            __syncthreads();
            SAVI_PH(33);
        }
        for (int i = tid; i < K * Ds; i += NT) dh[i] += a.grad_slots[((size_t)b * d.T + t) * K * Ds + i];
        __syncthreads();

        for (int it = d.I - 1; it >= 0; --it) {
            const int64_t s = (int64_t)t * d.I + it;
            const float* r_hp = frow(fb, a.sl.hp, s, b, B, K, Ds);
            const float* r_q = frow(fb, a.sl.q, s, b, B, K, Ds);
            const float* r_qk = frow(fb, a.sl.qk, s, b, B, K, D);
            const float* r_ux = frow(fb, a.sl.ux, s, b, B, K, D);
            const float* r_r = frow(fb, a.sl.r, s, b, B, K, Ds);
            const float* r_z = frow(fb, a.sl.z, s, b, B, K, Ds);
            const float* r_n = frow(fb, a.sl.n, s, b, B, K, Ds);
            const float* r_ghn = frow(fb, a.sl.ghn, s, b, B, K, Ds);
            const float* r_ss = fb + a.sl.ssum + (s * B + b) * KP;
            float* s_dq = lead ? frow(W, a.wl.dq, s, b, B, K, Ds) : cs + a.wl.sh_dq;
            float* s_st = lead ? frow(W, a.wl.st, s, b, B, K, Ds) : cs + a.wl.sh_st;
            float* s_dqk = lead ? frow(W, a.wl.dqk, s, b, B, K, D) : cs + a.wl.sh_dqk;
            float* s_du = lead ? frow(W, a.wl.du, s, b, B, K, Ds) : cs + a.wl.sh_du;
            float* s_dgi = lead ? frow(W, a.wl.dgi, s, b, B, K, 3 * Ds) : cs + a.wl.sh_dgi;
            float* s_dgh = lead ? frow(W, a.wl.dgh, s, b, B, K, 3 * Ds) : cs + a.wl.sh_dgh;
            (void)r_q;
            SAVI_PH(34);
            const float* cur = dh;
            if (it < d.I - 1) {
                // ---- residual MLP backward (steve.py:92-93) ----
                const int64_t sm = (int64_t)t * (d.I - 1) + it;
                const float* r_hg = frow(fb, a.sl.hg, sm, b, B, K, Ds);
                const float* r_a = frow(fb, a.sl.a, sm, b, B, K, M);
                float* s_dhm = lead ? frow(W, a.wl.dhm, sm, b, B, K, Ds) : cs + a.wl.sh_dhm;
                float* s_da = lead ? frow(W, a.wl.da, sm, b, B, K, M) : cs + a.wl.sh_da;
                float* s_m = lead ? frow(W, a.wl.m, sm, b, B, K, Ds) : cs + a.wl.sh_m;
                cta_copy(s_dhm, dh, K * Ds);
                if (lead) cta_colsum_atomic(G + po.b2, dh, Ds, K, Ds);
                LIN(s_da, M, dh, Ds, po.w2, po.w2_t, nullptr, 0, r_a, M, K, Ds, M, 1.0f);
                if (lead) cta_colsum_atomic(G + po.b1, s_da, M, K, M);
                cta_ln(s_m, Ds, r_hg, Ds, P + po.ln_m_w, P + po.ln_m_b, K, Ds, d.ln_eps);
                LIN(t0, Ds, s_da, M, po.w1, po.w1_t, nullptr, 0, nullptr, 0, K, M, Ds, 1.0f);
                cta_ln_bwd(dhg, Ds, dh, Ds, t0, Ds, r_hg, Ds, P + po.ln_m_w, G + po.ln_m_w, G + po.ln_m_b, K, Ds, d.ln_eps, lead);
                cur = dhg;
            }
            SAVI_PH(35);
            // ---- GRUCell backward (steve.py:87-89) ----
            for (int i = tid * 4; i < K * Ds; i += NT * 4) {
                const int k = i / Ds, c = i - k * Ds;
                const float4 g4 = ld4(cur + i), r4 = ld4(r_r + i), z4 = ld4(r_z + i), n4 = ld4(r_n + i), gn4 = ld4(r_ghn + i), hp4 = ld4(r_hp + i);
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
                const float nn[4] = {n4.x, n4.y, n4.z, n4.w}, gn[4] = {gn4.x, gn4.y, gn4.z, gn4.w}, hh[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
                float dr[4], dz[4], dn[4], dnr[4], dhn[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dn[e] = gg[e] * (1.0f - zz[e]) * (1.0f - nn[e] * nn[e]);
                    dz[e] = gg[e] * (hh[e] - nn[e]) * zz[e] * (1.0f - zz[e]);
                    dr[e] = dn[e] * gn[e] * rr[e] * (1.0f - rr[e]);
                    dnr[e] = dn[e] * rr[e];
                    dhn[e] = gg[e] * zz[e];
                }
                float* gi = s_dgi + (size_t)k * 3 * Ds + c; float* gh = s_dgh + (size_t)k * 3 * Ds + c;
                st4(gi, make_float4(dr[0], dr[1], dr[2], dr[3])); st4(gi + Ds, make_float4(dz[0], dz[1], dz[2], dz[3]));
                st4(gi + 2 * Ds, make_float4(dn[0], dn[1], dn[2], dn[3]));
                st4(gh, make_float4(dr[0], dr[1], dr[2], dr[3])); st4(gh + Ds, make_float4(dz[0], dz[1], dz[2], dz[3]));
                st4(gh + 2 * Ds, make_float4(dnr[0], dnr[1], dnr[2], dnr[3]));
                st4(dh + i, make_float4(dhn[0], dhn[1], dhn[2], dhn[3]));
            }
            __syncthreads();
            if (lead) { cta_colsum_atomic(G + po.bih, s_dgi, 3 * Ds, K, 3 * Ds); cta_colsum_atomic(G + po.bhh, s_dgh, 3 * Ds, K, 3 * Ds); }
            SAVI_PH(36);
            LIN(dh, Ds, s_dgh, 3 * Ds, po.whh, po.whh_t, dh, Ds, nullptr, 0, K, 3 * Ds, Ds, 1.0f);
            SAVI_PH(37);
            LIN(s_du, Ds, s_dgi, 3 * Ds, po.wih, po.wih_t, nullptr, 0, nullptr, 0, K, 3 * Ds, Ds, 1.0f);
            SAVI_PH(38);
            // ---- attention step backward ----
            if (MMA) dux = lead ? frow(W, a.wl.duxs, s, b, B, K, D) : cs + a.wl.dux;
            LIN(dux, D, s_du, Ds, po.wv, po.wv_t, nullptr, 0, nullptr, 0, K, Ds, D, 1.0f);
            {
                const int warp = tid >> 5, lane = tid & 31;
                for (int k = warp; k < KP; k += NW) {
                    float sdot = 0.f;
                    if (k < K) for (int c = lane; c < D; c += 32) sdot = fmaf(dux[(size_t)k * D + c], r_ux[(size_t)k * D + c], sdot);
                    sdot = warp_sum(sdot);
                    if (lane == 0) { cvec[k] = sdot; cvec[KP + k] = (k < K) ? 1.0f / r_ss[k] : 0.f; }
                }
                __syncthreads();
            }
            SAVI_PH(39);
            float* part = W + a.wl.part + (((size_t)b * 2 + (s & 1)) * CN + rank) * ((size_t)K * D);
            const TokT* ga = (a.grad_attn && it == d.I - 1)
                                 ? reinterpret_cast<const TokT*>(a.grad_attn) + ((size_t)b * d.T + t) * d.N * K : nullptr;
            if constexpr (MMA) {
                bf16* coef = reinterpret_cast<bf16*>(reinterpret_cast<unsigned char*>(W) + a.wl.coef) +
                             ((((size_t)b * d.T + t) * d.I + it) * 2 * d.KC) * d.N;
                token_pass_bwd_mma<MT>(d, xh_t, n_lo, n_hi, r_qk, dux, cvec, ga, coef, part, smem, a.stages);
            } else {
                token_pass_bwd<TokT, KMAX>(d, xh_t, n_lo, n_hi, r_qk, dux, cvec, ga, dxh_t, it != d.I - 1, part, smem, a.TN);
            }
            SAVI_PH(40);
            __threadfence();
            sync_clip(CN);
            SAVI_PH(41);
            {
                const float* pb = W + a.wl.part + (((size_t)b * 2 + (s & 1)) * CN) * ((size_t)K * D);
                for (int i = tid * 4; i < K * D; i += NT * 4) {
                    float4 v = __ldcg(reinterpret_cast<const float4*>(pb + i));
                    for (int r = 1; r < CN; ++r) {
                        const float4 t = __ldcg(reinterpret_cast<const float4*>(pb + (size_t)r * K * D + i));
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    st4(s_dqk + i, v);
                }
                __syncthreads();
            }
            SAVI_PH(42);
            LIN(s_dq, Ds, s_dqk, D, po.wk_t, po.wk, nullptr, 0, nullptr, 0, K, D, Ds, d.qscale);
            SAVI_PH(43);
            cta_ln(s_st, Ds, r_hp, Ds, P + po.ln_s_w, P + po.ln_s_b, K, Ds, d.ln_eps);
            SAVI_PH(44);
            LIN(t0, Ds, s_dq, Ds, po.wq, po.wq_t, nullptr, 0, nullptr, 0, K, Ds, Ds, 1.0f);
            SAVI_PH(45);
            cta_ln_bwd(dh, Ds, dh, Ds, t0, Ds, r_hp, Ds, P + po.ln_s_w, G + po.ln_s_w, G + po.ln_s_b, K, Ds, d.ln_eps, lead);
        }
    }
    // ---- slot initialisation backward (steve.py:56-57) ----
    if (lead) {
        for (int c = tid; c < Ds; c += NT) {
            const float es = expf(P[po.slot_log_sigma + c]);
            float smu = 0.f, sls = 0.f;
            for (int k = 0; k < K; ++k) {
                const float g = dh[(size_t)k * Ds + c];
                const float nz = a.noise[((size_t)b * K + k) * Ds + c];
                smu += g; sls = fmaf(g * es, nz, sls);
                if (a.grad_noise) a.grad_noise[((size_t)b * K + k) * Ds + c] = g * es;
            }
            atomicAdd(G + po.slot_mu + c, smu);
            atomicAdd(G + po.slot_log_sigma + c, sls);
        }
    }
#undef LIN
}

// ---------------------------------------------------------------------------
// K5: token LayerNorm backward (steve.py:60): d inputs, d gamma, d beta.
// One warp per token; each lane owns 4-element groups.  D <= 512.
// ---------------------------------------------------------------------------
template <typename TokT>
__global__ void __launch_bounds__(NT) ln_tokens_bwd_kernel(const TokT* __restrict__ x, const float* __restrict__ dxhat,
                                                           const float2* __restrict__ stats, const float* __restrict__ g,
                                                           TokT* __restrict__ dx, float* dg_glob, float* db_glob,
                                                           int64_t rows, int D) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = D >> 2;                         // 4-element groups per row (<= 128)
    float ag[4][4], ab[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { ag[j][e] = 0.f; ab[j][e] = 0.f; }
    const int64_t warp0 = (int64_t)blockIdx.x * NW + warp, nwarps = (int64_t)gridDim.x * NW;
    for (int64_t row = warp0; row < rows; row += nwarps) {
        const float2 st = stats[row];
        float z[4][4], dz[4][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gi = lane + 32 * j;
            if (gi < groups) {
                const float4 xv = Tok<TokT>::load4(x + row * D + gi * 4);
                const float4 dv = ld4(dxhat + row * D + gi * 4);
                const float4 gg = __ldg(reinterpret_cast<const float4*>(g + gi * 4));
                const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w}, gw[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    z[j][e] = (xx[e] - st.x) * st.y;
                    dz[j][e] = dd[e] * gw[e];
                    s1 += dz[j][e]; s2 = fmaf(dz[j][e], z[j][e], s2);
                    ag[j][e] = fmaf(dd[e], z[j][e], ag[j][e]); ab[j][e] += dd[e];
                }
            }
        }
        s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gi = lane + 32 * j;
            if (gi < groups) {
                float4 o;
                o.x = st.y * (dz[j][0] - s1 - z[j][0] * s2); o.y = st.y * (dz[j][1] - s1 - z[j][1] * s2);
                o.z = st.y * (dz[j][2] - s1 - z[j][2] * s2); o.w = st.y * (dz[j][3] - s1 - z[j][3] * s2);
                Tok<TokT>::store4(dx + row * D + gi * 4, o);
            }
        }
    }
    // block-level reduction of d gamma / d beta, then one atomic per column per block
    __shared__ float red[2][512];
    for (int i = threadIdx.x; i < 2 * 512; i += NT) (&red[0][0])[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int gi = lane + 32 * j;
        if (gi < groups) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { atomicAdd(&red[0][gi * 4 + e], ag[j][e]); atomicAdd(&red[1][gi * 4 + e], ab[j][e]); }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += NT) { atomicAdd(dg_glob + c, red[0][c]); atomicAdd(db_glob + c, red[1][c]); }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
template <typename TokT, int KMAX, bool MMA>
static cudaError_t launch_bwd_t(const BwdArgs& a, cudaStream_t st) {
    auto kern = savi_bwd_kernel<TokT, KMAX, MMA>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.d.B * a.d.CN);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = a.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.d.CN; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <typename TokT>
static cudaError_t launch_bwd_k(const BwdArgs& a, cudaStream_t st) {
    const int K = a.d.K;
    if constexpr (sizeof(TokT) == 2) {
        if (a.d.mma) {
            if (K <= 16) return launch_bwd_t<TokT, 16, true>(a, st);
            if (K <= 32) return launch_bwd_t<TokT, 32, true>(a, st);
            return launch_bwd_t<TokT, 64, true>(a, st);
        }
    }
    if (K <= 8) return launch_bwd_t<TokT, 8, false>(a, st);
    if (K <= 16) return launch_bwd_t<TokT, 16, false>(a, st);
    if (K <= 24) return launch_bwd_t<TokT, 24, false>(a, st);
    if (K <= 32) return launch_bwd_t<TokT, 32, false>(a, st);
    return launch_bwd_t<TokT, 64, false>(a, st);
}

#ifndef SAVI_TOK
#error "compile with -DSAVI_TOK=float -DSAVI_SUFFIX=f32 (or __nv_bfloat16 / bf16)"
#endif
#define SAVI_CAT2(a, b) a##b
#define SAVI_CAT(a, b) SAVI_CAT2(a, b)

cudaError_t SAVI_CAT(savi_launch_bwd_clip_, SAVI_SUFFIX)(const BwdArgs& a, cudaStream_t st) { return launch_bwd_k<SAVI_TOK>(a, st); }

cudaError_t SAVI_CAT(savi_launch_ln_bwd_, SAVI_SUFFIX)(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st) {
    const Dims& d = a.d;
    const int64_t rows = (int64_t)d.B * d.T * d.N;
    int grid = (int)((rows + NW * 8 - 1) / (NW * 8));
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    const float2* stats = reinterpret_cast<const float2*>(a.saved + a.sl.stats);
    ln_tokens_bwd_kernel<SAVI_TOK><<<grid, NT, 0, st>>>(reinterpret_cast<const SAVI_TOK*>(inputs), a.ws + a.wl.dxhat, stats,
                                                       a.packed + a.po.ln_in_w, reinterpret_cast<SAVI_TOK*>(grad_inputs),
                                                       a.grad_params + a.po.ln_in_w, a.grad_params + a.po.ln_in_b, rows, d.D);
    return cudaGetLastError();
}

#if defined(SAVI_IS_BF16)
size_t savi_dx_smem_bytes(const Dims& d) { return dx_smem_bytes(dx_pick_ig(d.I, d.KC, d.D, 227 * 1024), d.KC, d.D); }

cudaError_t savi_launch_dx_mma(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st) {
    const Dims& d = a.d;
    DxArgs x;
    x.x = reinterpret_cast<const bf16*>(inputs);
    x.stats = reinterpret_cast<const float2*>(a.saved + a.sl.stats);
    x.coef = reinterpret_cast<const bf16*>(reinterpret_cast<const unsigned char*>(a.ws) + a.wl.coef);
    x.qk = reinterpret_cast<const float*>(a.saved + a.sl.fbase) + a.sl.qk;
    x.dux = a.ws + a.wl.duxs;
    x.gamma = a.packed + a.po.ln_in_w;
    x.dx = reinterpret_cast<bf16*>(grad_inputs);
    x.dgamma = a.grad_params + a.po.ln_in_w; x.dbeta = a.grad_params + a.po.ln_in_b;
    x.B = d.B; x.T = d.T; x.N = d.N; x.D = d.D; x.K = d.K; x.I = d.I; x.KC = d.KC; x.tiles_per_cta = 4;
    x.IG = dx_pick_ig(d.I, d.KC, d.D, 227 * 1024);
    if (x.IG < 1) return cudaErrorInvalidValue;
    const int tiles = (d.N + TMMA_TN - 1) / TMMA_TN;
    dim3 grid((tiles + x.tiles_per_cta - 1) / x.tiles_per_cta, d.B * d.T);
    const int smem = (int)dx_smem_bytes(x.IG, d.KC, d.D);
    cudaError_t e;
#define DX_LAUNCH(ND_) \
    e = cudaFuncSetAttribute(dx_finalize_kernel<ND_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e != cudaSuccess) return e; \
    dx_finalize_kernel<ND_><<<grid, DX_NT, smem, st>>>(x);
    if (d.D <= 64) { DX_LAUNCH(8) } else if (d.D <= 128) { DX_LAUNCH(16) } else if (d.D <= 192) { DX_LAUNCH(24) } else { DX_LAUNCH(32) }
#undef DX_LAUNCH
    return cudaGetLastError();
}
#endif
