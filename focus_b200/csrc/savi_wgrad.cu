// Weight-gradient GEMM and the backward launch sequence.
#include "savi_dev.cuh"
#include <cstdlib>
#include "savi_args.h"
#include "savi_smallgemm.cuh"

// ---------------------------------------------------------------------------
// K4: weight gradients.  dW[o][c] += alpha * sum_r dY[r][o] X[r][c] for every weight
// matrix in one launch; r runs over all steps and clips (tall-skinny TN GEMM,
// split along r, fp32 atomics into the zeroed flat gradient buffer).
// ---------------------------------------------------------------------------
constexpr int WG_T = 64, WG_R = 16;
__global__ void __launch_bounds__(256) wgrad_kernel(const __grid_constant__ WgradArgs wa) {
    const WgradJob& jb = wa.job[blockIdx.z];
    const int to = (jb.O + WG_T - 1) / WG_T, tc = (jb.C + WG_T - 1) / WG_T;
    if ((int)blockIdx.x >= to * tc) return;
    const int o0 = (blockIdx.x / tc) * WG_T, c0 = (blockIdx.x % tc) * WG_T;
    const int r_begin = blockIdx.y * wa.rows_per_split;
    if (r_begin >= jb.R) return;
    const int r_end = min(jb.R, r_begin + wa.rows_per_split);
    __shared__ __align__(16) float ys[WG_R][WG_T];
    __shared__ __align__(16) float xs[WG_R][WG_T];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lr = tid >> 4, lc = (tid & 15) * 4;        // loader: row lr, 4 columns at lc
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r0 = r_begin; r0 < r_end; r0 += WG_R) {
        const int r = r0 + lr;
        float4 yv = make_float4(0.f, 0.f, 0.f, 0.f), xv = yv;
        if (r < r_end) {
            const float* yp = jb.dY + (size_t)r * jb.ldy + o0 + lc;
            const float* xp = jb.X + (size_t)r * jb.ldx + c0 + lc;
            if (o0 + lc + 3 < jb.O) yv = ld4(yp);
            else { if (o0 + lc < jb.O) yv.x = yp[0]; if (o0 + lc + 1 < jb.O) yv.y = yp[1]; if (o0 + lc + 2 < jb.O) yv.z = yp[2]; }
            if (c0 + lc + 3 < jb.C) xv = ld4(xp);
            else { if (c0 + lc < jb.C) xv.x = xp[0]; if (c0 + lc + 1 < jb.C) xv.y = xp[1]; if (c0 + lc + 2 < jb.C) xv.z = xp[2]; }
        }
        __syncthreads();
        st4(&ys[lr][lc], yv);
        st4(&xs[lr][lc], xv);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < WG_R; ++rr) {
            const float4 y = ld4(&ys[rr][ty * 4]);
            const float4 x = ld4(&xs[rr][tx * 4]);
            const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(yy[i], x.x, acc[i][0]); acc[i][1] = fmaf(yy[i], x.y, acc[i][1]);
                acc[i][2] = fmaf(yy[i], x.z, acc[i][2]); acc[i][3] = fmaf(yy[i], x.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = o0 + ty * 4 + i;
        if (o < jb.O) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + tx * 4 + j;
                if (c < jb.C) atomicAdd(jb.dW + (size_t)o * jb.C + c, jb.alpha * acc[i][j]);
            }
        }
    }
}


// Chain rule through the folded weights of the tcgen05 path (savi_layout.h: ParamOff::wqk, wg), s = Ds^-1/2:
//   wqk[d][c] = s sum_a Wk[a][d] Wq[a][c]  =>  dWq[a][c] = s sum_d Wk[a][d] dwqk[d][c],   dWk[a][d] = s sum_c Wq[a][c] dwqk[d][c]
//   wg[g][d]  =   sum_a Wih[g][a] Wv[a][d]  =>  dWih[g][a] = sum_d dwg[g][d] Wv[a][d],      dWv[a][d] = sum_g Wih[g][a] dwg[g][d]
// (project_q / project_k / project_v / gru.weight_ih receive no other contribution on this path: plain stores.)
static cudaError_t launch_fold_grads(const float* P, const float* dwqk, const float* dwg, float* G, const ParamOff& po,
                                     int D, int Ds, float s, cudaStream_t st) {
    SmallGemmArgs ga; ga.count = 0;
    ga.g[ga.count++] = SmallGemm{P + po.wk, dwqk, G + po.wq, Ds, Ds, D, D, 1, Ds, 1, s, 0};
    ga.g[ga.count++] = SmallGemm{P + po.wq, dwqk, G + po.wk, Ds, D, Ds, Ds, 1, 1, Ds, s, 0};
    ga.g[ga.count++] = SmallGemm{dwg, P + po.wv, G + po.wih, 3 * Ds, Ds, D, D, 1, 1, D, 1.0f, 0};
    // dWv contracts over the 3 Ds gate rows: one product per gate block, accumulated into the (zeroed) gradient buffer, so that no
    // CTA walks more than one 128-deep chunk
    for (int gate = 0; gate < 3; ++gate)
        ga.g[ga.count++] = SmallGemm{P + po.wih + (size_t)gate * Ds * Ds, dwg + (size_t)gate * Ds * D, G + po.wv, Ds, D, Ds, 1, Ds, D, 1, 1.0f, 1};
    return launch_small_gemms(ga, st);
}

cudaError_t savi_launch_forward(const FwdArgs& a, const void* inputs, cudaStream_t st, int* launches) {
    return a.d.tok_bytes == 4 ? savi_launch_forward_f32(a, inputs, st, launches) : savi_launch_forward_bf16(a, inputs, st, launches);
}

static void add_job(WgradArgs& wa, const float* dY, int ldy, const float* X, int ldx, float* dW, int64_t R, int O, int C, float alpha) {
    if (R <= 0) return;
    WgradJob& j = wa.job[wa.njobs++];
    j.dY = dY; j.X = X; j.dW = dW; j.R = (int)R; j.O = O; j.C = C; j.ldy = ldy; j.ldx = ldx; j.alpha = alpha;
}

cudaError_t savi_launch_backward(const BwdArgs& a, const void* inputs, void* grad_inputs, cudaStream_t st, int* launches) {
    const Dims& d = a.d;
    cudaError_t e = cudaMemsetAsync(a.grad_params, 0, (size_t)a.po.total * sizeof(float), st);
    if (e != cudaSuccess) return e;
    const bool umma_bwd = d.umma != 0;     // the tcgen05 forward saves only what the tcgen05 backward reads (no q, no U)
    const bool overlap_dx = umma_bwd && !savi_options().no_overlap;
    if (umma_bwd) {
        e = cudaMemsetAsync(reinterpret_cast<unsigned char*>(a.ws) + a.wl.flags, 0, ((size_t)d.B * d.T + 1) * sizeof(int), st);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.ws + a.wl.dwqk, 0, (size_t)(a.wl.dwg - a.wl.dwqk + (int64_t)3 * d.Ds * d.D) * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    savi_prof_begin(3, st);
    if (umma_bwd) {
        WImg wi;
        savi_wimg_layout(d.D, d.Ds, d.M, d.blocks, wi);
        e = savi_launch_bwd_umma(a, reinterpret_cast<const unsigned char*>(a.packed) + savi_wimg_base(a.po.packed_total), wi, st);
    } else if (d.tok_bytes == 4) e = savi_launch_bwd_clip_f32(a, st); else e = savi_launch_bwd_clip_bf16(a, st);
    savi_prof_end(3, st);
    if (e != cudaSuccess) return e;
    *launches += 1;
    if (umma_bwd) {
        // d_inputs directly behind the clip kernel as its programmatic dependent: its CTAs start on the SMs the clip grid
        // leaves idle (148 - B*CN) and on every SM a finished clip CTA frees, gated per (clip, frame) by the staging flags
        savi_prof_begin(5, st);
        e = savi_launch_dx_umma(a, inputs, grad_inputs, overlap_dx, st);
        savi_prof_end(5, st);
        if (e != cudaSuccess) return e;
        *launches += 1;
    }

    const float* fb = reinterpret_cast<const float*>(a.saved + a.sl.fbase);
    const float* W = a.ws;
    float* G = a.grad_params;
    const int Ds = d.Ds, D = d.D, M = d.M;
    const int64_t R = (int64_t)d.S * d.B * d.K, Rm = (int64_t)d.Sm * d.B * d.K, Rb = (int64_t)(d.T - 1) * d.B * d.K;
    WgradArgs wa; wa.njobs = 0;
    // the tcgen05 forward saves s~ and LN_m(h'): its backward does not restage them
    const float* st_rows = umma_bwd ? fb + a.sl.st : W + a.wl.st;
    const float* m_rows = umma_bwd ? fb + a.sl.m : W + a.wl.m;
    if (umma_bwd) {
        // the tcgen05 kernels multiply by the folded weights wqk = Ds^-1/2 Wk^T Wq and wg = Wih Wv (savi_layout.h): their
        // gradients are reduced here into scratch and mapped back to the four parameters by fold_grads_kernel below
        float* Wm = a.ws;
        add_job(wa, W + a.wl.dqk, D, st_rows, Ds, Wm + a.wl.dwqk, R, D, Ds, 1.0f);
        add_job(wa, W + a.wl.dgi, 3 * Ds, fb + a.sl.ux, D, Wm + a.wl.dwg, R, 3 * Ds, D, 1.0f);
    } else {
        add_job(wa, W + a.wl.dq, Ds, st_rows, Ds, G + a.po.wq, R, Ds, Ds, 1.0f);
        add_job(wa, fb + a.sl.q, Ds, W + a.wl.dqk, D, G + a.po.wk, R, Ds, D, d.qscale);
        add_job(wa, W + a.wl.du, Ds, fb + a.sl.ux, D, G + a.po.wv, R, Ds, D, 1.0f);
        add_job(wa, W + a.wl.dgi, 3 * Ds, fb + a.sl.u, Ds, G + a.po.wih, R, 3 * Ds, Ds, 1.0f);
    }
    add_job(wa, W + a.wl.dgh, 3 * Ds, fb + a.sl.hp, Ds, G + a.po.whh, R, 3 * Ds, Ds, 1.0f);
    add_job(wa, W + a.wl.da, M, m_rows, Ds, G + a.po.w1, Rm, M, Ds, 1.0f);
    add_job(wa, W + a.wl.dhm, Ds, fb + a.sl.a, M, G + a.po.w2, Rm, Ds, M, 1.0f);
    for (int j = 0; j < d.blocks; ++j) {
        const int64_t ro = (int64_t)j * Rb;      // block j's rows are contiguous: f = j*(T-1)+t
        const BlockOff& bo = a.po.blk[j];
        add_job(wa, W + a.wl.pdq + ro * Ds, Ds, fb + a.sl.py + ro * Ds, Ds, G + bo.pq, Rb, Ds, Ds, 1.0f);
        add_job(wa, W + a.wl.pdk + ro * Ds, Ds, fb + a.sl.py + ro * Ds, Ds, G + bo.pk, Rb, Ds, Ds, 1.0f);
        add_job(wa, W + a.wl.pdv + ro * Ds, Ds, fb + a.sl.py + ro * Ds, Ds, G + bo.pv, Rb, Ds, Ds, 1.0f);
        add_job(wa, W + a.wl.pdx1 + ro * Ds, Ds, fb + a.sl.po + ro * Ds, Ds, G + bo.po, Rb, Ds, Ds, 1.0f);
        add_job(wa, W + a.wl.pdf + ro * 4 * Ds, 4 * Ds, fb + a.sl.pl2 + ro * Ds, Ds, G + bo.f1, Rb, 4 * Ds, Ds, 1.0f);
        add_job(wa, W + a.wl.pdx2 + ro * Ds, Ds, fb + a.sl.pf + ro * 4 * Ds, 4 * Ds, G + bo.f2, Rb, Ds, 4 * Ds, 1.0f);
    }
    if (wa.njobs > 0 && umma_bwd) {
        savi_prof_begin(4, st);
        // behind d_inputs as ITS programmatic dependent: starts once every d_inputs CTA is resident, waits for the clip kernel's
        // CTA counter, then fills the SMs the last d_inputs wave leaves idle
        const int* done = overlap_dx ? reinterpret_cast<const int*>(reinterpret_cast<const unsigned char*>(a.ws) + a.wl.flags) + d.B * d.T : nullptr;
        e = savi_launch_wgrad_umma(wa, done, d.B * d.CN, st);
        if (e != cudaSuccess) return e;
        e = launch_fold_grads(a.packed, a.ws + a.wl.dwqk, a.ws + a.wl.dwg, G, a.po, D, Ds, d.qscale, st);
        savi_prof_end(4, st);
        if (e != cudaSuccess) return e;
        *launches += 2;
    } else if (wa.njobs > 0) {
        int maxR = 0, maxTiles = 0;
        for (int j = 0; j < wa.njobs; ++j) {
            maxR = wa.job[j].R > maxR ? wa.job[j].R : maxR;
            int tl = ((wa.job[j].O + WG_T - 1) / WG_T) * ((wa.job[j].C + WG_T - 1) / WG_T);
            maxTiles = tl > maxTiles ? tl : maxTiles;
        }
        wa.rows_per_split = 512;
        int splits = (maxR + wa.rows_per_split - 1) / wa.rows_per_split;
        if (splits > 256) { wa.rows_per_split = ((maxR + 255) / 256 + WG_R - 1) / WG_R * WG_R; splits = (maxR + wa.rows_per_split - 1) / wa.rows_per_split; }
        savi_prof_begin(4, st);
        wgrad_kernel<<<dim3(maxTiles, splits, wa.njobs), 256, 0, st>>>(wa);
        savi_prof_end(4, st);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        *launches += 1;
    }
    if (umma_bwd) return cudaSuccess;
    savi_prof_begin(5, st);
    if (d.mma) e = savi_launch_dx_mma(a, inputs, grad_inputs, st);
    else if (d.tok_bytes == 4) e = savi_launch_ln_bwd_f32(a, inputs, grad_inputs, st);
    else e = savi_launch_ln_bwd_bf16(a, inputs, grad_inputs, st);
    savi_prof_end(5, st);
    *launches += 1;
    return e;
}
