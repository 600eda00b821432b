/*
 * focus_steve.h — C ABI of the components NEXT to the slot-attention path, in the same library
 * (libfocus_savi.so).  Same conventions as focus_savi.h: plain device pointers and sizes, caller-owned memory, all
 * work enqueued on `stream`, 0 or a negative SAVI_E* code, savi_last_error() for the message, no CPU fallback.
 *
 *  N1  token-encoder tail       reference slowfast/models/STEVE/steve.py:307-309 (STEVE.forward), :342-344 (STEVE.encode)
 *  N2  attention-map consumer   reference slowfast/models/STEVE/steve.py:314-319 (STEVE.forward), :349-355 (STEVE.encode)
 *  N4  FG-ARI contingency table reference slowfast/utils/metrics.py:40-83 (compute_mask_ari / evaluate_ari),
 *                               called from tools/steve_eval_net.py:107-108
 */
#ifndef FOCUS_STEVE_H
#define FOCUS_STEVE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* N1.  The tokens the slot-attention module consumes are made from the CNN feature map by
 *     emb_set = mlp(layer_norm(emb.permute(0,2,3,1).flatten(1,2)))        (LayerNorm(C) -> Linear(C,C) + ReLU -> Linear(C,C))
 * Forward only (STEVE.encode / evaluation), one tcgen05 kernel: the map is read once, the tokens are written once.
 * bf16 tensor-core operands with fp32 accumulation (the accuracy class of the reference under bf16 autocast).
 *   emb   [BT, C, HW]  fp32, channels-first (the CNN + positional-embedding output, steve.py:301-304)
 *   ln_w, ln_b [C]; w1, w2 [C, C] (nn.Linear layout: out x in); b1, b2 [C]      fp32
 *   out   [BT, HW, C]  out_dtype: SAVI_DTYPE_F32 or SAVI_DTYPE_BF16 (C = 192: bf16 only)
 *   ws    steve_token_mlp_ws_bytes(C) bytes of scratch (the packed bf16 weight images)
 * C in {64, 128, 192}. */
int64_t steve_token_mlp_ws_bytes(int C);
int steve_token_mlp(const float* emb, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                    const float* w2, const float* b2, void* out, int out_dtype, int64_t BT, int HW, int C,
                    float ln_eps, void* ws, void* stream);

/* N2.  The slot-attention maps are consumed as per-slot image overlays:
 *     attns.transpose(-1,-2).reshape(B,T,K,1,He,We).repeat_interleave(H/He,-2).repeat_interleave(W/We,-1)   -> up
 *     video.unsqueeze(2) * up + (1 - up)                                                                     -> overlay
 * One pass instead of the reference's six (transpose copy, two repeat_interleave copies, mul, rsub, add):
 *   attn     [BT, He*We, K]    token dtype of the slot-attention output (attn_dtype: SAVI_DTYPE_F32 / SAVI_DTYPE_BF16)
 *   video    [BT, C, H, W]     fp32
 *   overlay  [BT, K, C, H, W]  fp32   (written; may be NULL)
 *   up       [BT, K, 1, H, W]  fp32   (written; may be NULL)  — the third output of STEVE.encode
 * H % He == 0 and W % We == 0 (the reference's repeat_interleave factors). */
int steve_attention_overlay(const void* attn, int attn_dtype, const float* video, float* overlay, float* up,
                            int64_t BT, int K, int C, int H, int W, int He, int We, void* stream);

/* N4.  FG-ARI: for every sample b, the contingency table between the ground-truth segments and the argmax
 * segmentation of the predicted masks,
 *     table[b][i][j] = #{ d : (uint8(true_mask[b,i,d]) & 1) and argmax_j' pred_mask[b,j',d] == j }
 * (metrics.py:50-57 after :72-76; first maximum wins, as torch.argmax).  Integer-exact; the ARI itself is the reference's
 * closed form on that table (evaluated on the host in float64 by focus_b200.neighbors.evaluate_ari).
 *   true_mask [B, N0, D] fp32      pred_mask [B, N1, D] fp32      tables [B, N0, N1] int32 (overwritten)
 * N0, N1 <= 64. */
int steve_ari_tables(const float* true_mask, const float* pred_mask, int32_t* tables,
                     int B, int N0, int N1, int64_t D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FOCUS_STEVE_H */
