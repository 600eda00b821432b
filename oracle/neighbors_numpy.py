"""numpy restatement of the components next to the slot-attention path.

TEST INFRASTRUCTURE ONLY (same rules as oracle/savi_numpy.py): nothing under focus_b200/ imports this file.

  token_mlp         : /root/reference/slowfast/models/STEVE/steve.py:307-309 (STEVE.forward), :342-344 (STEVE.encode)
  attention_overlay : /root/reference/slowfast/models/STEVE/steve.py:314-319 (STEVE.forward), :349-355 (STEVE.encode)
  ari_tables / evaluate_ari : /root/reference/slowfast/utils/metrics.py:10-36 (compute_ari), :40-57 (compute_mask_ari),
                              :58-83 (evaluate_ari)
Pinned in tests/test_neighbors.py against the reference functions themselves (imported from /root/reference, or from the
oracle/_ref copies on the GPU box) on seeded inputs, ties and degenerate tables included.
"""
import numpy as np


def _round_bf16(a):
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32).astype(np.asarray(a).dtype).reshape(np.shape(a))


def token_mlp(emb, ln_w, ln_b, w1, b1, w2, b2, eps=1e-5, operand_dtype=None):
    """emb [BT,C,H,W] -> emb_set [BT,H*W,C] = mlp(layer_norm(emb.permute(0,2,3,1).flatten(1,2)))  (steve.py:307-308), float64.
    operand_dtype="bf16" models the CUDA kernel's tensor-core operands: it folds the LayerNorm affine into the first Linear
    (W1' = W1 diag(ln_w), b1' = b1 + W1 ln_b — the same function in exact arithmetic) and rounds the normalised token, W1', the
    hidden activations and W2 to bfloat16 before each product (accumulation and biases stay full precision)."""
    x = np.asarray(emb, np.float64)
    BT, C, H, W = x.shape
    x = x.transpose(0, 2, 3, 1).reshape(BT, H * W, C)
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    xh = xc / np.sqrt((xc * xc).mean(-1, keepdims=True) + eps)
    ln_w, ln_b, w1, b1 = (np.asarray(t, np.float64) for t in (ln_w, ln_b, w1, b1))
    if operand_dtype == "bf16":
        r = _round_bf16
        h = np.maximum(r(xh) @ r(w1 * ln_w[None, :]).T + (b1 + w1 @ ln_b), 0.0)
    else:
        r = lambda a: a
        h = np.maximum((xh * ln_w + ln_b) @ w1.T + b1, 0.0)
    return r(h) @ r(np.asarray(w2, np.float64)).T + np.asarray(b2, np.float64)


def attention_overlay(video, attns, H_enc, W_enc):
    """video [B,T,C,H,W], attns [B,T,H_enc*W_enc,K] -> (overlay [B,T,K,C,H,W], up [B,T,K,1,H,W]) in float32 arithmetic
    (the reference runs this in the dtype of `video`, float32)."""
    video = np.asarray(video, np.float32)
    B, T, C, H, W = video.shape
    K = attns.shape[-1]
    up = np.asarray(attns, np.float32).swapaxes(-1, -2).reshape(B, T, K, 1, H_enc, W_enc)      # steve.py:314-316
    up = np.repeat(np.repeat(up, H // H_enc, axis=-2), W // W_enc, axis=-1)                  # :317-318
    overlay = video[:, :, None] * up + (np.float32(1.0) - up)                                # :319
    return overlay, up


def ari_tables(true_mask, pred_mask):
    """[B,N0,D], [B,N1,D] -> int64 [B,N0,N1]: argmax one-hot of the predictions (metrics.py:72-76, first maximum wins) against
    the low bit of the uint8-truncated ground truth (the `.byte()` + `&` of :50-55)."""
    true_mask, pred_mask = np.asarray(true_mask), np.asarray(pred_mask)
    B, N1, D = pred_mask.shape
    N0 = true_mask.shape[1]
    arg = pred_mask.argmax(1)                                      # [B,D]
    t = (true_mask.astype(np.int64).astype(np.uint8) & 1).astype(np.int64)   # float -> int -> uint8 (C truncation / wrap), low bit
    out = np.zeros((B, N0, N1), np.int64)
    for b in range(B):
        for j in range(N1):
            out[b, :, j] = t[b][:, arg[b] == j].sum(1)
    return out


def ari_from_table(table):
    """metrics.py:10-36 with comb(x, 2) = x (x - 1) / 2."""
    table = np.asarray(table, np.float64)
    c2 = lambda x: x * (x - 1.0) / 2.0
    a, b = table.sum(1), table.sum(0)
    n = a.sum()
    ca, cb, cn, ct = c2(a).sum(), c2(b).sum(), c2(n), c2(table).sum()
    if cb == ca == cn == ct:
        return 1.0
    return float((ct - ca * cb / cn) / (0.5 * (ca + cb) - (ca * cb) / cn))


def evaluate_ari(true_mask, pred_mask):
    tabs = ari_tables(true_mask, pred_mask)
    return float(sum(ari_from_table(t) for t in tabs) / len(tabs))
