"""B200-native versions of the components next to the slot-attention path, with the
reference's own call signatures.  CUDA only (include/focus_steve.h); no CPU / PyTorch fallback.

  token_mlp(emb, layer_norm, mlp)                 reference slowfast/models/STEVE/steve.py:307-309 / :342-344: the encoder tail
                                                  that turns the CNN map into the slot-attention module's `inputs`
  attention_overlay(video, attns, H_enc, W_enc)   reference slowfast/models/STEVE/steve.py:314-319 (STEVE.forward) and
                                                  :349-355 (STEVE.encode): the per-slot attention overlays
  evaluate_ari(true_mask, pred_mask)              reference slowfast/utils/metrics.py:58-83, the FG-ARI of
                                                  tools/steve_eval_net.py:107-108
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def token_mlp(emb, layer_norm, mlp, out_dtype=torch.bfloat16):
    """emb [BT, C, H, W] fp32 (CNN + positional embedding, channels first), `layer_norm` = steve_encoder.layer_norm,
    `mlp` = steve_encoder.mlp (Sequential(Linear, ReLU, Linear)) -> emb_set [BT, H*W, C] =
        mlp(layer_norm(emb.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)))
    in one kernel.  Forward only: use it under torch.no_grad() (STEVE.encode, evaluation); training keeps the reference's
    own modules, whose backward autograd provides."""
    if not emb.is_cuda:
        raise RuntimeError("focus_b200.neighbors.token_mlp has no CPU path")
    if torch.is_grad_enabled() and (emb.requires_grad or any(p.requires_grad for p in list(layer_norm.parameters()) + list(mlp.parameters()))):
        raise RuntimeError("focus_b200.neighbors.token_mlp is forward-only: call it under torch.no_grad() (evaluation / STEVE.encode)")
    BT, C, H, W = emb.shape
    lin1, lin2 = mlp[0], mlp[2]
    if tuple(lin1.weight.shape) != (C, C) or tuple(lin2.weight.shape) != (C, C) or tuple(layer_norm.normalized_shape) != (C,):
        raise ValueError("token_mlp expects LayerNorm(C) -> Linear(C, C) -> ReLU -> Linear(C, C) with C = emb.shape[1] = %d" % C)
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be torch.float32 or torch.bfloat16")
    dev = emb.device
    f = lambda t: t.detach().float().contiguous()
    emb = f(emb)
    out = torch.empty(BT, H * W, C, dtype=out_dtype, device=dev)
    ws = torch.empty(int(_lib.lib.steve_token_mlp_ws_bytes(C)), dtype=torch.uint8, device=dev)
    args = [f(layer_norm.weight), f(layer_norm.bias), f(lin1.weight), f(lin1.bias), f(lin2.weight), f(lin2.bias)]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.steve_token_mlp(_p(emb), *[_p(t) for t in args], _p(out),
                                            _lib.SAVI_DTYPE_F32 if out_dtype == torch.float32 else _lib.SAVI_DTYPE_BF16,
                                            BT, H * W, C, float(layer_norm.eps), _p(ws), _stream(dev)), "steve_token_mlp")
    return out


def attention_overlay(video, attns, H_enc, W_enc, want_overlay=True, want_up=True):
    """video [B,T,C,H,W] fp32, attns [B,T,H_enc*W_enc,K] (the slot-attention module's second output, fp32 or bf16)
    -> (overlay [B,T,K,C,H,W], up [B,T,K,1,H,W]), both fp32:
        up      = attns.transpose(-1,-2).reshape(B,T,K,1,H_enc,W_enc).repeat_interleave(H//H_enc,-2).repeat_interleave(W//W_enc,-1)
        overlay = video.unsqueeze(2) * up + (1. - up)
    Forward only: the reference never differentiates these maps (they feed the visualisation / the evaluation masks,
    tools/steve_train_net.py:149, tools/steve_eval_net.py:75-108), so the result is detached."""
    if not (video.is_cuda and attns.is_cuda):
        raise RuntimeError("focus_b200.neighbors.attention_overlay has no CPU path")
    B, T, C, H, W = video.shape
    K = attns.shape[-1]
    if tuple(attns.shape) != (B, T, H_enc * W_enc, K):
        raise ValueError("attns must be [B, T, H_enc*W_enc, K], got %s" % (tuple(attns.shape),))
    dev = video.device
    video = video.detach().float().contiguous()
    attns = attns.detach()
    if attns.dtype not in (torch.float32, torch.bfloat16):
        attns = attns.float()
    attns = attns.contiguous()
    overlay = torch.empty(B, T, K, C, H, W, dtype=torch.float32, device=dev) if want_overlay else None
    up = torch.empty(B, T, K, 1, H, W, dtype=torch.float32, device=dev) if want_up else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.steve_attention_overlay(_p(attns), _lib.SAVI_DTYPE_F32 if attns.dtype == torch.float32 else _lib.SAVI_DTYPE_BF16,
                                                    _p(video), _p(overlay), _p(up), B * T, K, C, H, W, H_enc, W_enc, _stream(dev)),
                   "steve_attention_overlay")
    return overlay, up


def ari_tables(true_mask, pred_mask):
    """[B,N0,D], [B,N1,D] -> int32 contingency tables [B,N0,N1] of metrics.py:40-57 (on the device)."""
    if not (true_mask.is_cuda and pred_mask.is_cuda):
        raise RuntimeError("focus_b200.neighbors.ari_tables has no CPU path")
    B, N1, D = pred_mask.shape
    N0 = true_mask.shape[1]
    if true_mask.shape[0] != B or true_mask.shape[2] != D:
        raise ValueError("true_mask %s does not match pred_mask %s" % (tuple(true_mask.shape), tuple(pred_mask.shape)))
    dev = pred_mask.device
    tm = true_mask.detach().float().contiguous()
    pm = pred_mask.detach().float().contiguous()
    tables = torch.empty(B, N0, N1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.steve_ari_tables(_p(tm), _p(pm), _p(tables), B, N0, N1, D, _stream(dev)), "steve_ari_tables")
    return tables


def _comb2(x):
    x = np.asarray(x, np.float64)
    return x * (x - 1.0) / 2.0            # scipy.special.comb(x, 2) for the non-negative integer counts of a contingency table


def ari_from_table(table):
    """The reference's closed form (metrics.py:10-36) on one [N0,N1] integer table, float64."""
    table = np.asarray(table, np.float64)
    a, b = table.sum(1), table.sum(0)
    n = a.sum()
    ca, cb, cn, ct = _comb2(a).sum(), _comb2(b).sum(), _comb2(n), _comb2(table).sum()
    if cb == ca == cn == ct:
        return 1.0
    return float((ct - ca * cb / cn) / (0.5 * (ca + cb) - (ca * cb) / cn))


def evaluate_ari(true_mask, pred_mask):
    """Same signature and value as the reference's metrics.evaluate_ari: true_mask [B,N0,D], pred_mask [B,N1,D] -> average
    ARI over the batch.  One kernel launch + one [B,N0,N1] int32 read-back instead of a per-sample CPU loop over an
    [N0,N1,D] boolean broadcast."""
    tables = ari_tables(true_mask, pred_mask).cpu().numpy()
    return float(sum(ari_from_table(t) for t in tables) / tables.shape[0])
