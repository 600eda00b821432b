"""Integration tier: the UNMODIFIED reference code with our module swapped in.

The reference package is imported from /root/reference in the build container and from oracle/_ref (byte-identical
copies made by oracle/make_ref.py, git-ignored, shipped by gpurun) on the GPU box.
  * CPU tests: the travel recipe reproduces the reference files, and the stock STEVE model accepts our module as
    `steve_encoder.savi` (construction, state_dict, parameter groups) exactly as INTEGRATION.md describes.
  * GPU tests: reference STEVE vs the same model with `.savi` swapped (steve.py:229-232, 311, 346): one training-style
    forward + backward (tools/steve_train_net.py:95-126) and one `encode`; CUDA-autocast output dtypes.
"""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import _load_reference as LR
from oracle import make_ref
from tests._util import err

needs_ref = pytest.mark.skipif(not LR.reference_available(), reason="reference sources not present (run oracle/make_ref.py)")


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="build container only")
def test_make_ref_copies_are_byte_identical():
    make_ref.make()
    assert make_ref.check()
    for rel in make_ref.FILES:
        with open(os.path.join(make_ref.REF_ROOT, rel), "rb") as f, open(os.path.join(make_ref.OUT, rel), "rb") as g:
            assert f.read() == g.read(), rel


def _swap(model):
    """replace steve_encoder.savi by the B200 module and load the reference weights."""
    from focus_b200 import SlotAttentionVideo
    old = model.steve_encoder.savi
    blocks = len(old.predictor.blocks)
    heads = old.predictor.blocks[0].attn.num_heads if blocks else 1
    new = SlotAttentionVideo(old.num_iterations, old.num_slots, old.input_size, old.slot_size, old.mlp_hidden_size,
                             blocks, heads, 0.0)
    new.load_state_dict(old.state_dict(), strict=True)
    model.steve_encoder.savi = new
    return model


@needs_ref
def test_stock_steve_accepts_the_swapped_module_cpu():
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    ref = steve.STEVE(LR.steve_config())
    ours = _swap(copy.deepcopy(ref))
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    for (k, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
        assert a.shape == b.shape and torch.equal(a, b), k
    # optimizer parameter groups are formed by name (slowfast/models/optimizer.py:19-21): same names, same order
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    with pytest.raises(RuntimeError):                       # CPU tensors: no fallback, loud
        ours(torch.rand(1, 2, 3, 32, 32), 1.0, False)


def _steve_pair(dev, **kw):
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    ref = steve.STEVE(LR.steve_config(**kw)).to(dev)
    ours = _swap(copy.deepcopy(ref)).to(dev)
    return ref, ours


def _train_step(model, video, seed):
    """The body of slot_train_epoch (tools/steve_train_net.py:95-126): forward(video, tau, hard), loss = mse + cross_entropy,
    backward, clip_grad_norm_ over all parameters."""
    torch.manual_seed(seed)                                   # gumbel noise, slot noise, dropout: same draws in both models
    for p in model.parameters():
        p.grad = None
    recon, ce, mse, attns = model(video, 1.0, False)
    loss = mse + ce
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1e9)
    return recon, float(ce), float(mse), attns, float(gn)


@needs_ref
@pytest.mark.gpu
def test_steve_training_step_with_swapped_savi_fp32():
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref, ours = _steve_pair(dev)
    video = torch.rand(3, 3, 3, 32, 32, generator=torch.Generator().manual_seed(5)).to(dev)
    r = _train_step(ref, video, 7)
    o = _train_step(ours, video, 7)
    assert abs(o[1] - r[1]) <= 1e-4 * abs(r[1]) and abs(o[2] - r[2]) <= 1e-4 * abs(r[2]), (o[1:3], r[1:3])
    assert err(o[3].detach().cpu().numpy(), r[3].detach().cpu().numpy()) < 1e-4            # attention overlays
    assert abs(o[4] - r[4]) <= 1e-3 * r[4]                                                   # global gradient norm
    gr = dict(ref.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    worst, worst_name = 0.0, None
    for n, p in ours.named_parameters():
        if not p.requires_grad:                                                               # steve_encoder.pos.pe is a frozen grid
            continue
        assert p.grad is not None, n                                                          # clip_grad_norm_ needs every .grad
        g = gr[n].grad
        den = float(g.abs().max())
        if den <= 1e-6 * gmax:                      # structurally zero (savi.norm_slots.bias): absolute bound instead
            den = gmax
        e = float((p.grad - g).abs().max()) / den
        if e > worst:
            worst, worst_name = e, n
    # whole-model gradients through the swapped module (cuDNN convolutions and fp32 eager everywhere else), every tensor by its own max
    assert worst < 2e-3, (worst_name, worst)


@needs_ref
@pytest.mark.gpu
def test_steve_encode_with_swapped_savi_is_argmax_identical():
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, ours = _steve_pair(dev)
    ref.eval(); ours.eval()
    video = torch.rand(2, 4, 3, 32, 32, generator=torch.Generator().manual_seed(6)).to(dev)
    with torch.no_grad():
        torch.manual_seed(3); s_r, vis_r, a_r = ref.encode(video)
        torch.manual_seed(3); s_o, vis_o, a_o = ours.encode(video)
    assert s_o.shape == s_r.shape and a_o.shape == a_r.shape and vis_o.shape == vis_r.shape
    assert err(s_o.cpu().numpy(), s_r.cpu().numpy()) < 1e-4
    # segmentation = argmax over slots of the attention maps (tools/steve_eval_net.py:93-104)
    assert torch.equal(a_o.argmax(2), a_r.argmax(2))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16])
def test_cuda_autocast_output_dtypes_match_the_reference(amp_dtype):
    """the dtypes `savi(emb_set)` returns under torch.autocast("cuda"), probed on the reference itself."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    ref = LR.reference_slot_attention_video(2, 7, 128, 128, 128, 1, 4, 0.0).to(dev)
    from focus_b200 import SlotAttentionVideo
    ours = SlotAttentionVideo(2, 7, 128, 128, 128, 1, 4, 0.0).to(dev)
    ours.load_state_dict(ref.state_dict(), strict=True)
    for in_dtype in (amp_dtype, torch.float32):           # emb_set leaves an autocast nn.Linear (half); fp32 for completeness
        x = torch.randn(2, 3, 256, 128, device=dev).to(in_dtype)
        with torch.autocast("cuda", dtype=amp_dtype):
            torch.manual_seed(1); s_r, a_r = ref(x)
            torch.manual_seed(1); s_o, a_o = ours(x)
        assert (s_o.dtype, a_o.dtype) == (s_r.dtype, a_r.dtype), (in_dtype, s_o.dtype, a_o.dtype, s_r.dtype, a_r.dtype)
        # and the values are the same computation within the half-precision class of the reference's own run
        assert err(a_o.detach().float().cpu().numpy(), a_r.detach().float().cpu().numpy()) < 0.25
