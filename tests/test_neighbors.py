"""SURVEY.md §8f N2 (attention-map consumer) and N4 (FG-ARI): oracle pinned to the reference's own code on CPU; the CUDA
kernels against the oracle (and the reference) on the GPU.  Integer work (the contingency tables) must be exact; the overlay
is fp32 elementwise arithmetic in the reference's operation order and must be bit-identical."""
import os

import numpy as np
import pytest
import torch

from oracle import _load_reference as LR
from oracle import neighbors_numpy as ON
from tests._util import ROOT

needs_ref = pytest.mark.skipif(not LR.reference_available(), reason="reference sources not present (run oracle/make_ref.py)")


def _masks(B, N0, N1, D, seed, ties=False):
    rng = np.random.default_rng(seed)
    seg = rng.integers(0, N0 + 1, size=(B, D))                       # ground-truth segment id per pixel; N0 = background (dropped)
    true = np.stack([(seg == i) for i in range(N0)], 1).astype(np.float32)
    pred = rng.random((B, N1, D)).astype(np.float32)
    if ties:                                                          # exact ties: the FIRST maximum must win (torch.argmax)
        pred = np.round(pred * 4) / 4
        true[:, 0, ::7] = 2.0                                         # .byte() & 1 -> 0
        true[:, 1, ::5] = 0.5                                         # truncates to 0
        true[:, 2, ::11] = 3.0                                        # low bit 1
    return true, pred


# ---- CPU: the oracle against the reference ------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("case", [(3, 5, 7, 1000, False), (2, 24, 15, 4096, True), (1, 1, 1, 17, False), (2, 3, 4, 64, True)])
def test_oracle_ari_equals_reference_metrics(case):
    B, N0, N1, D, ties = case
    M = LR.load_reference_metrics()
    true, pred = _masks(B, N0, N1, D, seed=sum(case[:4]), ties=ties)
    want = M.evaluate_ari(torch.from_numpy(true), torch.from_numpy(pred))
    got = ON.evaluate_ari(true, pred)
    assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (got, want)
    # and the table itself against compute_mask_ari's construction (metrics.py:50-57)
    onehot = np.zeros_like(pred)
    arg = torch.from_numpy(pred).argmax(1).numpy()
    for b in range(B):
        onehot[b, arg[b], np.arange(D)] = 1.0
        m0 = torch.from_numpy(true[b])[:, None].byte()
        m1 = torch.from_numpy(onehot[b])[None, :].byte()
        assert np.array_equal((m0 & m1).sum(-1).numpy(), ON.ari_tables(true, pred)[b])


def test_oracle_ari_perfect_and_degenerate_tables():
    t = np.eye(4) * 10
    assert ON.ari_from_table(t) == 1.0
    assert ON.ari_from_table(np.array([[5.0]])) == 1.0               # one segment each: the reference's "perfect case" branch
    true, pred = _masks(2, 4, 4, 256, 0)
    assert abs(ON.evaluate_ari(true, true[:, :4] + 0.0) - 1.0) < 1e-12


@needs_ref
def test_oracle_overlay_equals_reference_encode_cpu():
    """STEVE.encode (steve.py:331-357) on CPU: its attns_vis / attns outputs against the oracle applied to the attention
    maps its own slot-attention module returned."""
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    model = steve.STEVE(LR.steve_config()).eval()
    video = torch.rand(2, 3, 3, 32, 32)
    grabbed = {}
    h = model.steve_encoder.savi.register_forward_hook(lambda m, i, o: grabbed.update(attn=o[1].detach()))
    with torch.no_grad():
        _, vis, up = model.encode(video)
    h.remove()
    o_vis, o_up = ON.attention_overlay(video.numpy(), grabbed["attn"].numpy(), 16, 16)
    assert np.array_equal(o_up, up.numpy()) and np.array_equal(o_vis, vis.numpy())


@needs_ref
def test_oracle_token_mlp_equals_reference_modules_cpu():
    """steve.py:307-308 executed with the reference's own STEVEEncoder modules against the oracle (fp64)."""
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    enc = steve.STEVE(LR.steve_config()).steve_encoder.double()
    with torch.no_grad():
        for p_ in list(enc.layer_norm.parameters()) + [enc.mlp[0].bias, enc.mlp[2].bias]:
            p_.add_(0.3 * torch.randn_like(p_))
        emb = torch.randn(3, 128, 16, 16, dtype=torch.float64)
        want = enc.mlp(enc.layer_norm(emb.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)))
    got = ON.token_mlp(emb.numpy(), *[t.detach().numpy() for t in (enc.layer_norm.weight, enc.layer_norm.bias, enc.mlp[0].weight, enc.mlp[0].bias,
                                                                    enc.mlp[2].weight, enc.mlp[2].bias)])
    assert np.abs(got - want.numpy()).max() <= 1e-12 * np.abs(want.numpy()).max()


def test_library_exports_the_neighbour_symbols():
    import ctypes
    import re
    from focus_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "focus_steve.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(steve_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS_STEVE)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    assert all(hasattr(raw, n) for n in declared)
    with pytest.raises(RuntimeError, match="no CPU path"):
        from focus_b200 import neighbors
        neighbors.evaluate_ari(torch.zeros(1, 2, 8), torch.zeros(1, 2, 8))


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [(3, 5, 7, 1000, False), (2, 24, 15, 4096, True), (1, 1, 1, 17, False), (4, 24, 24, 6 * 64 * 64 + 3, True),
                                  (2, 64, 64, 5000, True)])
def test_cuda_ari_tables_are_exact(case):
    from focus_b200 import neighbors
    B, N0, N1, D, ties = case
    true, pred = _masks(B, N0, N1, D, seed=sum(case[:4]), ties=ties)
    tabs = neighbors.ari_tables(torch.from_numpy(true).cuda(), torch.from_numpy(pred).cuda()).cpu().numpy()
    assert np.array_equal(tabs, ON.ari_tables(true, pred))
    got = neighbors.evaluate_ari(torch.from_numpy(true).cuda(), torch.from_numpy(pred).cuda())
    assert abs(got - ON.evaluate_ari(true, pred)) <= 1e-12
    if LR.reference_available():
        want = LR.load_reference_metrics().evaluate_ari(torch.from_numpy(true), torch.from_numpy(pred))
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


@pytest.mark.gpu
def test_cuda_ari_full_size_properties():
    """MOVi-E evaluation size (tools/steve_eval_net.py:107): 64 clips x 24 gt segments x 6 frames of 128 x 128 pixels."""
    from focus_b200 import neighbors
    B, N0, N1, D = 64, 24, 24, 6 * 128 * 128
    g = torch.Generator().manual_seed(0)
    seg = torch.randint(0, N0 + 1, (B, D), generator=g).cuda()
    true = torch.stack([(seg == i) for i in range(N0)], 1).float()
    pred = torch.rand(B, N1, D, generator=g).cuda()
    tabs = neighbors.ari_tables(true, pred)
    assert torch.equal(tabs.sum((1, 2)), (seg < N0).sum(1).int())                 # every foreground pixel counted exactly once
    assert torch.equal(tabs.sum(1), torch.stack([(pred.argmax(1) == j) & (seg < N0) for j in range(N1)], 1).sum(2).int())
    onehot = torch.nn.functional.one_hot(seg.clamp(max=N0 - 1), N0).permute(0, 2, 1).float()
    assert abs(neighbors.evaluate_ari(onehot, onehot * 2.0 + torch.rand_like(onehot) * 0.5) - 1.0) < 1e-12   # same partition -> ARI 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 3, 7, 3, 32, 32, 16, 16, torch.float32), (1, 2, 24, 3, 128, 128, 64, 64, torch.bfloat16),
                                  (3, 1, 15, 3, 64, 64, 64, 64, torch.float32), (2, 2, 5, 1, 16, 32, 4, 8, torch.bfloat16),
                                  (1, 1, 64, 4, 8, 8, 8, 4, torch.float32)])
def test_cuda_overlay_is_bit_identical(case):
    from focus_b200 import neighbors
    B, T, K, C, H, W, He, We, adt = case
    g = torch.Generator().manual_seed(sum(case[:8]))
    video = torch.rand(B, T, C, H, W, generator=g)
    attn = torch.softmax(torch.randn(B, T, He * We, K, generator=g) * 3, -1).to(adt)
    vis, up = neighbors.attention_overlay(video.cuda(), attn.cuda(), He, We)
    o_vis, o_up = ON.attention_overlay(video.numpy(), attn.float().numpy(), He, We)
    assert np.array_equal(up.cpu().numpy(), o_up) and np.array_equal(vis.cpu().numpy(), o_vis)
    only_up = neighbors.attention_overlay(video.cuda(), attn.cuda(), He, We, want_overlay=False)
    assert only_up[0] is None and torch.equal(only_up[1], up)


@needs_ref
@pytest.mark.gpu
def test_cuda_overlay_equals_reference_encode_on_gpu():
    from focus_b200 import neighbors
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = steve.STEVE(LR.steve_config()).to(dev).eval()
    video = torch.rand(2, 3, 3, 32, 32, device=dev)
    grabbed = {}
    h = model.steve_encoder.savi.register_forward_hook(lambda m, i, o: grabbed.update(attn=o[1].detach()))
    with torch.no_grad():
        _, vis, up = model.encode(video)
    h.remove()
    o_vis, o_up = neighbors.attention_overlay(video, grabbed["attn"], 16, 16)
    assert torch.equal(o_up, up) and torch.equal(o_vis, vis)


def _encoder_tail(C, seed):
    from focus_b200.slot_attention import _linear
    torch.manual_seed(seed)
    ln = torch.nn.LayerNorm(C)
    mlp = torch.nn.Sequential(_linear(C, C, weight_init="kaiming"), torch.nn.ReLU(), _linear(C, C))     # steve.py:224-227
    with torch.no_grad():
        for p_ in list(ln.parameters()) + [mlp[0].bias, mlp[2].bias]:
            p_.add_(0.3 * torch.randn_like(p_))
    return ln, mlp


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(6, 128, 32, 32, torch.bfloat16), (2, 128, 32, 32, torch.float32), (3, 64, 16, 16, torch.bfloat16),
                                  (2, 192, 64, 64, torch.bfloat16), (5, 128, 9, 7, torch.bfloat16), (1, 64, 4, 5, torch.float32)])
def test_cuda_token_mlp_against_oracle(case):
    from focus_b200 import neighbors
    from tests._util import err
    BT, C, H, W, odt = case
    ln, mlp = _encoder_tail(C, seed=C + H)
    emb = torch.randn(BT, C, H, W, generator=torch.Generator().manual_seed(BT)) * 1.5 + 0.2
    with torch.no_grad():
        got = neighbors.token_mlp(emb.cuda(), ln.cuda(), mlp.cuda(), out_dtype=odt)
    assert got.shape == (BT, H * W, C) and got.dtype == odt
    prm = [t.detach().cpu().numpy() for t in (ln.weight, ln.bias, mlp[0].weight, mlp[0].bias, mlp[2].weight, mlp[2].bias)]
    exact = ON.token_mlp(emb.numpy(), *prm)
    quant = ON.token_mlp(emb.numpy(), *prm, operand_dtype="bf16")
    g = got.float().cpu().numpy()
    assert err(g, exact) < 2e-2                                           # accuracy tier: the bf16 bar against the unquantised oracle
    # implementation tier: operands rounded as the kernel rounds them (a LayerNorm output or hidden activation that sits on a
    # bf16 rounding boundary may round the other way in fp32 than in fp64: a few 2^-9 operand flips, ~1e-3 on the output)
    assert err(g, quant) < (6e-3 if odt == torch.bfloat16 else 2e-3)


@needs_ref
@pytest.mark.gpu
def test_cuda_token_mlp_in_the_reference_encode_path():
    """The reference's STEVE.encode up to `emb_set` (steve.py:336-344) against the fused kernel fed with the same CNN map."""
    from focus_b200 import neighbors
    from tests._util import err
    steve = LR.load_reference_steve()
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    model = steve.STEVE(LR.steve_config()).to(dev).eval()
    video = torch.rand(2, 3, 3, 32, 32, device=dev)
    with torch.no_grad():
        emb = model.steve_encoder.pos(model.steve_encoder.cnn(video.flatten(end_dim=1)))
        want = model.steve_encoder.mlp(model.steve_encoder.layer_norm(emb.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)))
        got = neighbors.token_mlp(emb, model.steve_encoder.layer_norm, model.steve_encoder.mlp, out_dtype=torch.float32)
    assert err(got.cpu().numpy(), want.cpu().numpy()) < 2e-2
    with pytest.raises(RuntimeError, match="forward-only"):
        neighbors.token_mlp(emb, model.steve_encoder.layer_norm, model.steve_encoder.mlp)
