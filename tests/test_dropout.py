"""Predictor dropout in training mode (reference transformer.py:12-13, 44, 48, 68; constructor default 0.1, steve.py:16).

  * CPU: the oracle's dropout semantics (three multiplicative masks per block, forward and closed-form backward) against the
    UNMODIFIED reference module in train mode, whose F.dropout calls are replaced by a mask-replaying stand-in (so both sides use
    the same masks and the comparison is exact, fp64);
  * GPU: the three kernel families against the oracle with injected masks;
  * GPU: the default path draws its masks with the reference's own RNG consumption — with equal seeds the CUDA module in
    train mode reproduces the reference module in train mode on the same GPU (fp32, 1e-5 class), outputs and gradients.
"""
import numpy as np
import pytest
import torch

from oracle import _load_reference as LR
from oracle import savi_numpy as O
from tests._util import TOL_BF16, TOL_FP32, err, grad_errs, l2_err, relu_fed

needs_ref = pytest.mark.skipif(not LR.reference_available(), reason="reference sources not present (run oracle/make_ref.py)")


def _masks(rng, T, blocks, B, H, K, Ds, p):
    Sp = (T - 1) * blocks
    mk = lambda *s: ((rng.random(s) >= p).astype(np.float64) / (1.0 - p))
    return mk(Sp, B, H, K, K), mk(Sp, B, K, Ds), mk(Sp, B, K, Ds)


def _flat(masks):
    return torch.from_numpy(np.concatenate([m.reshape(-1) for m in masks]).astype(np.float32))


def _problem(seed, B, T, N, D, Ds, M, K, I, blocks, heads):
    rng = np.random.default_rng(seed)
    return dict(B=B, T=T, N=N, D=D, Ds=Ds, M=M, K=K, I=I, blocks=blocks, heads=heads,
                params={k: v.astype(np.float32) for k, v in O.random_params(K, D, Ds, M, blocks, seed=seed).items()},
                x=rng.standard_normal((B, T, N, D)).astype(np.float32), noise=rng.standard_normal((B, K, Ds)).astype(np.float32),
                g_slots=rng.standard_normal((B, T, K, Ds)).astype(np.float32), g_attn=rng.standard_normal((B, T, N, K)).astype(np.float32)), rng


@needs_ref
@pytest.mark.parametrize("cfg", [(2, 3, 40, 16, 16, 24, 5, 2, 2, 4), (1, 2, 33, 24, 16, 20, 3, 1, 1, 2)])
def test_oracle_dropout_equals_reference_train_mode_cpu(cfg):
    B, T, N, D, Ds, M, K, I, blocks, heads = cfg
    p = 0.25
    fx, rng = _problem(21, *cfg)
    masks = _masks(rng, T, blocks, B, heads, K, Ds, p)
    ref = LR.reference_slot_attention_video(I, K, D, Ds, M, blocks, heads, p).double().train()
    ref.load_state_dict({k: torch.from_numpy(v).double() for k, v in fx["params"].items()})
    # the reference's draw order: every frame (also the last, whose result it discards), every block: attention, proj_o, FFN
    queue = []
    for t in range(T):
        for j in range(blocks):
            f = j * (T - 1) + t
            for m, shp in zip(masks, ((B, heads, K, K), (B, K, Ds), (B, K, Ds))):
                queue.append(torch.from_numpy(m[f]) if t < T - 1 else torch.ones(shp, dtype=torch.float64))
    import torch.nn.functional as Fn
    orig_drop, orig_norm = Fn.dropout, torch.Tensor.normal_
    calls = []

    def replay(inp, p_=0.5, training=True, inplace=False):
        assert training and abs(p_ - p) < 1e-12
        m = queue[len(calls)]
        calls.append(tuple(inp.shape))
        assert tuple(inp.shape) == tuple(m.shape)
        return inp * m

    x = torch.from_numpy(fx["x"]).double().requires_grad_(True)
    Fn.dropout = replay
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(torch.from_numpy(fx["noise"]).double())
    try:
        slots, attn = ref(x)
    finally:
        Fn.dropout, torch.Tensor.normal_ = orig_drop, orig_norm
    assert len(calls) == 3 * blocks * T
    ((slots * torch.from_numpy(fx["g_slots"]).double()).sum() + (attn * torch.from_numpy(fx["g_attn"]).double()).sum()).backward()
    s, a, sv = O.forward(fx["params"], fx["x"].astype(np.float64), fx["noise"].astype(np.float64), I, heads, keep=True, drop_masks=masks)
    dx, G, _ = O.backward(fx["params"], sv, fx["g_slots"].astype(np.float64), fx["g_attn"].astype(np.float64))
    assert err(s, slots.detach().numpy()) < 1e-12 and err(a, attn.detach().numpy()) < 1e-12
    assert err(dx, x.grad.numpy()) < 1e-11
    RG = {n: (q.grad.numpy() if q.grad is not None else np.zeros(q.shape)) for n, q in ref.named_parameters()}
    assert max(e for e, _ in grad_errs(G, RG).values()) < 1e-10
    # and the masks matter: without them the outputs differ
    s0, _ = O.forward(fx["params"], fx["x"].astype(np.float64), fx["noise"].astype(np.float64), I, heads)
    assert err(s0, s) > 1e-3


CASES = [
    (torch.float32, (2, 3, 200, 64, 64, 128, 7, 3, 1, 4)),          # SIMT fp32 family
    (torch.float32, (1, 3, 96, 32, 32, 32, 11, 2, 2, 2)),           # two blocks
    (torch.bfloat16, (3, 3, 384, 128, 128, 128, 24, 3, 1, 4)),      # tcgen05 family, CTA pair
    (torch.bfloat16, (2, 4, 1032, 128, 128, 128, 11, 2, 2, 8)),     # tcgen05, 2 blocks x 8 heads
    (torch.bfloat16, (2, 3, 304, 192, 192, 192, 15, 2, 1, 4)),      # mma.sync family (the shipped yaml dims)
]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,cfg", CASES)
def test_cuda_dropout_against_oracle_with_injected_masks(dtype, cfg):
    from focus_b200 import SlotAttentionVideo, _lib
    B, T, N, D, Ds, M, K, I, blocks, heads = cfg
    p = 0.2
    fx, rng = _problem(33, *cfg)
    masks = _masks(rng, T, blocks, B, heads, K, Ds, p)
    m = SlotAttentionVideo(I, K, D, Ds, M, blocks, heads, p).cuda().train()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in fx["params"].items()}, strict=True)
    path = _lib.query(m.make_shape(B, T, N, dtype)).path
    x = torch.from_numpy(fx["x"]).cuda().to(dtype).requires_grad_(True)
    s, a = m(x, noise=torch.from_numpy(fx["noise"]).cuda(), dropout_masks=_flat(masks))
    ga = torch.from_numpy(fx["g_attn"]).cuda().to(dtype)
    torch.autograd.backward([s, a], [torch.from_numpy(fx["g_slots"]).cuda().to(s.dtype), ga])
    torch.cuda.synchronize()
    x64, ga64 = x.detach().double().cpu().numpy(), ga.double().cpu().numpy()
    q = {} if dtype == torch.float32 else dict(token_dtype="bf16", weight_dtype="f16" if path == 2 else None)
    rs, ra, sv = O.forward(fx["params"], x64, fx["noise"].astype(np.float64), I, heads, keep=True, drop_masks=masks, **q)
    rdx, RG, _ = O.backward(fx["params"], sv, fx["g_slots"].astype(np.float64), ga64,
                            bwd_weight_dtype="bf16" if (dtype != torch.float32 and path == 2) else None)
    tol = TOL_FP32 if dtype == torch.float32 else TOL_BF16
    assert err(s.detach().float().cpu().numpy(), rs) < tol and err(a.detach().float().cpu().numpy(), ra) < tol
    assert err(x.grad.float().cpu().numpy(), rdx) < tol
    G = {n: q_.grad.detach().cpu().numpy() for n, q_ in m.named_parameters()}
    bad = {}
    for k, (e, z) in grad_errs(G, RG).items():
        if tol > 1e-3 and relu_fed(k) and not z:
            if not (l2_err(G[k], RG[k]) < tol and e < 3 * tol):
                bad[k] = (e, l2_err(G[k], RG[k]))
        elif not e < tol:
            bad[k] = e
    assert not bad, bad
    # the masks were applied: evaluation mode (no dropout) gives different slots from the second frame on
    with torch.no_grad():
        s_eval, _ = m.eval()(x.detach(), noise=torch.from_numpy(fx["noise"]).cuda())
    assert err(s_eval.float().cpu().numpy()[:, 1:], rs[:, 1:]) > 1e-3


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(2, 3, 256, 64, 64, 128, 7, 2, 1, 4), (3, 4, 128, 32, 32, 32, 5, 3, 2, 2)])
def test_default_dropout_draw_matches_reference_rng_on_gpu(cfg):
    """Same seed, same GPU, train mode, dropout 0.1 (the reference's constructor default), fp32: the module's own mask draw
    consumes the generator exactly as the reference does, so outputs AND gradients agree to the fp32 class."""
    from focus_b200 import SlotAttentionVideo
    B, T, N, D, Ds, M, K, I, blocks, heads = cfg
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    ref = LR.reference_slot_attention_video(I, K, D, Ds, M, blocks, heads, 0.1).to(dev).train()
    ours = SlotAttentionVideo(I, K, D, Ds, M, blocks, heads, 0.1).to(dev).train()
    ours.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, N, D, generator=g).to(dev)
    gs = torch.randn(B, T, K, Ds, generator=g).to(dev)
    out = []
    for mod in (ref, ours):
        xx = x.clone().requires_grad_(True)
        torch.manual_seed(99)
        s, a = mod(xx)
        (s * gs).sum().backward()
        out.append((s.detach().cpu().numpy(), a.detach().cpu().numpy(), xx.grad.cpu().numpy(),
                    {n: q.grad.detach().cpu().numpy() for n, q in mod.named_parameters()}, torch.cuda.get_rng_state(dev).clone()))
    (rs, ra, rdx, RG, rng_r), (s, a, dx, G, rng_o) = out
    assert torch.equal(rng_r, rng_o)                               # the generator ends in the same state: same consumption
    assert err(s, rs) < 1e-4 and err(a, ra) < 1e-4 and err(dx, rdx) < 1e-4      # fp32 eager (cuBLAS) vs fp32 kernels, chaotic map
    # per tensor; a gradient that is structurally zero (norm_slots.bias) is fp32 rounding noise on BOTH sides here: absolute bound
    gmax = max(float(np.abs(g_).max()) for g_ in RG.values())
    worst = {}
    for n, g_ in RG.items():
        den = float(np.abs(g_).max())
        worst[n] = float(np.abs(G[n] - g_).max()) / (den if den > 1e-5 * gmax else gmax)
    assert max(worst.values()) < 1e-3, max(worst.items(), key=lambda kv: kv[1])
    torch.manual_seed(99)
    with torch.no_grad():
        s_eval, _ = ours.eval()(x)
    assert err(s_eval.cpu().numpy()[:, 1:], rs[:, 1:]) > 1e-3       # and dropout really acted
