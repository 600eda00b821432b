"""Generate tests/golden/*.npz from the UNMODIFIED reference module.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these
fixtures are what pins oracle/savi_numpy.py and oracle/savi_torch.py.  Each
fixture holds the reference's parameters, the inputs, the injected slot noise,
upstream gradients, and the reference's outputs / autograd gradients in fp64
(`*_f64`), plus — for the C1 anchor — the reference's own fp32 forward as run
"as-is" on CPU (`*_f32`).  Inputs come from numpy's PCG64 so they are
platform-independent.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle._load_reference import reference_slot_attention_video  # noqa: E402


def run_reference(ref, x, noise, g_slots, g_attn):
    x = x.clone().requires_grad_(True)
    orig = torch.Tensor.normal_
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(noise)   # inject the single RNG draw (steve.py:56)
    try:
        slots, attn = ref(x)
    finally:
        torch.Tensor.normal_ = orig
    loss = (slots * g_slots).sum()
    if g_attn is not None:
        loss = loss + (attn * g_attn).sum()
    for p in ref.parameters():
        p.grad = None
    loss.backward()
    grads = {n: (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in ref.named_parameters()}
    return slots.detach(), attn.detach(), x.grad, grads


def make(name, B, T, N, D, Ds, M, K, I, blocks, heads, seed, with_g_attn=True, perturb=True,
         f32_forward=False, sub=1):
    torch.manual_seed(seed)
    ref = reference_slot_attention_video(I, K, D, Ds, M, blocks, heads, 0.0)
    if perturb:   # make zero-initialised biases / unit LN weights non-trivial
        with torch.no_grad():
            for _, p in ref.named_parameters():
                if p.ndim == 1:
                    p.add_(0.2 * torch.randn_like(p))
    rng = np.random.default_rng(seed + 1000)
    x = rng.standard_normal((B, T, N, D)).astype(np.float32)
    noise = rng.standard_normal((B, K, Ds)).astype(np.float32)
    g_slots = rng.standard_normal((B, T, K, Ds)).astype(np.float32)
    g_attn = rng.standard_normal((B, T, N, K)).astype(np.float32) if with_g_attn else None
    out = dict(cfg=np.array([B, T, N, D, Ds, M, K, I, blocks, heads, seed, int(with_g_attn), sub]),
               noise=noise, g_slots=g_slots)
    if sub == 1:
        out["x"] = x
        if g_attn is not None:
            out["g_attn"] = g_attn
    for k, v in ref.state_dict().items():
        out["param/" + k] = v.numpy().copy()
    if f32_forward:
        with torch.no_grad():
            s32, a32, _, _ = (lambda r: (r[0], r[1], None, None))(
                _fwd_only(ref, torch.from_numpy(x), torch.from_numpy(noise)))
        out["slots_f32"] = s32.numpy()
        out["attn_f32"] = a32.numpy()[:, :, ::sub]
    ref64 = ref.double()
    tt = lambda a: None if a is None else torch.from_numpy(a).double()
    s, a, dx, grads = run_reference(ref64, tt(x), tt(noise), tt(g_slots), tt(g_attn))
    # fp64-computed; for the big subsampled fixture stored rounded to fp32 (6e-8, far below the 1e-5 bar)
    st = np.float64 if sub == 1 else np.float32
    out["slots_f64"] = s.numpy().astype(st)
    out["attn_f64"] = a.numpy()[:, :, ::sub].astype(st)
    out["dx_f64"] = dx.numpy()[:, :, ::sub].astype(st)
    for k, v in grads.items():
        out["grad/" + k] = v.numpy().astype(np.float64 if sub == 1 else np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "%.1f KB" % (os.path.getsize(path) / 1024))


def _fwd_only(ref, x, noise):
    orig = torch.Tensor.normal_
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(noise)
    try:
        return ref(x)
    finally:
        torch.Tensor.normal_ = orig


def make_bf16_reference(name, out_name):
    """The reference's OWN bf16 accuracy on fixture `name`: the unmodified module run (a) under
    torch.autocast(bfloat16) with fp32 parameters (the AMP flavour of tools/steve_train_net.py:95, in bf16) and
    (b) converted with .bfloat16(), each compared with the same module in fp64 on the same bf16-rounded inputs.
    Stored: the max-normalised error of every output / gradient tensor (`autocast/<tensor>`, `bf16/<tensor>`), which is
    what tests/test_cuda_parity.py bounds the CUDA bf16 path by (the BPTT gradient of this module amplifies any
    2^-9 rounding ~100x, so "within the reference's own bf16 error" is the meaningful gradient bar)."""
    import copy
    sys.path.insert(0, os.path.join(HERE, ".."))
    from _util import load_fixture, err
    fx = load_fixture(name)
    ref = reference_slot_attention_video(fx["I"], fx["K"], fx["D"], fx["Ds"], fx["M"], fx["blocks"], fx["heads"], 0.0)
    ref.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in fx["params"].items()})
    t = torch.from_numpy
    xb, noise, gs = t(fx["x"]).bfloat16(), t(fx["noise"]), t(fx["g_slots"])
    gab = t(fx["g_attn"]).bfloat16() if fx["g_attn"] is not None else None
    truth = run_reference(copy.deepcopy(ref).double(), xb.double(), noise.double(), gs.double(), None if gab is None else gab.double())
    with torch.autocast("cpu", dtype=torch.bfloat16):
        auto = run_reference(copy.deepcopy(ref), xb, noise, gs, gab)
    half = run_reference(copy.deepcopy(ref).bfloat16(), xb, noise.bfloat16(), gs.bfloat16(), gab)
    out = {}
    for tag, r in (("autocast", auto), ("bf16", half)):
        out[tag + "/slots"] = err(r[0].float().numpy(), truth[0].numpy())
        out[tag + "/attn"] = err(r[1].float().numpy(), truth[1].numpy())
        out[tag + "/d_inputs"] = err(r[2].float().numpy(), truth[2].numpy())
        for k in truth[3]:
            out[tag + "/grad/" + k] = err(r[3][k].float().numpy(), truth[3][k].numpy())
    path = os.path.join(HERE, out_name + ".npz")
    np.savez_compressed(path, **{k: np.float64(v) for k, v in out.items()})
    print(out_name, {k: "%.2e" % v for k, v in out.items() if "grad" not in k})


if __name__ == "__main__":
    if "--bf16-reference" in sys.argv:          # only the reference-bf16 accuracy record (the fixtures above stay as committed)
        make_bf16_reference("c1", "c1_refbf16")
        make_bf16_reference("tiny_a", "tiny_a_refbf16")
        sys.exit(0)
    #            B  T  N    D    Ds   M    K   I  blocks heads seed
    make("tiny_a", 2, 3, 64, 32, 32, 48, 5, 3, 1, 4, 11)
    make("tiny_b", 1, 2, 50, 24, 16, 20, 7, 2, 2, 2, 12)                       # D != Ds, 2 predictor blocks, ragged N
    make("tiny_c", 2, 1, 40, 16, 16, 16, 3, 1, 0, 1, 13, with_g_attn=False)   # I=1, T=1, no predictor blocks
    make("tiny_d", 1, 2, 33, 16, 32, 8, 2, 2, 1, 1, 14)                        # K=2 (base_lite.yaml), odd N
    # BASELINE config 1 (reference init as-is under manual_seed(0); inputs regenerated from the
    # numpy seed by the test, token-dim outputs subsampled every 8th token to keep the file small)
    make("c1", 2, 6, 1024, 128, 128, 128, 24, 3, 1, 4, 0, perturb=False, f32_forward=True, sub=8)
