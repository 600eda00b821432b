"""bf16 (tensor-core path) vs oracle at the token quantisation point; prints all errors."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from focus_b200 import SlotAttentionVideo, _lib
from oracle import savi_numpy as O
from tests._util import load_fixture, err, grad_scale

def run(name, cluster=0):
    fx = load_fixture(name)
    m = SlotAttentionVideo(fx["I"], fx["K"], fx["D"], fx["Ds"], fx["M"], fx["blocks"], fx["heads"], 0.0).cuda()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in fx["params"].items()})
    m.cluster = cluster
    x = torch.from_numpy(fx["x"]).cuda().bfloat16().requires_grad_(True)
    noise = torch.from_numpy(fx["noise"]).cuda()
    xo = x.detach().float().cpu().numpy().astype(np.float64)
    s, a = m(x, noise=noise)
    gs = torch.from_numpy(fx["g_slots"]).cuda()
    ga = None if fx["g_attn"] is None else torch.from_numpy(fx["g_attn"]).cuda().bfloat16()
    torch.autograd.backward([s, a] if ga is not None else [s], [gs.to(s.dtype), ga] if ga is not None else [gs.to(s.dtype)])
    torch.cuda.synchronize()
    gao = None if ga is None else ga.float().cpu().numpy().astype(np.float64)
    for td in (None, "bf16"):
        s_ref, a_ref, sv = O.forward(fx["params"], xo, fx["noise"].astype(np.float64), fx["I"], fx["heads"], keep=True, token_dtype=td)
        dx_ref, G_ref, _ = O.backward(fx["params"], sv, fx["g_slots"].astype(np.float64), gao)
        gsc = grad_scale(G_ref)
        worst = max((float(np.abs(p.grad.cpu().numpy() - G_ref[n]).max() / gsc), n) for n, p in m.named_parameters())
        print("%-7s cn=%d oracle[%s]: slots %.2e attn %.2e dx %.2e grads %.2e (%s)" % (name, cluster, td,
              err(s.detach().float().cpu().numpy(), s_ref), err(a.detach().float().cpu().numpy(), a_ref),
              err(x.grad.float().cpu().numpy(), dx_ref), worst[0], worst[1]), flush=True)

print(torch.cuda.get_device_name(0))
for n in ["tiny_a", "tiny_d", "c1"]:
    for cn in ([1, 2] if n != "c1" else [0, 1]):
        run(n, cn)
