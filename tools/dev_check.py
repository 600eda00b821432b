"""Bring-up helper (GPU box): CUDA path vs the numpy fp64 oracle on the golden fixtures, printing every error."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from focus_b200 import SlotAttentionVideo
from oracle import savi_numpy as O
from tests._util import load_fixture, err, grad_scale


def run(name, dtype=torch.float32, cluster=0, bwd=True):
    fx = load_fixture(name)
    m = SlotAttentionVideo(fx["I"], fx["K"], fx["D"], fx["Ds"], fx["M"], fx["blocks"], fx["heads"], 0.0).cuda()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in fx["params"].items()})
    m.cluster = cluster
    x = torch.from_numpy(fx["x"]).cuda().to(dtype).requires_grad_(True)
    noise = torch.from_numpy(fx["noise"]).cuda()
    # oracle on exactly the values the kernel sees (bf16-rounded inputs)
    xo = x.detach().float().cpu().numpy().astype(np.float64)
    s_ref, a_ref, sv = O.forward(fx["params"], xo, fx["noise"].astype(np.float64), fx["I"], fx["heads"], keep=True)
    t0 = time.time()
    s, a = m(x, noise=noise)
    torch.cuda.synchronize()
    print("%-7s %-8s cn=%d fwd %.1f ms  slots %.2e  attn %.2e" % (
        name, str(dtype).split('.')[-1], cluster, (time.time() - t0) * 1e3,
        err(s.detach().float().cpu().numpy(), s_ref), err(a.detach().float().cpu().numpy(), a_ref)), flush=True)
    if not bwd:
        return
    gs = torch.from_numpy(fx["g_slots"]).cuda()
    ga = None if fx["g_attn"] is None else torch.from_numpy(fx["g_attn"]).cuda().to(dtype)
    loss = (s.float() * gs).sum() + (0 if ga is None else (a.float() * ga.float()).sum())
    loss.backward()
    torch.cuda.synchronize()
    gao = None if ga is None else ga.float().cpu().numpy().astype(np.float64)
    dx_ref, G_ref, _ = O.backward(fx["params"], sv, fx["g_slots"].astype(np.float64), gao)
    gsc = grad_scale(G_ref)
    worst, wname = 0, ""
    for n, p in m.named_parameters():
        e = float(np.abs(p.grad.cpu().numpy() - G_ref[n]).max() / gsc)
        if e > worst:
            worst, wname = e, n
    print("        dx %.2e  worst param grad %.2e (%s)" % (err(x.grad.float().cpu().numpy(), dx_ref), worst, wname), flush=True)
    if "-v" in sys.argv:
        for n, p in m.named_parameters():
            print("          %-50s %.2e" % (n, float(np.abs(p.grad.cpu().numpy() - G_ref[n]).max() / gsc)))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    names = ["tiny_a", "tiny_b", "tiny_c", "tiny_d", "c1"]
    for n in names:
        for cn in ([1, 2] if n != "c1" else [1, 8]):
            run(n, torch.float32, cn)
    for n in names:
        run(n, torch.bfloat16, 0)
