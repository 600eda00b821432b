"""Development check of the tcgen05 forward kernel against the mma.sync kernels (same process, env toggle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from focus_b200 import SlotAttentionVideo

def run(B, T, N, K, I, blocks, cluster, bwd=True, seed=0):
    torch.manual_seed(seed)
    m = SlotAttentionVideo(I, K, 128, 128, 128, blocks, 4, 0.0).cuda()
    m.cluster = cluster
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, N, 128, generator=g).bfloat16().cuda().requires_grad_(True)
    noise = torch.randn(B, K, 128, generator=g).cuda()
    gs = torch.randn(B, T, K, 128, generator=g).cuda()
    ga = torch.randn(B, T, N, K, generator=g).bfloat16().cuda()
    res = {}
    for mode in ("v1", "umma"):
        if mode == "v1": os.environ["SAVI_DISABLE_UMMA"] = "1"
        else: os.environ.pop("SAVI_DISABLE_UMMA", None)
        s, at = m(x, noise=noise)
        out = [s.float().detach().clone(), at.float().detach().clone()]
        if bwd:
            torch.autograd.backward([s, at], [gs.to(s.dtype), ga])
            out.append(x.grad.float().clone()); x.grad = None
            out.append(torch.cat([p.grad.flatten() for p in m.parameters()]).clone())
            for p in m.parameters(): p.grad = None
        torch.cuda.synchronize()
        res[mode] = out
    names = ["slots", "attn", "dx", "dparams"]
    errs = []
    for n, a, b in zip(names, res["v1"], res["umma"]):
        e = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-30)
        errs.append("%s %.2e" % (n, e))
    bad = any(not torch.isfinite(t).all() for t in res["umma"])
    print("B=%d T=%d N=%d K=%d I=%d blocks=%d cluster=%d : %s %s" % (B, T, N, K, I, blocks, cluster, "  ".join(errs), "NONFINITE" if bad else ""), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        for spec in sys.argv[1:]:
            run(*[int(v) for v in spec.split(",")])
        sys.exit(0)
    stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    if stage <= 0: run(1, 1, 128, 24, 1, 0, 1)
    if stage <= 1: run(1, 1, 512, 24, 3, 0, 1)
    if stage <= 2: run(2, 1, 512, 24, 3, 0, 2)
    if stage <= 3: run(2, 3, 512, 24, 3, 1, 2)
    if stage <= 4: run(3, 2, 1000, 11, 2, 2, 0)
    if stage <= 5: run(64, 6, 1024, 24, 3, 1, 0)
