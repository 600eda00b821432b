"""One (or a few) forward+backward steps of a bench config — the command profiled under ncu."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--clips", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--fwd-only", action="store_true")
a = ap.parse_args()
c = dict(bench.CONFIGS[a.config])
if a.clips:
    c["B"] = a.clips
dt = torch.float32 if c["dtype"] == "fp32" else torch.bfloat16
m = bench.make_params_like(c).cuda()
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["T"], c["N"], c["D"], generator=g).to(dt).cuda().requires_grad_(True)
noise = torch.randn(c["B"], c["K"], c["Ds"], generator=g).cuda()
gs = torch.randn(c["B"], c["T"], c["K"], c["Ds"], generator=g).cuda()
ga = torch.randn(c["B"], c["T"], c["N"], c["K"], generator=g).to(dt).cuda()
for _ in range(a.steps):
    s, at = m(x, noise=noise)
    if not a.fwd_only:
        torch.autograd.backward([s, at], [gs.to(s.dtype), ga])
    x.grad = None
torch.cuda.synchronize()
print("ok")
