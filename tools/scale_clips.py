"""Kernel times of the C2 shape at several clip counts (is the clip kernel bound by shared L2 bandwidth?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from focus_b200 import _lib
c = dict(bench.CONFIGS["c2"])
m = bench.make_params_like(c).cuda()
_lib.profile_enable(True)
for B in [int(v) for v in (sys.argv[1:] or ["8", "16", "32", "64", "74"])]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, c["T"], c["N"], c["D"], generator=g).bfloat16().cuda().requires_grad_(True)
    noise = torch.randn(B, c["K"], c["Ds"], generator=g).cuda()
    gs = torch.randn(B, c["T"], c["K"], c["Ds"], generator=g).cuda()
    ga = torch.randn(B, c["T"], c["N"], c["K"], generator=g).bfloat16().cuda()
    for _ in range(3):
        s, at = m(x, noise=noise)
        torch.autograd.backward([s, at], [gs.to(s.dtype), ga]); x.grad = None
    torch.cuda.synchronize()
    print(B, {k: round(v, 3) for k, v in _lib.profile_read().items()}, flush=True)
