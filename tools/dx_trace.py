"""Development: when do the CTAs of the overlapped d_inputs kernel start / pass their frame flag / finish, relative to the
backward clip kernel (globaltimer, ns)?  usage: SAVI_DX_TRACE=1 python tools/dx_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SAVI_DX_TRACE"] = "1"
import torch, bench
from focus_b200 import _lib
c = dict(bench.CONFIGS["c2"])
m = bench.make_params_like(c).cuda()
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["T"], c["N"], c["D"], generator=g).bfloat16().cuda().requires_grad_(True)
noise = torch.randn(c["B"], c["K"], c["Ds"], generator=g).cuda()
gs = torch.randn(c["B"], c["T"], c["K"], c["Ds"], generator=g).bfloat16().cuda()
ga = torch.randn(c["B"], c["T"], c["N"], c["K"], generator=g).bfloat16().cuda()
def step():
    s, at = m(x, noise=noise)
    torch.autograd.backward([s, at], [gs, ga]); x.grad = None
for _ in range(3): step()
torch.cuda.synchronize()
buf = torch.zeros(64 + 3 * 4096 + 2, dtype=torch.int64, device="cuda")
buf[64 + 3 * 4096] = 2 ** 62
_lib.lib.savi_debug_set_phase_buffer(buf.data_ptr())
step(); torch.cuda.synchronize()
_lib.lib.savi_debug_set_phase_buffer(None)
v = buf.cpu()
b0, b1 = int(v[64 + 3 * 4096]), int(v[64 + 3 * 4096 + 1])
tr = v[64:64 + 3 * 384].view(384, 3)
print("overlap %s | bwd clip kernel, first CTA start -> last CTA end: %.1f us" % ("off" if os.environ.get("SAVI_NO_OVERLAP") else "on", (b1 - b0) / 1e3))
for i in list(range(0, 384, 16)):
    s, w, e = [(int(t) - b0) / 1e3 for t in tr[i]]
    print("dx CTA %3d (frame t=%d): start %8.1f  flag passed %8.1f  end %8.1f us after the clip kernel's start" % (i, 5 - i // 64, s, w, e))
print("last dx CTA ends %.1f us after the clip kernel ends" % ((int(tr[:, 2].max()) - b1) / 1e3))
