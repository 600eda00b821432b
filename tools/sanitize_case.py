"""One small forward+backward on each kernel family (for compute-sanitizer runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from focus_b200 import SlotAttentionVideo, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "umma"
cfg = {"umma": (2, 7, 128, 128, 128, 1, 4, torch.bfloat16, 2, 2, 200),
       "mma": (2, 7, 64, 64, 64, 1, 4, torch.bfloat16, 2, 2, 200),
       "simt": (2, 7, 64, 64, 64, 1, 4, torch.float32, 2, 2, 200)}[which]
I, K, D, Ds, M, blocks, heads, dt, B, T, N = cfg
torch.manual_seed(0)
m = SlotAttentionVideo(I, K, D, Ds, M, blocks, heads, 0.0).cuda()
print(which, "path", _lib.PATH_NAMES[_lib.query(m.make_shape(B, T, N, dt)).path])
x = torch.randn(B, T, N, D, device="cuda").to(dt).requires_grad_(True)
s, a = m(x)
torch.autograd.backward([s, a], [torch.randn_like(s), torch.randn_like(a)])
torch.cuda.synchronize()
print("ok", float(s.float().abs().sum()), float(x.grad.float().abs().sum()))
