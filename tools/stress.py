"""Stability check on one B200: many steps of C2 back to back (outputs and gradients must stay bit-identical from step to
step: the forward is deterministic, the backward's atomics only reorder fp32 sums), then a large-N clip against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from focus_b200 import SlotAttentionVideo
from oracle import savi_numpy as O

c = dict(bench.CONFIGS["c2"])
m = bench.make_params_like(c).cuda()
g = torch.Generator().manual_seed(1)
x = torch.randn(c["B"], c["T"], c["N"], c["D"], generator=g).bfloat16().cuda().requires_grad_(True)
noise = torch.randn(c["B"], c["K"], c["Ds"], generator=g).cuda()
gs = torch.randn(c["B"], c["T"], c["K"], c["Ds"], generator=g).bfloat16().cuda()
ga = torch.randn(c["B"], c["T"], c["N"], c["K"], generator=g).bfloat16().cuda()
ref = None
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
worst = 0.0
for i in range(steps):
    m.zero_grad(set_to_none=True); x.grad = None
    s, a = m(x, noise=noise)
    torch.autograd.backward([s, a], [gs, ga])
    if i % 50 == 0 or i == steps - 1:
        cur = (s.detach().clone(), a.detach().clone(), x.grad.clone(), m.gru.weight_hh.grad.clone())
        if ref is None:
            ref = cur
        else:
            assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]), "forward not reproducible at step %d" % i
            assert torch.equal(cur[2], ref[2]), "d_inputs not reproducible at step %d" % i
            worst = max(worst, float((cur[3] - ref[3]).abs().max() / ref[3].abs().max()))
torch.cuda.synchronize()
print("C2 x %d steps: forward and d_inputs bit-identical; weight-gradient drift (atomic order) %.2e" % (steps, worst))

# one large-N clip on the tcgen05 path against the fp64 oracle
I, K, D, N, T = 3, 24, 128, 16384, 2
torch.manual_seed(3)
mm = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
rng = np.random.default_rng(5)
xs = torch.from_numpy(rng.standard_normal((1, T, N, D)).astype(np.float32)).bfloat16().cuda().requires_grad_(True)
nz = torch.from_numpy(rng.standard_normal((1, K, D)).astype(np.float32)).cuda()
gsl = torch.from_numpy(rng.standard_normal((1, T, K, D)).astype(np.float32)).cuda()
s, a = mm(xs, noise=nz)
(s.float() * gsl).sum().backward()
P = {k: v.detach().cpu().numpy() for k, v in mm.state_dict().items()}
x64 = xs.detach().float().cpu().numpy().astype(np.float64)
rs, ra, sv = O.forward(P, x64, nz.cpu().numpy().astype(np.float64), I, 4, keep=True, token_dtype="bf16", weight_dtype="f16")
dx, G, _ = O.backward(P, sv, gsl.cpu().numpy().astype(np.float64), None, bwd_weight_dtype="bf16")
es = O.max_norm_err(s.detach().float().cpu().numpy(), rs); ea = O.max_norm_err(a.detach().float().cpu().numpy(), ra)
ed = O.max_norm_err(xs.grad.float().cpu().numpy(), dx)
gmax = max(float(np.abs(v).max()) for v in G.values())
eg = max(float(np.abs(p.grad.cpu().numpy() - G[n]).max()) / gmax for n, p in mm.named_parameters())
print("N=16384 clip vs oracle (at the quantisation points: bf16 tokens, fp16 / bf16 weight images): slots %.2e attn %.2e d_inputs %.2e param grads %.2e (bar 2e-2)" % (es, ea, ed, eg))
assert max(es, ea, ed, eg) < 2e-2
