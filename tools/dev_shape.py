"""Per-tensor errors of one random-shape case (same generator as tests/test_cuda_parity.py shape sweep)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import savi_numpy as O
from tests import test_cuda_parity as T
from tests._util import err, grad_scale

shape = tuple(int(v) for v in sys.argv[1].split(","))
dtype = torch.bfloat16 if len(sys.argv) < 3 or sys.argv[2] == "bf16" else torch.float32
B, Tt, N, D, Ds, M, K, I, blocks, heads = shape
rng = np.random.default_rng(hash(shape) % 2 ** 31)
fx = dict(B=B, T=Tt, N=N, D=D, Ds=Ds, M=M, K=K, I=I, blocks=blocks, heads=heads, sub=1,
          params={k: v.astype(np.float32) for k, v in O.random_params(K, D, Ds, M, blocks, seed=3).items()},
          x=rng.standard_normal((B, Tt, N, D)).astype(np.float32) * 1.5 + 0.3,
          noise=rng.standard_normal((B, K, Ds)).astype(np.float32),
          g_slots=rng.standard_normal((B, Tt, K, Ds)).astype(np.float32),
          g_attn=rng.standard_normal((B, Tt, N, K)).astype(np.float32))
s, a, dx, G, x64, ga64 = T._run_cuda(fx, dtype)
rs, ra, rdx, RG = T._oracle(fx, x64, ga64, token_dtype="bf16" if dtype == torch.bfloat16 else None)
print("slots %.2e attn %.2e dx %.2e" % (err(s, rs), err(a, ra), err(dx, rdx)))
gs = grad_scale(RG)
for k, g in RG.items():
    print("  %-48s %.2e   |ref|max %.2e" % (k, float(np.abs(G[k] - g).max() / gs), float(np.abs(g).max())))
