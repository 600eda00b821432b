import os, sys
sys.path.insert(0, '/root/repo')
import torch
from focus_b200 import neighbors
from focus_b200.slot_attention import _linear
dev = torch.device("cuda", 0)
BT, C, H, W = 384, 128, 32, 32
torch.manual_seed(0)
ln = torch.nn.LayerNorm(C).to(dev)
mlp = torch.nn.Sequential(_linear(C, C, weight_init="kaiming"), torch.nn.ReLU(), _linear(C, C)).to(dev)
emb = torch.randn(BT, C, H, W, device=dev)
with torch.no_grad():
    for _ in range(5):
        out = neighbors.token_mlp(emb, ln, mlp, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok", out.shape)
