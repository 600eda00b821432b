"""Pinned host -> device copy bandwidth for the C2 input batch (the floor of bench.py's e2e step)."""
import torch
n = 64 * 6 * 1024 * 128
h = torch.empty(n, dtype=torch.bfloat16).pin_memory()
d = torch.empty(n, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("H2D %.1f MB in %.3f ms = %.1f GB/s" % (n * 2 / 1e6, ms, n * 2 / ms / 1e6))
