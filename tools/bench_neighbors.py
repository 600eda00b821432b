"""Measurement of the §8f neighbours (N1 token-encoder tail, N2 attention overlay, N4 FG-ARI tables) on one B200:
CUDA-event time of each kernel on MOVi-E-sized inputs (inputs > L2, 3 warm-up + 10 timed), algorithmic bytes / time against the
measured HBM peak, and the reference's own code for the same work (PyTorch eager on the GPU for N1 / N2, its CPU loop for N4).
    python tools/bench_neighbors.py > profiles/r02_neighbors.json"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from focus_b200 import neighbors
from focus_b200.slot_attention import _linear

dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10, warm=3):
    ev = []
    for i in range(warm + n):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        if i >= warm:
            ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(x.elapsed_time(y) for x, y in ev)


out = {"peak_gbs": peak}
# ---- N1: 64 clips x 6 frames, 32 x 32 tokens, C = 128 (the C2 token count) ----
BT, C, H, W = 384, 128, 32, 32
torch.manual_seed(0)
ln = torch.nn.LayerNorm(C).to(dev)
mlp = torch.nn.Sequential(_linear(C, C, weight_init="kaiming"), torch.nn.ReLU(), _linear(C, C)).to(dev)
emb = torch.randn(BT, C, H, W, device=dev)
with torch.no_grad():
    ms = timed(lambda: neighbors.token_mlp(emb, ln, mlp, out_dtype=torch.bfloat16))
    ref = timed(lambda: mlp(ln(emb.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2))))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref16 = timed(lambda: mlp(ln(emb.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2))))
by = BT * H * W * C * (4 + 2)
out["N1_token_mlp"] = {"shape": [BT, C, H, W], "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                       "reference_eager_fp32_ms": ref, "reference_eager_bf16_autocast_ms": ref16}
# ---- N2: 128 x 128 video, 64 x 64 attention grid, K = 24 (C2 slots), 96 frames (fits memory: overlay 4.5 GB) ----
BTv, K, Cv, Hv, He = 96, 24, 3, 128, 64
video = torch.rand(1, BTv, Cv, Hv, Hv, device=dev)
attn = torch.softmax(torch.randn(1, BTv, He * He, K, device=dev), -1).bfloat16()
ms = timed(lambda: neighbors.attention_overlay(video, attn, He, He))


def ref_overlay():
    a = attn.float().transpose(-1, -2).reshape(1, BTv, K, 1, He, He).repeat_interleave(Hv // He, dim=-2).repeat_interleave(Hv // He, dim=-1)
    return video.unsqueeze(2) * a + (1. - a), a


ref = timed(ref_overlay)
by = BTv * (K * Cv * Hv * Hv * 4 + K * Hv * Hv * 4 + Cv * Hv * Hv * 4 + He * He * K * 2)
out["N2_attention_overlay"] = {"shape": [BTv, K, Cv, Hv, Hv, He], "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6,
                               "frac": by / ms / 1e6 / peak, "reference_eager_ms": ref}
del video, attn
torch.cuda.empty_cache()
# ---- N4: 64 clips, 24 + 24 segments, 6 frames of 128 x 128 ----
B, N0, N1, D = 64, 24, 24, 6 * 128 * 128
g = torch.Generator().manual_seed(0)
seg = torch.randint(0, N0 + 1, (B, D), generator=g).to(dev)
true = torch.stack([(seg == i) for i in range(N0)], 1).float()
pred = torch.rand(B, N1, D, generator=g).to(dev)
ms = timed(lambda: neighbors.ari_tables(true, pred))
by = B * (N0 + N1) * D * 4
out["N4_ari_tables"] = {"shape": [B, N0, N1, D], "ms": ms, "algorithmic_bytes": by, "achieved_gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
t0 = time.perf_counter()
val = neighbors.evaluate_ari(true, pred)
torch.cuda.synchronize()
out["N4_ari_tables"]["evaluate_ari_end_to_end_ms"] = (time.perf_counter() - t0) * 1e3
try:
    from oracle import _load_reference as LR
    M = LR.load_reference_metrics()
    tc, pc = true[:4].cpu(), pred[:4].cpu()
    t0 = time.perf_counter()
    M.evaluate_ari(tc, pc)
    out["N4_ari_tables"]["reference_cpu_ms_per_64_clips"] = (time.perf_counter() - t0) * 1e3 * (B / 4)
    out["N4_ari_tables"]["reference_note"] = "slowfast/utils/metrics.py:evaluate_ari on 4 of the 64 clips (host CPU), scaled x16"
except Exception as e:
    out["N4_ari_tables"]["reference_error"] = str(e)[:200]
print(json.dumps(out, indent=1))
