"""BASELINE.json configs[4]: slot-attention microbench sweep, N x K x D, fp32 and bf16, one B200, against the k/v-bytes roofline.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.txt

Per point: B clips (64, or enough that k+v >= 512 MB where noted), T = 2 frames, 3 iterations, 1 predictor block;
3 warm-up + 5 timed forward+backward steps (CUDA events), inputs resident in HBM.  Columns: kernel family the library
dispatches to, ms/step, frames/s, and the fraction of the HBM roofline (4 I N Ds e bytes per frame at the measured peak).
"""
import argparse, itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from focus_b200 import SlotAttentionVideo, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
peak, _ = bench.measured_peaks()
T, I = 2, 3
Ns = [1024, 4096, 16384] if not a.quick else [1024, 4096]
Ks = [7, 11, 24, 32, 64] if not a.quick else [7, 24, 64]
Ds = [64, 128, 192, 256] if not a.quick else [128, 192]
print("%-5s %6s %3s %4s %5s  %-14s %9s %12s %9s" % ("dtype", "N", "K", "D", "B", "path", "ms/step", "frames/s", "roofline"))
for dt, N, K, D in itertools.product(("bf16", "fp32"), Ns, Ks, Ds):
    e = 2 if dt == "bf16" else 4
    for regime in ("B=64", "k+v>=512MB"):
        B = 64 if regime == "B=64" else max(64, -(-(512 << 20) // (T * 2 * N * D * e)))
        if regime != "B=64" and (B == 64 or a.quick):
            continue
        if B * T * N * (D * e * 3 + K * e * 2) > 60e9:       # bounded memory sweep
            continue
        tdt = torch.bfloat16 if dt == "bf16" else torch.float32
        torch.manual_seed(0)
        m = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
        path = _lib.PATH_NAMES[_lib.query(m.make_shape(B, T, N, tdt)).path]
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(B, T, N, D, generator=g, device="cuda").to(tdt).requires_grad_(True)
        noise = torch.randn(B, K, D, generator=g, device="cuda")
        gs = torch.randn(B, T, K, D, generator=g, device="cuda").to(tdt)
        ga = torch.randn(B, T, N, K, generator=g, device="cuda").to(tdt)

        def step():
            m.zero_grad(set_to_none=True)
            s, at = m(x, noise=noise)
            torch.autograd.backward([s, at], [gs, ga]); x.grad = None
        try:
            for _ in range(3):
                step()
        except RuntimeError as ex:                               # shape outside the library's limits: reported, not hidden
            print("%-5s %6d %3d %4d %5d  %-14s unsupported: %s" % (dt, N, K, D, B, path, str(ex)[:90]), flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        roof_ms = B * T * 4 * I * N * D * e / (peak * 1e9) * 1e3
        print("%-5s %6d %3d %4d %5d  %-14s %9.3f %12.0f %8.1f%%" % (dt, N, K, D, B, path, ms, B * T / (ms * 1e-3), 100 * roof_ms / ms), flush=True)
        del m, x, noise, gs, ga
        torch.cuda.empty_cache()
