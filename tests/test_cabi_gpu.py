"""GPU: properties of the C ABI that include/focus_savi.h promises — a stream capture of pack + forward + backward
replays as a CUDA graph (no allocation, no host synchronisation, shape-static launches, programmatic dependent launches
included) and reproduces the eager results bit for bit (forward) / to atomics order (parameter gradients)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _buffers(m, B, T, N, dt):
    from focus_b200 import _lib
    dev = torch.device("cuda", 0)
    shape = m.make_shape(B, T, N, dt)
    sz = _lib.query(shape)
    u8 = dict(dtype=torch.uint8, device=dev)
    K, Ds, D = m.num_slots, m.slot_size, m.input_size
    g = torch.Generator().manual_seed(3)
    buf = dict(packed=torch.empty(sz.packed_bytes, **u8), saved=torch.empty(sz.saved_bytes, **u8),
               fws=torch.empty(max(sz.fwd_ws_bytes, 16), **u8), bws=torch.empty(max(sz.bwd_ws_bytes, 16), **u8),
               x=torch.randn(B, T, N, D, generator=g).to(dt).to(dev), noise=torch.randn(B, K, Ds, generator=g).to(dev),
               slots=torch.empty(B, T, K, Ds, device=dev), attn=torch.empty(B, T, N, K, dtype=dt, device=dev),
               gs=torch.randn(B, T, K, Ds, generator=g).to(dev), ga=torch.randn(B, T, N, K, generator=g).to(dt).to(dev),
               gx=torch.empty(B, T, N, D, dtype=dt, device=dev), gp=torch.empty(sz.param_floats, device=dev),
               gn=torch.empty(B, K, Ds, device=dev))
    return shape, sz, buf


def _enqueue(m, shape, b, stream):
    from focus_b200 import _lib
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(stream.cuda_stream)
    params = [p.detach() for p in m._ordered_params()]
    ptrs = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
    _lib.check(_lib.lib.savi_pack_params(ctypes.byref(shape), ptrs, P(b["packed"]), st), "pack")
    _lib.check(_lib.lib.savi_forward(ctypes.byref(shape), P(b["packed"]), P(b["x"]), P(b["noise"]), P(b["slots"]), P(b["attn"]),
                                     P(b["saved"]), P(b["fws"]), ctypes.c_void_p(0), st), "forward")
    _lib.check(_lib.lib.savi_backward(ctypes.byref(shape), P(b["packed"]), P(b["x"]), P(b["noise"]), P(b["saved"]), P(b["gs"]), P(b["ga"]),
                                      P(b["gx"]), P(b["gp"]), P(b["gn"]), P(b["bws"]), ctypes.c_void_p(0), st), "backward")


@pytest.mark.parametrize("cfg", [
    (torch.bfloat16, 3, 24, 128, 128, 128, 1, 4, 8, 3, 1024),       # tcgen05 family (overlapped dependent launches inside the graph)
    (torch.bfloat16, 2, 15, 192, 192, 192, 1, 4, 3, 2, 512),        # mma.sync family
    (torch.float32, 3, 7, 64, 64, 128, 1, 4, 3, 2, 200),            # SIMT fp32 family
])
def test_pack_forward_backward_replay_as_a_cuda_graph(cfg):
    from focus_b200 import SlotAttentionVideo
    dt, I, K, D, Ds, M, blocks, heads, B, T, N = cfg
    torch.manual_seed(0)
    m = SlotAttentionVideo(I, K, D, Ds, M, blocks, heads, 0.0).cuda()
    shape, sz, b = _buffers(m, B, T, N, dt)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _enqueue(m, shape, b, s)                     # eager, also warms every cudaFuncSetAttribute
    s.synchronize()
    eager = {k: b[k].clone() for k in ("slots", "attn", "gx", "gp", "gn")}
    for k in ("slots", "attn", "gx", "gp", "gn"):
        b[k].zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        _enqueue(m, shape, b, torch.cuda.current_stream())
    for _ in range(3):                               # replays reuse the same buffers: nothing may depend on stale state
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(b["slots"], eager["slots"]) and torch.equal(b["attn"], eager["attn"])
    assert torch.equal(b["gx"], eager["gx"]) and torch.equal(b["gn"], eager["gn"])
    gmax = float(eager["gp"].abs().max())
    assert float((b["gp"] - eager["gp"]).abs().max()) <= 1e-5 * gmax      # atomics may reorder the fp32 sums
    assert torch.isfinite(b["gp"]).all()
