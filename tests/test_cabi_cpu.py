"""CPU (no GPU): the C-ABI library loads and exports every symbol include/focus_savi.h declares,
shape validation / sizing / parameter layout behave, and the module mirrors the reference's
constructor, state_dict and initialisation.  No compute entry point is called here."""
import ctypes
import os
import re

import pytest
import torch

from oracle import savi_numpy as O
from oracle._load_reference import reference_available, reference_slot_attention_video
from tests._util import ROOT


def _lib():
    from focus_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "focus_savi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(savi_[a-z_]+)\s*\(", hdr))
    assert {"savi_forward", "savi_backward", "savi_query", "savi_pack_params", "savi_param_layout"} <= declared
    L = _lib()
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert declared == set(L.EXPORTS)
    assert L.lib.savi_version() == 2


def test_query_sizes_and_validation():
    L = _lib()
    ok = L.SaviShape(B=64, T=6, N=1024, D=128, Ds=128, M=128, K=24, I=3, blocks=1, heads=4, dtype=1, cluster=0, eps=1e-8, ln_eps=1e-5)
    sz = L.query(ok)
    assert sz.n_params == 33 and sz.param_floats == 380288 and sz.cluster == 2
    tokens = 64 * 6 * 1024 * 128 * 2
    # tcgen05 path: ONE copy of the LayerNorm'd tokens (the swizzled image, 100 MB) + ~200 MB of fp32 slot-side records
    assert 2 * tokens < sz.saved_bytes < 3 * tokens, sz.saved_bytes
    # heads only matter when there are predictor blocks (the reference builds no attention otherwise: transformer.py:94-101)
    L.query(L.SaviShape(B=1, T=2, N=8, D=16, Ds=16, M=16, K=4, I=1, blocks=0, heads=3, dtype=0, cluster=0, eps=1e-8, ln_eps=1e-5))
    with pytest.raises(RuntimeError, match="65535"):  # grid limit of the token-parallel kernels
        L.query(L.SaviShape(B=11000, T=6, N=8, D=16, Ds=16, M=16, K=4, I=1, blocks=0, heads=1, dtype=0, cluster=0, eps=1e-8, ln_eps=1e-5))
    with pytest.raises(RuntimeError, match="unknown option"):
        L.set_option("no_such_option", 1)
    L.set_option("no_overlap", 0)
    for field, bad in (("K", 65), ("K", 0), ("D", 12), ("Ds", 130), ("heads", 3), ("blocks", 5), ("dtype", 7), ("cluster", 3), ("I", 0)):
        s = L.SaviShape(B=1, T=1, N=8, D=16, Ds=16, M=16, K=4, I=1, blocks=1, heads=2, dtype=0, cluster=0, eps=1e-8, ln_eps=1e-5)
        setattr(s, field, bad)
        with pytest.raises(RuntimeError):
            L.query(s)
        assert L.lib.savi_last_error()
    # cluster choice: as many CTAs per clip as fit one wave of 148 SMs with >= 64 tokens each
    for B, N, want in ((2, 1024, 8), (64, 1024, 2), (200, 1024, 1), (2, 100, 1), (16, 4096, 8)):
        s = L.SaviShape(B=B, T=1, N=N, D=16, Ds=16, M=16, K=4, I=1, blocks=0, heads=1, dtype=0, cluster=0, eps=1e-8, ln_eps=1e-5)
        assert L.query(s).cluster == want, (B, N)


def test_query_reports_the_kernel_family():
    L = _lib()
    mk = lambda **kw: L.SaviShape(**{**dict(B=64, T=6, N=1024, D=128, Ds=128, M=128, K=24, I=3, blocks=1, heads=4, dtype=1,
                                          cluster=0, eps=1e-8, ln_eps=1e-5), **kw})
    assert L.PATH_NAMES[L.query(mk()).path] == "tcgen05-bf16"                       # C2
    assert L.PATH_NAMES[L.query(mk(K=11, I=2, T=24)).path] == "tcgen05-bf16"        # C4
    assert L.PATH_NAMES[L.query(mk(N=4096, D=192, Ds=192, M=192)).path] == "mma.sync-bf16"   # C3
    assert L.PATH_NAMES[L.query(mk(K=32)).path] == "mma.sync-bf16"
    assert L.PATH_NAMES[L.query(mk(K=64)).path] == "mma.sync-bf16"                  # d_inputs operands staged in iteration groups
    assert L.PATH_NAMES[L.query(mk(B=2, dtype=0)).path] == "simt-fp32"              # C1


@pytest.mark.parametrize("cfg", [(3, 24, 128, 128, 128, 1, 4), (2, 15, 192, 192, 192, 1, 4), (3, 7, 64, 32, 1024, 4, 8), (1, 3, 16, 16, 16, 0, 1)])
def test_module_mirrors_reference_contract(cfg):
    from focus_b200 import SlotAttentionVideo
    I, K, D, Ds, M, blocks, heads = cfg
    m = SlotAttentionVideo(I, K, D, Ds, M, blocks, heads, 0.0)
    want = O.param_shapes(K, D, Ds, M, blocks)
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys())
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in want.items())
    assert [n for n, _ in m.named_parameters()] == list(want.keys())
    # flat layout of the C ABI == state_dict order
    L = _lib()
    shape = m.make_shape(2, 2, 64, torch.float32)
    off, num = L.param_layout(shape, len(want))
    assert num == [p.numel() for p in m._ordered_params()]
    assert off == [sum(num[:i]) for i in range(len(num))]
    assert [tuple(p.shape) for p in m._ordered_params()] == [tuple(v) for v in want.values()]
    for attr in ("num_iterations", "num_slots", "input_size", "slot_size", "mlp_hidden_size", "epsilon"):
        assert hasattr(m, attr)


@pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")
def test_initialisation_is_rng_identical_to_reference():
    from focus_b200 import SlotAttentionVideo
    for cfg in [(3, 24, 128, 128, 128, 1, 4, 0.0), (2, 7, 64, 32, 96, 3, 2, 0.0), (1, 3, 16, 16, 16, 0, 1, 0.0)]:
        torch.manual_seed(5)
        ours = SlotAttentionVideo(*cfg).state_dict()
        torch.manual_seed(5)
        ref = reference_slot_attention_video(*cfg).state_dict()
        assert list(ours) == list(ref)
        assert all(torch.equal(ours[k], ref[k]) for k in ref)
        # and a reference checkpoint loads strictly, both directions (utils/checkpoint.py:366-382 matches name+shape)
        SlotAttentionVideo(*cfg).load_state_dict(ref, strict=True)
        reference_slot_attention_video(*cfg).load_state_dict(ours, strict=True)


def test_no_cpu_fallback():
    from focus_b200 import SlotAttentionVideo
    m = SlotAttentionVideo(2, 4, 16, 16, 16, 1, 2, 0.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 2, 8, 16))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "focus_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src, fn
