"""Import the UNMODIFIED reference SlotAttentionVideo from /root/reference.

TEST INFRASTRUCTURE ONLY.  In the build container the package is imported from
/root/reference; on the GPU box (no such path) from oracle/_ref/, the byte-identical
copies oracle/make_ref.py makes (git-ignored, shipped with the gpurun snapshot).
Nothing under focus_b200/ imports this file.  Users: tests/golden/make_golden.py
(fixture generation), tests/test_reference_integration.py, bench.py's reference arm
and its gpu_eager_baseline leg.

The reference package cannot be imported normally: slowfast/__init__.py pulls
iopath and slowfast/models/__init__.py pulls fvcore (neither installed).  The
STEVE sub-package itself only needs torch plus `..build.MODEL_REGISTRY`
(/root/reference/slowfast/models/build.py:7-9), so we register hollow parent
packages and a tiny stand-in for fvcore's Registry, then import
slowfast.models.STEVE.steve unmodified.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_TRAVEL_ROOT = os.path.join(_HERE, "_ref")      # byte-identical copies made by oracle/make_ref.py (git-ignored; ships to the GPU box)


def _has_steve(root):
    return os.path.isfile(os.path.join(root, "slowfast", "models", "STEVE", "steve.py"))


REFERENCE_ROOT = os.environ.get("FOCUS_REFERENCE_ROOT", "/root/reference")
if not _has_steve(REFERENCE_ROOT) and _has_steve(_TRAVEL_ROOT):
    REFERENCE_ROOT = _TRAVEL_ROOT


def reference_available() -> bool:
    return _has_steve(REFERENCE_ROOT)


def load_reference_steve():
    """Return the reference module slowfast.models.STEVE.steve (unmodified source)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "slowfast.models.STEVE.steve" in sys.modules:
        return sys.modules["slowfast.models.STEVE.steve"]

    if "fvcore.common.registry" not in sys.modules:
        class Registry(dict):
            def __init__(self, name):
                super().__init__()
                self._name = name

            def register(self, obj=None):
                if obj is None:
                    return lambda o: self.register(o)
                self[obj.__name__] = obj
                return obj

            def get(self, name):
                return self[name]

        fv = types.ModuleType("fvcore")
        fvc = types.ModuleType("fvcore.common")
        fvr = types.ModuleType("fvcore.common.registry")
        fvr.Registry = Registry
        fv.common = fvc
        fvc.registry = fvr
        sys.modules.setdefault("fvcore", fv)
        sys.modules.setdefault("fvcore.common", fvc)
        sys.modules["fvcore.common.registry"] = fvr

    for name, rel in (("slowfast", "slowfast"), ("slowfast.models", "slowfast/models")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            sys.modules[name] = m
    return importlib.import_module("slowfast.models.STEVE.steve")


def reference_slot_attention_video(*args, **kwargs):
    return load_reference_steve().SlotAttentionVideo(*args, **kwargs)


def steve_config(img_size=32, dim=128, slot_size=128, mlp_hidden=128, num_slots=7, num_iters=2, blocks=1, heads=4,
                 vocab=64, cnn_hidden=32, dec_blocks=1, dec_heads=4, dropout=0.0):
    """A cfg object with the fields STEVE.__init__ reads (steve.py:160-275; yaml names of configs/movi_e/base.yaml)."""
    ns = types.SimpleNamespace
    return ns(MODEL=ns(CNN_NAME="base"),
              SLOTS=ns(NUM_ITERS=num_iters, NUM_SLOTS=num_slots, CNN_HID_SIZE=cnn_hidden, SIZE=slot_size, MLP_HID_SIZE=mlp_hidden,
                       IMG_CHANNELS=3, IMG_SIZE=img_size, VOCAB_SIZE=vocab, DIM=dim, NUM_PREDICTOR_BLOCKS=blocks,
                       NUM_PREDICTOR_HEADS=heads, PREDICTOR_DROPOUT=dropout,
                       DECODER=ns(DIM=dim, NUM_BLOCKS=dec_blocks, NUM_HEADS=dec_heads, DROPOUT=dropout)))


def load_reference_metrics():
    """The reference's slowfast/utils/metrics.py (evaluate_ari, compute_mask_ari, compute_ari), loaded by file path: the
    slowfast.utils package itself pulls iopath / fvcore."""
    import importlib.util
    path = os.path.join(REFERENCE_ROOT, "slowfast", "utils", "metrics.py")
    if not os.path.isfile(path):
        raise RuntimeError("reference metrics.py not present under %s" % REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location("_focus_reference_metrics", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
