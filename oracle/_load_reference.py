"""Import the UNMODIFIED reference SlotAttentionVideo from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Works only in the build
container where /root/reference is mounted; the GPU box has no such path, so
nothing on the product path, in `-m gpu` tests, smoke() or bench.py imports
this file.  It is used by tests/golden/make_golden.py (fixture generation) and
by CPU tests that are skipped when the reference tree is absent.

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

REFERENCE_ROOT = os.environ.get("FOCUS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "slowfast", "models", "STEVE", "steve.py"))


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
