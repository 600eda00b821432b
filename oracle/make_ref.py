"""Make the UNMODIFIED reference STEVE package travel to the GPU box.

TEST INFRASTRUCTURE ONLY.  `/root/reference` exists only in the build container; `oracle/_ref/` is git-ignored
(it never enters the history) but is NOT gpurun-ignored, so whatever this script puts there ships with the
snapshot.  It copies, byte for byte, the four source files of the reference's STEVE package

    slowfast/models/STEVE/{steve.py, transformer.py, utils.py, dvae.py}   (+ the empty __init__.py)
    slowfast/utils/metrics.py                                             (FG-ARI, for the §8f N4 parity tests)

into `oracle/_ref/slowfast/models/STEVE/` and writes `oracle/_ref/slowfast/models/build.py`, a stand-in of our own for
the only symbol steve.py takes from the rest of PySlowFast (`MODEL_REGISTRY`, reference slowfast/models/build.py:7-9;
the real file pulls fvcore + the whole model zoo).  `oracle/_load_reference.py` then imports the package from
`/root/reference` when present, else from `oracle/_ref` — so on the B200 box
  * `bench.py --impl reference` times the real module on the host cores (kind "reference"),
  * `bench.py`'s `gpu_eager_baseline` leg runs it through PyTorch eager on the GPU,
  * tests/test_reference_integration.py swaps our module into the reference `STEVE` model.
A sha256 manifest is written next to the copies; `python oracle/make_ref.py --check` verifies it against the
reference tree (run by the CPU tests in the build container).

    python oracle/make_ref.py            # (re)create oracle/_ref from /root/reference
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FOCUS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["slowfast/models/STEVE/__init__.py", "slowfast/models/STEVE/steve.py", "slowfast/models/STEVE/transformer.py",
         "slowfast/models/STEVE/utils.py", "slowfast/models/STEVE/dvae.py",
         "slowfast/utils/metrics.py"]                 # FG-ARI; loaded by path, never through slowfast.utils
BUILD_STUB = '''"""Stand-in (ours, not reference code) for slowfast/models/build.py: only MODEL_REGISTRY is needed by STEVE/steve.py."""
from fvcore.common.registry import Registry

MODEL_REGISTRY = Registry("MODEL")
'''


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make():
    if not os.path.isfile(os.path.join(REF_ROOT, FILES[1])):
        raise SystemExit("reference tree not found at %s" % REF_ROOT)
    manifest = {}
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_ROOT, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(OUT, "slowfast", "models", "build.py"), "w") as f:
        f.write(BUILD_STUB)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest}, f, indent=1)
    return manifest


def check():
    """True when oracle/_ref holds byte-identical copies of the reference files."""
    mp = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isfile(mp):
        return False
    with open(mp) as f:
        manifest = json.load(f)["sha256"]
    for rel in FILES:
        p = os.path.join(OUT, rel)
        if not os.path.isfile(p) or _sha(p) != manifest.get(rel):
            return False
        src = os.path.join(REF_ROOT, rel)
        if os.path.isfile(src) and _sha(src) != manifest[rel]:
            return False
    return True


if __name__ == "__main__":
    if "--check" in sys.argv:
        ok = check()
        print("oracle/_ref", "matches the reference" if ok else "is missing or differs")
        sys.exit(0 if ok else 1)
    print(json.dumps(make(), indent=1))
