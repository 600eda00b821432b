"""setup.py — builds the C-ABI CUDA library in-tree (focus_b200/libfocus_savi.so) with nvcc for sm_100a.

    python setup.py build_ext --inplace        # = python focus_b200/build.py
    pip install -e . --no-build-isolation      # development install; the library stays in the source tree

The library has no Python / torch dependency (plain `extern "C"`, loaded with ctypes), so there is no
setuptools Extension to compile: build_ext just drives focus_b200/build.py (one nvcc call per translation
unit, `-gencode arch=compute_100a,code=sm_100a -lineinfo`).
"""
import os
import sys

from setuptools import Command, find_packages, setup

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildCudaLibrary(Command):
    description = "compile focus_b200/csrc/*.cu into focus_b200/libfocus_savi.so (nvcc, sm_100a)"
    user_options = [("inplace", "i", "accepted for compatibility; the library is always built in-tree"),
                    ("force", "f", "rebuild even if the library is newer than its sources")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = True
        self.force = False

    def finalize_options(self):
        pass

    def run(self):
        # load the build script by path: importing the package would load (and symbol-check) a stale library first
        import importlib.util
        spec = importlib.util.spec_from_file_location("_focus_b200_build", os.path.join(ROOT, "focus_b200", "build.py"))
        b = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(b)
        print("built", b.build(force=bool(self.force), verbose=False))


setup(
    name="focus_b200",
    version="0.1.0",
    description="B200-native (sm_100a) video slot-attention encoder for srv902/FOCUS (STEVE): drop-in SlotAttentionVideo",
    packages=find_packages(include=["focus_b200", "focus_b200.*"]),
    package_data={"focus_b200": ["libfocus_savi.so", "csrc/*"]},
    python_requires=">=3.9",
    cmdclass={"build_ext": BuildCudaLibrary},
)
