"""Builds focus_b200/libfocus_savi.so in-tree with nvcc for sm_100a (no torch headers)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfocus_savi.so")
# (source, object name, extra defines): the clip kernels are compiled once per token dtype
UNITS = [("savi_api.cu", "savi_api.o", []),
         ("savi_wgrad.cu", "savi_wgrad.o", []),
         ("savi_fwd_umma.cu", "savi_fwd_umma.o", []),
         ("savi_bwd_umma.cu", "savi_bwd_umma.o", []),
         ("savi_dx_umma.cu", "savi_dx_umma.o", []),
         ("savi_wgrad_umma.cu", "savi_wgrad_umma.o", []),
         ("savi_allreduce.cu", "savi_allreduce.o", []),
         ("steve_neighbors.cu", "steve_neighbors.o", []),
         ("steve_token_mlp.cu", "steve_token_mlp.o", []),
         ("savi_fwd.cu", "savi_fwd_f32.o", ["-DSAVI_TOK=float", "-DSAVI_SUFFIX=f32"]),
         ("savi_fwd.cu", "savi_fwd_bf16.o", ["-DSAVI_TOK=__nv_bfloat16", "-DSAVI_SUFFIX=bf16"]),
         ("savi_bwd.cu", "savi_bwd_f32.o", ["-DSAVI_TOK=float", "-DSAVI_SUFFIX=f32"]),
         ("savi_bwd.cu", "savi_bwd_bf16.o", ["-DSAVI_TOK=__nv_bfloat16", "-DSAVI_SUFFIX=bf16", "-DSAVI_IS_BF16"])]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _newest_source_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(force=False, verbose=False, out=None, extra_flags=()):
    """Compile the CUDA library if it is missing or older than its sources. Returns the .so path.
    `out` / `extra_flags`: development A/B builds (e.g. focus_b200/variants/x.so with -DSAVI_WIMG_SPLIT=1), loaded through
    the FOCUS_SAVI_LIB override of focus_b200/_lib.py."""
    OUT = out or globals()["OUT"]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build") if out is None else os.path.join(HERE, "build", os.path.basename(OUT) + ".d")
    os.makedirs(objdir, exist_ok=True)
    extra = (["-Xptxas", "-v"] if verbose else []) + os.environ.get("SAVI_NVCC_EXTRA", "").split() + list(extra_flags)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)

    def cc(unit):
        src, oname, defs = unit
        obj = os.path.join(objdir, oname)
        cmd = [nvcc] + NVCC_FLAGS + extra + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (oname, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(cc, UNITS))
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a not in ("--force", "-v")]
    if args:        # python -m focus_b200.build [--force] [-v] OUT.so [nvcc flags ...]
        print(build(force=True, verbose="-v" in sys.argv, out=os.path.abspath(args[0]), extra_flags=args[1:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
