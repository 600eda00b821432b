"""CPU: the reference arm of bench.py prints exactly one JSON line with the contract's keys (the GPU arm shares the code that
builds `config`, `metric`, `unit`); the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import torch

from tests._util import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    assert d["value"] > 0


def test_reference_arm_never_loads_the_product():
    """The reference arm times the UNMODIFIED reference module: neither the focus_b200 package nor its CUDA library may be
    imported / mapped into that process (`reference.native_so_loaded` listed libfocus_savi.so)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'c1', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read();"
            "assert not any(m == 'focus_b200' or m.startswith('focus_b200.') for m in sys.modules), 'focus_b200 imported';"
            "assert 'libfocus_savi' not in maps, 'CUDA library mapped'; print('CLEAN', file=sys.stderr)") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "CLEAN" in r.stderr, r.stderr[-2000:]


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
