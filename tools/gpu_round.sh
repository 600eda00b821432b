#!/bin/bash
# one GPU call: sanity (smoke) -> GPU test suite -> C2 bench -> phase breakdown -> C4 bench.   usage: tools/gpu_round.sh TAG [pytest args]
cd "$(dirname "$0")/.."
TAG=${1:-r02x}; shift
O=gpurun_out; mkdir -p $O
timeout 600 python __graft_entry__.py --smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${TAG}_smoke.log
grep -E "^smoke|Error|error" $O/${TAG}_smoke.log | tail -5
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 "$@" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -vE "^\s*$" $O/${TAG}_pytest.log | tail -60
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench rc=$?"; cut -c1-400 $O/${TAG}_bench_c2.json
python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_c2.json")); print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "kernels", d["kernels_ms"], "roofline", d["roofline"]["frac"])
except Exception as e: print("no bench line", e)
PY
timeout 300 python tools/phase_times.py c2 > $O/${TAG}_phase_cycles_c2.txt 2>&1; head -75 $O/${TAG}_phase_cycles_c2.txt
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.json 2>/dev/null; cut -c1-200 $O/${TAG}_bench_c4.json
