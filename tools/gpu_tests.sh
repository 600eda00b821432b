#!/bin/bash
# GPU test suite (+ optional extra commands): usage tools/gpu_tests.sh TAG [pytest args]
cd "$(dirname "$0")/.."
TAG=${1:-r02x}; shift
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 "$@" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -vE "^\s*$" $O/${TAG}_pytest.log | tail -${TAIL:-80}
