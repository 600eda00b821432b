#!/bin/bash
# round-2 GPU call 1: sanity (smoke) -> full GPU test suite -> C2 bench (fp16 weight images vs the bf16 hi+lo variant) -> phase breakdown
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02a_smi.txt 2>&1
timeout 600 python __graft_entry__.py --smoke > $O/r02a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r02a_smoke.log
tail -3 $O/r02a_smoke.log
if ! grep -q "smoke rc=0" $O/r02a_smoke.log; then echo "SMOKE FAILED: trying the split-weights variant"; FOCUS_SAVI_LIB=focus_b200/variants/split.so timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3; fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 $O/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench_c2.json 2> $O/r02a_bench_c2.err; echo "bench rc=$?"; cat $O/r02a_bench_c2.json | cut -c1-600
FOCUS_SAVI_LIB=focus_b200/variants/split.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02a_bench_c2_split.json 2> $O/r02a_bench_c2_split.err; echo "bench split rc=$?"; cut -c1-300 $O/r02a_bench_c2_split.json
timeout 300 python tools/phase_times.py c2 > $O/r02a_phase_cycles_c2.txt 2>&1; cat $O/r02a_phase_cycles_c2.txt | head -70
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_c4.json 2>/dev/null; cut -c1-300 $O/r02a_bench_c4.json
