#!/bin/bash
# ncu --set full of the non-recurrence kernels of the step (token LayerNorm, d_inputs, weight gradients).  usage: tools/gpu_ncu_aux.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-r02x}; O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > $O/${TAG}_plain.log 2> $O/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:"ln_tokens_fwd_img128|dx_umma|wgrad_umma" -s 9 -c 3 -o $O/${TAG}_prof_aux $CMD > $O/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -2 $O/${TAG}_ncu.log
