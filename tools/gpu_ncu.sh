#!/bin/bash
# ncu evidence for the bench command (one ncu-using gpurun call): launch list + full capture of the two clip kernels.
# usage: tools/gpu_ncu.sh TAG
cd "$(dirname "$0")/.."
TAG=${1:-r02x}; O=gpurun_out; mkdir -p $O
CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > $O/${TAG}_plain.log 2> $O/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_ncu_launches_c2.csv $CMD > $O/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"; tail -2 $O/${TAG}_ncu1.log
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph"
$CMD2 > $O/${TAG}_plain2.log 2> $O/${TAG}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:savi_.*_umma_kernel -s 6 -c 2 -o $O/${TAG}_prof_clip $CMD2 > $O/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"; tail -2 $O/${TAG}_ncu2.log; ls -la $O/${TAG}_prof_clip.ncu-rep
