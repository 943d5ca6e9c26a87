#!/bin/bash
# Runs on the GPU box (via gpurun): plain run first, then the ncu launch list and one full capture of the top kernel.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex] [extra bench args]
set -u
TAG=${1:-r01}
KREGEX=${2:-prove_kernel}
shift 2 || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --ring 2 --e2e-steps 0 $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 2 -f -o $OUT/${TAG}_${KREGEX} $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
