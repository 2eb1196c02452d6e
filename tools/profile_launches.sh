#!/bin/bash
# tools/profile_launches.sh <tag> — ncu launch list (gpu__time_duration per launch) of a short bench run.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 6 --warmup 3 --quick --no-cpu --boards-log2 25 --e2e-log2 20"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
