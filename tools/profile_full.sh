#!/bin/bash
# tools/profile_full.sh <tag> — ONE full ncu capture of the step kernel (run under gpurun, after the
# plain command exited 0).  The launch list is a separate call (tools/profile_launches.sh).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 6 --warmup 3 --quick --no-cpu --boards-log2 25 --e2e-log2 20"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
cat $OUT/${TAG}_plain.json | head -c 600
ncu --set full --clock-control none --import-source on -k regex:step_lut_kernel -s 4 -c 2 -f -o $OUT/${TAG}_step $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT | tail -4
