#!/bin/bash
# tools/profile_expand.sh <tag> — ncu --set full of the expand kernels on 2^20 boards (run under gpurun)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_other.py > $OUT/${TAG}_expand_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_expand_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"expand_" -s 2 -c 2 -f -o $OUT/${TAG}_expand python tools/profile_other.py > $OUT/${TAG}_expand_ncu.log 2>&1
echo "ncu rc=$?"
