#!/bin/bash
# tools/profile_other.sh <tag> — ncu --set full of the rollout and expand kernels (run under gpurun)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_other.py > $OUT/${TAG}_other_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_other_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|expand_" -s 3 -c 4 -f -o $OUT/${TAG}_other python tools/profile_other.py > $OUT/${TAG}_other_ncu.log 2>&1
echo "ncu rc=$?"
ls -la $OUT | grep other
