#!/bin/bash
# tools/final_evidence.sh — the one-GPU evidence run of a round (under gpurun): GPU tests, smoke, both bench
# arms, the ncu launch list and one full capture of the step kernel, both soak scripts.  Outputs in gpurun_out/r3_*.
set -u
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/r3_tF.log 2>&1; echo "pytest rc=$?" >> $O/r3_tF.log; tail -3 $O/r3_tF.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r3_smokeF.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r3_smokeF.log
python bench.py --impl reference > $O/r3_bF_ref.json 2> $O/r3_bF_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r3_bF_1gpu.json 2> $O/r3_bF_1gpu.err; echo "bench rc=$?"
bash tools/profile_launches.sh r3F; 
bash tools/profile_full.sh r3F | tail -3
python tests/soak/fuzz_other.py 30 > $O/r3_fuzz_other.log 2>&1; echo "fuzz_other rc=$?"; tail -1 $O/r3_fuzz_other.log; cp $O/fuzz_other.json $O/r3_fuzz_other.json
python tests/soak/fuzz_step.py 100 22 > $O/r3_fuzz_step.log 2>&1; echo "fuzz_step rc=$?"; tail -1 $O/r3_fuzz_step.log | head -c 300; cp $O/fuzz_step.json $O/r3_fuzz_step.json 2>/dev/null
