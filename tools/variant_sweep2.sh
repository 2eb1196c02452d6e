#!/bin/bash
# like variant_sweep.sh, plus the rollout timing per variant
cp simulate_2048_b200/libg2048.so /tmp/orig.so
for f in variants/lib_*.so; do
  v=$(basename $f .so); v=${v#lib_}
  cp $f simulate_2048_b200/libg2048.so
  python bench.py --steps 20 --warmup 3 --quick --no-cpu --e2e-log2 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e9,1), round(d['roofline']['frac'],4), 'philox', round(d['philox_mode']['value']/1e9,1))"
  python tools/time_rollout.py 2>/dev/null | tr '\n' ';'; echo
done
cp /tmp/orig.so simulate_2048_b200/libg2048.so
