#!/usr/bin/env python
"""launch_summary.py <launches.csv> — per-kernel totals of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv`, tools/profile_launches.sh): number of
launches, total and mean duration, share of the profiled time.  Per-launch times under ncu are
cold-cache and serialised: read the shares, not the absolutes."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 5]
hdr = next(r for r in rows if 'Kernel Name' in r)
ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.defaultdict(list)
for r in rows:
    if r is hdr or r[ki] == 'Kernel Name':
        continue
    try:
        v = float(r[mi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)   # -> microseconds
    name = re.sub(r'\(.*', '', r[ki]).strip()
    if not re.search(r'g2048|step_|rollout|expand|scan_|mask_|observe|onehot|pack|reset|spawn|after_state|expect|nstep|'
                     r'selfplay|sample_|legal_|advance_|build_table|row_table', name):
        name = 'torch elementwise / copy / NCCL kernels (workload generation, statistics)'
    tot[name].append(v)
total = sum(sum(v) for v in tot.values())
for name, v in sorted(tot.items(), key=lambda kv: -sum(kv[1])):
    big = [x for x in v if x > 100.0]
    extra = f'  [launches > 100 us: n={len(big)}, mean {sum(big) / len(big):.1f} us]' if big and len(big) != len(v) else ''
    print(f'{name[:60]:60s} n={len(v):4d} total={sum(v):12.1f} us ({100 * sum(v) / total:5.1f} %) mean={sum(v) / len(v):9.2f} us{extra}')
