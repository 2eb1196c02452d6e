#!/usr/bin/env python
"""sass_report.py [lib.so] — per-kernel SASS evidence for DESIGN.md §2: instruction count, opcode mix by
issue pipe, and the memory / TMA mnemonics (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDG/STG widths).
Writes to stdout; profiles/r02_step_sass_mix.txt is this tool's output for the shipped library."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'simulate_2048_b200', 'libg2048.so')
ALU = {'LOP3', 'SHF', 'PRMT', 'ISETP', 'SEL', 'LEA', 'IADD3', 'VIADD', 'PLOP3', 'VIMNMX', 'MOV', 'CS2R', 'P2R', 'R2P', 'VOTE'}
WANT = ['step_lut_kernelILb0ELi640ELi1', 'step_lut_kernelILb1ELi640ELi1', 'rollout_kernelILb1ELb1ELb0ELi640',
        'expand_quad_kernelILi2ELb1', 'expand_children_kernelILi2', 'expand_moves_smem_kernel', 'selfplay_step_kernelILb1',
        'env_step_kernelILb1', 'rollout_stats_kernel']
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
print('# cuobjdump -sass', os.path.relpath(lib, ROOT), '(sm_100a)')
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0].strip()
    if not any(w in name for w in WANT):
        continue
    ops, mem = collections.Counter(), collections.Counter()
    for l in f.split('\n'):
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)([.A-Z0-9_]*)', l)
        if not m:
            continue
        op, full = m.group(1), m.group(1) + m.group(2)
        key = op
        if op == 'IMAD':
            key = 'IMAD.HI' if '.HI' in full else ('IMAD.WIDE' if 'WIDE' in full else 'IMAD')
        ops[key] += 1
        if op in ('LDG', 'STG', 'UBLKCP', 'SYNCS', 'LDS', 'ATOMG', 'ATOMS', 'RED', 'LDGSTS'):
            mem[full] += 1
    total = sum(ops.values())
    alu = sum(v for k, v in ops.items() if k in ALU)
    fma = sum(v for k, v in ops.items() if k.startswith('IMAD'))
    print(f'\n## {name}\n   {total} instructions; ALU pipe {alu}, IMAD (FMA-heavy pipe) {fma} '
          f'(of which IMAD.HI {ops["IMAD.HI"]}, IMAD.WIDE {ops["IMAD.WIDE"]}), POPC {ops["POPC"]}')
    print('   memory / TMA:', dict(sorted(mem.items())))
    print('   top opcodes :', ops.most_common(14))
