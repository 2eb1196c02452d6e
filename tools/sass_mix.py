#!/usr/bin/env python
"""sass_mix.py <file.sass> [start end] — opcode mix of a SASS listing by issue pipe (B200 rates
from profiles/r01_microbench_pipes.txt)."""
import collections
import re
import sys

PIPE = {
    'LOP3': 'alu', 'SHF': 'alu', 'PRMT': 'alu', 'ISETP': 'alu', 'SEL': 'alu', 'LEA': 'alu', 'IADD3': 'alu',
    'VIADD': 'alu', 'PLOP3': 'alu', 'VIMNMX': 'alu', 'IMNMX': 'alu', 'MOV': 'alu', 'IABS': 'alu', 'FLO': 'xu',
    'BREV': 'xu', 'POPC': 'xu', 'IMAD': 'fma', 'LDS': 'lsu', 'LDG': 'lsu', 'STG': 'lsu', 'LDL': 'lsu',
    'STL': 'lsu', 'LDC': 'lsu', 'BRA': 'cbu', 'BSSY': 'cbu', 'BSYNC': 'cbu', 'CALL': 'cbu', 'RET': 'cbu',
    'EXIT': 'cbu', 'BAR': 'cbu', 'WARPSYNC': 'cbu', 'NOP': 'none', 'S2UR': 'uni', 'UMOV': 'uni', 'ULEA': 'uni',
    'UIADD3': 'uni', 'USHF': 'uni', 'ULOP3': 'uni', 'UISETP': 'uni', 'UIMAD': 'uni', 'LDCU': 'uni', 'S2R': 'xu',
    'CS2R': 'alu', 'R2UR': 'uni', 'ELECT': 'uni', 'UBLKCP': 'uni', 'SYNCS': 'uni', 'UPLOP3': 'uni', 'USEL': 'uni',
    'R2P': 'alu', 'P2R': 'alu', 'SHFL': 'lsu', 'VOTE': 'alu', 'ATOM': 'lsu', 'ATOMS': 'lsu', 'ATOMG': 'lsu', 'RED': 'lsu',
}
lines = open(sys.argv[1]).read().split('\n')
if len(sys.argv) > 3:
    lines = lines[int(sys.argv[2]) - 1:int(sys.argv[3])]
ops = collections.Counter()
pipes = collections.Counter()
for l in lines:
    m = re.search(r'\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)([.A-Z0-9_]*)', l)
    if not m:
        continue
    op = m.group(1)
    full = op + m.group(2)
    key = op
    if op == 'IMAD':
        key = 'IMAD.HI' if '.HI' in full else ('IMAD.WIDE' if 'WIDE' in full else 'IMAD')
    ops[key] += 1
    pipes[PIPE.get(op, '?' + op)] += 1
print('total', sum(ops.values()))
print('pipes', dict(pipes))
print(ops.most_common())
