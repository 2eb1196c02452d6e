"""derive_orient.py - brute-force check/derivation of the per-action orientation words of g2048_fast.cuh:
F_a = P_a . S_a with S_a one delta swap inside each 32-bit half and P_a one byte permutation."""
import itertools
def delta(x, k, m):  # x: list of 8 nibble labels of a half (index = nibble position)
    # delta swap on 32-bit: t=(x^(x>>k))&m; x^=t^(t<<k)  -> swaps nibble p with p+k/4 for p in mask
    x = list(x); d = k // 4
    for p in range(8):
        if (m >> (4*p)) & 0xF:
            x[p], x[p+d] = x[p+d], x[p]
    return x
board = list(range(16))  # label = cell id 4r+c at nibble position 4r+c
def desired(a):
    O = [[None]*4 for _ in range(4)]
    for i in range(4):
        for j in range(4):
            if a == 0: r, c = i, j
            elif a == 2: r, c = i, 3-j
            elif a == 1: r, c = j, i
            else: r, c = 3-j, 3-i
            O[i][j] = 4*r+c
    return [O[i][j] for i in range(4) for j in range(4)]
stages = {0: (0, 0), 1: (12, 0x0000F0F0), 2: (4, 0x0F0F0F0F), 3: (20, 0x00000F0F)}
for a in range(4):
    k, m = stages[a]
    lo, hi = board[:8], board[8:]
    if m: lo, hi = delta(lo, k, m), delta(hi, k, m)
    src = lo + hi
    bytes_src = [(src[2*b], src[2*b+1]) for b in range(8)]
    want = desired(a)
    sel = []
    ok = True
    for b in range(8):
        w = (want[2*b], want[2*b+1])
        if w in bytes_src: sel.append(bytes_src.index(w))
        else: ok = False; sel.append(None)
    print(a, ok, sel)
    if ok:
        f_lo = sum(s << (4*i) for i, s in enumerate(sel[:4])); f_hi = sum(s << (4*i) for i, s in enumerate(sel[4:]))
        inv = [sel.index(b) for b in range(8)]
        b_lo = sum(s << (4*i) for i, s in enumerate(inv[:4])); b_hi = sum(s << (4*i) for i, s in enumerate(inv[4:]))
        print('   f_lo %04x f_hi %04x b_lo %04x b_hi %04x  mul_hi %08x mul_lo %08x mask %08x' % (f_lo, f_hi, b_lo, b_hi, (1 << (32-k)) & 0xFFFFFFFF if k else 0, (1 << k) if k else 0, m))
