"""fuzz_step.py [rounds] [log2 boards per round] — soak test of g2048_step against the C oracle.

Every round draws a fresh batch from one of several board generators (the SURVEY.md synthetic
recipe, dense low tiles that merge and fill up, boards with 16384 / 32768 tiles for the exact
path, near-empty boards, mid-game boards taken from capped device rollouts), with action and spawn
bytes drawn from the WHOLE byte range (high action bits, invalid ranks, absolute-cell spawn bytes),
and compares boards, rewards and flags of the injected-spawn and the Philox kernel with
oracle/g2048_oracle.c (pthreads) bit for bit.  Test infrastructure: run under gpurun, prints one
summary line per round and a JSON total (profiles/r01_fuzz_step.json keeps the last run; lives under tests/ because it uses oracle/).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle  # noqa: E402
from simulate_2048_b200 import ops  # noqa: E402

SHIFTS = (4 * np.arange(16, dtype=np.uint64))[None, :]


def pack(exps):
    return np.bitwise_or.reduce(exps.astype(np.uint64) << SHIFTS, axis=1)


def gen_boards(kind, n, rng, dev):
    if kind == 'synthetic':      # n tiles ~ U{1..16}, exponents ~ U{1..11}
        fill = rng.integers(1, 17, n)[:, None]
        order = np.argsort(rng.random((n, 16)), axis=1)
        return pack(np.where(order < fill, rng.integers(1, 12, (n, 16)), 0))
    if kind == 'dense_low':      # exponents 1..3 on 10..16 cells: merges, full boards, game over
        fill = rng.integers(10, 17, n)[:, None]
        order = np.argsort(rng.random((n, 16)), axis=1)
        return pack(np.where(order < fill, rng.integers(1, 4, (n, 16)), 0))
    if kind == 'big':            # tiles up to 32768: exact path, 15+15 overflow
        fill = rng.integers(1, 17, n)[:, None]
        order = np.argsort(rng.random((n, 16)), axis=1)
        return pack(np.where(order < fill, rng.integers(1, 16, (n, 16)), 0))
    if kind == 'sparse':         # 0..3 tiles
        fill = rng.integers(0, 4, n)[:, None]
        order = np.argsort(rng.random((n, 16)), axis=1)
        return pack(np.where(order < fill, rng.integers(1, 14, (n, 16)), 0))
    if kind == 'midgame':        # boards reached by random play after a capped number of steps
        cap = int(rng.integers(5, 200))
        r = ops.rollout(n, seed=int(rng.integers(1 << 62)), max_steps=cap, device=dev)
        return r.boards.cpu().numpy().view(np.uint64).copy()
    raise ValueError(kind)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    n = (1 << lg) + 37           # full tiles of the smem-table kernel + a ragged tail
    threads = len(os.sched_getaffinity(0))
    ops.init(0)
    dev = torch.device('cuda:0')
    c_oracle.build()
    rng = np.random.default_rng(20481)
    kinds = ['synthetic', 'dense_low', 'big', 'sparse', 'midgame']
    total = bad = 0
    per_kind = {}
    t0 = time.time()
    for rnd in range(rounds):
        kind = kinds[rnd % len(kinds)]
        boards = gen_boards(kind, n, rng, dev)
        actions = rng.integers(0, 256, n).astype(np.uint8)
        spawns = rng.integers(0, 256, n).astype(np.uint8)
        if rnd % 2 == 0:         # mostly valid rank bytes, so that real spawns dominate
            spawns &= np.uint8(0x1F)
        seed = int(rng.integers(1 << 62))
        off = int(rng.integers(1 << 40))
        d_b = torch.from_numpy(boards.view(np.int64)).to(dev)
        d_a = torch.from_numpy(actions).to(dev)
        d_s = torch.from_numpy(spawns).to(dev)
        mism = 0
        for mode, sp, d_sp in (('injected', spawns, d_s), ('philox', None, None)):
            out = ops.step(d_b, d_a, d_sp, seed=seed, env_offset=off, step_base=rnd)
            wb, wr, wf = c_oracle.step(boards, actions, sp, seed=seed, env_offset=off, step_base=rnd,
                                       threads=threads)
            gb = out.boards.cpu().numpy().view(np.uint64)
            m = (gb != wb) | (out.rewards.cpu().numpy() != wr) | (out.flags.cpu().numpy() != wf)
            k = int(m.sum())
            if k:
                i = int(np.flatnonzero(m)[0])
                print(f'MISMATCH {kind}/{mode} env {i}: board {boards[i]:016x} a={actions[i]} s={spawns[i]:02x} '
                      f'got {gb[i]:016x} want {wb[i]:016x}', flush=True)
            mism += k
            total += n
        bad += mism
        pk = per_kind.setdefault(kind, [0, 0])
        pk[0] += 2 * n
        pk[1] += mism
        print(f'round {rnd} {kind}: {2 * n} transitions, {mism} mismatches', flush=True)
    line = {'transitions': total, 'mismatches': bad, 'rounds': rounds, 'boards_per_round': n,
            'per_generator': {k: {'transitions': v[0], 'mismatches': v[1]} for k, v in per_kind.items()},
            'seconds': round(time.time() - t0, 1), 'oracle': 'oracle/g2048_oracle.c orc_step_mt',
            'modes': ['injected (whole spawn-byte range)', 'philox']}
    print(json.dumps(line))
    out_dir = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, 'fuzz_step.json'), 'w') as f:
        json.dump(line, f, indent=1)
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
