"""fuzz_other.py [rounds] — soak test of g2048_expand (fused small-batch kernel and the two-kernel
path, compact and dense), g2048_env_step trajectories and g2048_rollout against the C oracle on
fresh random batches.  Test infrastructure: run under gpurun; writes gpurun_out/fuzz_other.json."""
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
from tests.soak.fuzz_step import gen_boards  # noqa: E402


def host(t):
    a = t.cpu().numpy()
    return a.view(np.uint64) if a.dtype == np.int64 else a


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    ops.init(0)
    dev = torch.device('cuda:0')
    c_oracle.build()
    rng = np.random.default_rng(777)
    kinds = ['synthetic', 'dense_low', 'big', 'sparse', 'midgame']
    tot = {'expand_boards': 0, 'expand_children': 0, 'env_steps': 0, 'rollout_games': 0, 'rollout_steps': 0,
           'selfplay_steps': 0, 'sampled_children': 0}
    bad = 0
    t0 = time.time()
    for rnd in range(rounds):
        kind = kinds[rnd % len(kinds)]
        # ---- expand: one size below and one above the fused / two-kernel switch (2^17)
        for n in (int(rng.integers(1000, 100000)), (1 << 17) + int(rng.integers(0, 50000))):
            boards = gen_boards(kind, n, rng, dev)
            d = torch.from_numpy(boards.view(np.int64)).to(dev)
            for layout, code in (('compact', 2), ('dense', 1)):
                ex = ops.expand(d, children=layout)
                want = c_oracle.expand(boards, child_layout=code)
                ok = (np.array_equal(host(ex.after), want['after'].reshape(n, 4)) and
                      np.array_equal(ex.rewards.cpu().numpy().reshape(-1), np.asarray(want['rewards']).reshape(-1)) and
                      np.array_equal(ex.mask.cpu().numpy(), want['mask']) and
                      np.array_equal(ex.num_empty.cpu().numpy().reshape(-1), np.asarray(want['num_empty']).reshape(-1)) and
                      np.array_equal(ex.child_count.cpu().numpy().reshape(-1), np.asarray(want['child_count']).reshape(-1)) and
                      np.array_equal(host(ex.children).reshape(-1), np.asarray(want['children']).reshape(-1)))
                if not ok:
                    bad += 1
                    print(f'MISMATCH expand {kind} n={n} {layout}', flush=True)
                tot['expand_boards'] += n
                tot['expand_children'] += int(ex.child_count.sum().item())
        # ---- env_step: 40 steps of a GameState batch, random actions, Philox spawns
        n = int(rng.integers(3000, 60000))
        seed = int(rng.integers(1 << 62))
        b = c_oracle.reset(n, seed=seed)
        sc = np.zeros(n, np.int32); dn = np.zeros(n, np.uint8); tr = np.zeros(n, np.float32)
        db = torch.from_numpy(b.view(np.int64).copy()).to(dev)
        dsc = torch.zeros(n, dtype=torch.int32, device=dev); ddn = torch.zeros(n, dtype=torch.uint8, device=dev)
        dtr = torch.zeros(n, dtype=torch.float32, device=dev)
        for t in range(40):
            acts = rng.integers(0, 4, n).astype(np.uint8)
            (b, sc, dn, tr), rew, fl, obs, mt = c_oracle.env_step(b, sc, dn, tr, acts, None, seed=seed)
            drew = torch.empty(n, dtype=torch.float32, device=dev)
            dfl = torch.empty(n, dtype=torch.uint8, device=dev)
            dobs = torch.empty((n, 16), dtype=torch.float32, device=dev)
            ops.env_step(db, dsc, ddn, dtr, torch.from_numpy(acts).to(dev), None, seed=seed, rewards=drew, flags=dfl, obs=dobs)
            ok = (np.array_equal(host(db), b) and np.array_equal(dsc.cpu().numpy(), sc) and
                  np.array_equal(ddn.cpu().numpy(), dn) and np.array_equal(dtr.cpu().numpy(), tr) and
                  np.array_equal(drew.cpu().numpy(), rew) and np.array_equal(dfl.cpu().numpy(), fl) and
                  np.array_equal(dobs.cpu().numpy(), obs))
            if not ok:
                bad += 1
                print(f'MISMATCH env_step round {rnd} t={t}', flush=True)
                break
            tot['env_steps'] += n
        # ---- rollout: every episode against the oracle
        n = int(rng.integers(20000, 120000))
        seed = int(rng.integers(1 << 62)); off = int(rng.integers(1 << 40))
        out = ops.rollout(n, seed=seed, env_offset=off)
        want = c_oracle.rollout(n, seed=seed, env_offset=off)
        ok = (np.array_equal(host(out.boards), want['boards']) and np.array_equal(out.score.cpu().numpy(), want['score']) and
              np.array_equal(out.steps.cpu().numpy(), want['steps']) and
              np.array_equal(out.stats.cpu().numpy().view(np.uint64), want['stats']))
        if not ok:
            bad += 1
            print(f'MISMATCH rollout round {rnd}', flush=True)
        tot['rollout_games'] += n
        tot['rollout_steps'] += int(out.steps.sum().item())
        # ---- rollout from given boards (mid-game positions, a step cap) against the oracle
        n = int(rng.integers(5000, 90000))
        start = gen_boards('midgame' if rnd % 2 else 'dense_low', n, rng, dev)
        seed = int(rng.integers(1 << 62)); cap = int(rng.integers(20, 400))
        out = ops.rollout(n, seed=seed, boards=torch.from_numpy(start.view(np.int64).copy()).to(dev), max_steps=cap)
        want = c_oracle.rollout(n, seed=seed, from_reset=False, boards=start, max_steps=cap)
        ok = (np.array_equal(host(out.boards), want['boards']) and np.array_equal(out.score.cpu().numpy(), want['score']) and
              np.array_equal(out.steps.cpu().numpy(), want['steps']) and np.array_equal(out.maxexp.cpu().numpy(), want['maxexp']) and
              np.array_equal(out.stats.cpu().numpy().view(np.uint64), want['stats']))
        if not ok:
            bad += 1
            print(f'MISMATCH rollout-from-boards round {rnd}', flush=True)
        tot['rollout_games'] += n
        tot['rollout_steps'] += int(out.steps.sum().item())
        # ---- selfplay_step: sampled actions (= the standalone sampler's), recorded trajectory, state vs oracle
        n = int(rng.integers(3000, 40000))
        seed = int(rng.integers(1 << 62))
        b = c_oracle.reset(n, seed=seed)
        sc = np.zeros(n, np.int32); dn = np.zeros(n, np.uint8); tr = np.zeros(n, np.float32)
        db = torch.from_numpy(b.view(np.int64).copy()).to(dev)
        dsc = torch.zeros(n, dtype=torch.int32, device=dev); ddn = torch.zeros(n, dtype=torch.uint8, device=dev)
        dtr = torch.zeros(n, dtype=torch.float32, device=dev)
        dact = torch.empty(n, dtype=torch.uint8, device=dev)
        drew = torch.empty(n, dtype=torch.float32, device=dev)
        for t in range(30):
            pol = torch.rand((n, 4), device=dev) + 0.01
            legal = ops.legal_actions(db)
            alone = ops.sample_action(pol, legal, 0.7, seed=seed, step_ctr=dsc)
            ops.selfplay_step(db, dsc, ddn, dtr, pol, temperature=0.7, seed=seed, rewards=drew, actions_out=dact)
            acts = dact.cpu().numpy()
            (b, sc, dn, tr), rew, fl, obs, mt = c_oracle.env_step(b, sc, dn, tr, acts, None, seed=seed)
            lg = legal.cpu().numpy()
            ok = (np.array_equal(host(db), b) and np.array_equal(dsc.cpu().numpy(), sc) and
                  np.array_equal(ddn.cpu().numpy(), dn) and np.array_equal(dtr.cpu().numpy(), tr) and
                  np.array_equal(drew.cpu().numpy(), rew) and np.array_equal(alone.cpu().numpy(), acts) and
                  bool(lg[lg.any(axis=1), acts[lg.any(axis=1)]].all()))
            if not ok:
                bad += 1
                print(f'MISMATCH selfplay_step round {rnd} t={t}', flush=True)
                break
            tot['selfplay_steps'] += n
        # ---- sampled chance children against the oracle
        n = int(rng.integers(2000, 50000)); k = int(rng.integers(1, 9))
        boards = gen_boards(kind, n, rng, dev)
        ex = ops.expand(torch.from_numpy(boards.view(np.int64)).to(dev), children='none')
        seed = int(rng.integers(1 << 62))
        kids = ops.sample_children(ex, k, seed=seed, env_offset=rnd, draw_base=3 * rnd)
        want = c_oracle.sample_children(host(ex.after), ex.child_count.cpu().numpy(), k, seed=seed, env_offset=rnd, draw_base=3 * rnd)
        if not np.array_equal(host(kids), want):
            bad += 1
            print(f'MISMATCH sample_children round {rnd}', flush=True)
        tot['sampled_children'] += n * 4 * k
        print(f'round {rnd} {kind}: ok so far, mismatching checks: {bad}', flush=True)
    line = dict(tot, mismatching_checks=bad, rounds=rounds, seconds=round(time.time() - t0, 1),
                oracle='oracle/g2048_oracle.c (orc_expand, orc_env_step, orc_rollout, orc_sample_children)')
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'fuzz_other.json'), 'w') as f:
        json.dump(line, f, indent=1)
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
