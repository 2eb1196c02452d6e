"""g2048_selfplay_step: action sampling (reinforce/game/core.py:381-424), trajectory recording
(reinforce/training/self_play.py:346-368, replay_buffer.py:35-68), normalised reward
(twentyfortyeight/utils/normalize.py:6) and the DLPack hand-off."""

import numpy as np
import pytest

from tests.helpers import synthetic_packed

pytestmark = pytest.mark.gpu

TOL = 1e-6   # f32 logf / expf on the device against an f32 NumPy evaluation of the same formula


def test_action_distribution_and_draw(oracle):
    """probs_out against the restatement of core.py:401-421 (tolerance TOL, stated in include/g2048.h),
    and the drawn action against the inverse CDF of those probabilities at the env's Philox word
    (seed, env, step_count, tag 4) — exact."""
    import torch
    from oracle import np_engine as npe
    from simulate_2048_b200 import ops
    n = 20000
    rng = np.random.default_rng(4)
    boards = synthetic_packed(n, seed=9, max_exp=7)
    boards[:50] = 0x0121212112121212 ^ 0  # arbitrary pattern
    pol = rng.random((n, 4)).astype(np.float32)
    pol[::7] = 0.0                                    # no mass at all -> uniform over the legal actions
    pol[1::7, 1:] = 0.0                               # all the mass on one (maybe illegal) action
    pol /= np.maximum(pol.sum(axis=1, keepdims=True), 1e-30)
    mask = oracle.mask(boards) & 0xF
    legal = ((mask[:, None] >> np.arange(4)) & 1).astype(bool)
    for temperature in (1.0, 0.5, 0.25, 0.0):
        d_b = torch.from_numpy(boards.view(np.int64)).cuda()
        sc = torch.from_numpy(rng.integers(0, 300, n).astype(np.int32)).cuda()
        sc0 = sc.cpu().numpy().copy()
        done = torch.zeros(n, dtype=torch.uint8, device='cuda')
        tot = torch.zeros(n, dtype=torch.float32, device='cuda')
        probs = torch.empty((n, 4), dtype=torch.float32, device='cuda')
        acts = torch.empty(n, dtype=torch.uint8, device='cuda')
        ops.selfplay_step(d_b, sc, done, tot, torch.from_numpy(pol).cuda(), temperature=temperature, seed=77,
                          env_offset=1000, probs_out=probs, actions_out=acts)
        got = probs.cpu().numpy()
        want = np.stack([npe.jx_action_distribution(temperature, pol[i], legal[i]) for i in range(0, n, 10)])
        np.testing.assert_allclose(got[::10], want, rtol=0, atol=TOL)
        np.testing.assert_allclose(got.sum(axis=1), 1.0, atol=4e-6)
        a = acts.cpu().numpy()
        for i in range(0, n, 7):
            u = np.float32(oracle.philox_word(77, 1000 + i, int(sc0[i]), 4) >> 8) * np.float32(1.0 / 16777216.0)
            c, pick, last = np.float32(0), None, 0
            for k in range(4):
                if got[i, k] > 0:
                    c = np.float32(c + got[i, k])
                    last = k
                    if pick is None and u < c:
                        pick = k
            assert a[i] == (last if pick is None else pick)
        has = legal.any(axis=1)
        assert legal[has, a[has]].all()               # never an illegal action while one is legal
        if temperature == 0.0:                        # greedy: argmax of the masked policy
            sel = has & (np.where(legal, pol, 0).sum(axis=1) > 1e-6)
            np.testing.assert_array_equal(a[sel], np.where(legal, pol, -1)[sel].argmax(axis=1))


def test_draw_frequencies_follow_the_policy():
    import torch
    from simulate_2048_b200 import ops
    n = 400000
    boards = torch.full((n,), 0x0000000000001020, dtype=torch.int64, device='cuda')   # every action legal? no: see mask
    boards[:] = 0x0000100000020010                                                   # scattered tiles: all four legal
    pol = torch.tensor([0.1, 0.2, 0.3, 0.4], dtype=torch.float32, device='cuda').repeat(n, 1)
    acts = torch.empty(n, dtype=torch.uint8, device='cuda')
    ops.selfplay_step(boards, torch.zeros(n, dtype=torch.int32, device='cuda'),
                      torch.zeros(n, dtype=torch.uint8, device='cuda'),
                      torch.zeros(n, dtype=torch.float32, device='cuda'), pol, temperature=1.0, seed=5,
                      actions_out=acts)
    freq = np.bincount(acts.cpu().numpy(), minlength=4) / n
    np.testing.assert_allclose(freq, [0.1, 0.2, 0.3, 0.4], atol=0.004)


def test_recorded_policy_self_play_and_trajectories(oracle):
    """A whole self-play run through BatchedEnv.step_policy (one graph launch per iteration): the
    recorded trajectory replays through the oracle, policies / values are the ones fed in, the
    Trajectory records (replay_buffer.py:35-68) agree with the buffers and the final states."""
    import torch
    from simulate_2048_b200.selfplay import BatchedEnv, Trajectory, TrajectoryBuffer
    n, cap = 384, 500
    env = BatchedEnv(n, seed=21, graph=True)
    res = env.reset()
    key = env._key
    buf = TrajectoryBuffer(cap, n)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(0)
    fed_p, fed_v = [], []
    while buf.t < cap:
        p = torch.rand((n, 4), device='cuda', generator=gen) + 0.05
        v = torch.rand(n, device='cuda', generator=gen) * 100
        fed_p.append(p.cpu().numpy())
        fed_v.append(v.cpu().numpy())
        res = env.step_policy(p, v, temperature=1.0, record=buf)
        if buf.t % 16 == 0 and bool(res.done.all().item()):
            break
    steps = buf.t
    assert int(buf.t_dev.item()) == steps and bool(env.state.done.all().item())
    st = (oracle.reset(n, seed=key), np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    acts = buf.actions[:steps].cpu().numpy()
    obs_rec, legal_rec = buf.obs[:steps].cpu().numpy(), buf.legal[:steps].cpu().numpy()
    rew_rec = buf.rewards[:steps].cpu().numpy()
    for t in range(steps):
        np.testing.assert_array_equal(obs_rec[t], oracle.observe(st[0])[0])
        np.testing.assert_array_equal(legal_rec[t], ((oracle.mask(st[0])[:, None] >> np.arange(4)) & 1).astype(bool))
        np.testing.assert_array_equal(buf.active[t].cpu().numpy(), 1 - st[2])
        alive = st[2] == 0
        assert legal_rec[t][alive, acts[t][alive]].all()
        st, rew, _, _, _ = oracle.env_step(*st, acts[t], None, seed=key)
        np.testing.assert_array_equal(rew_rec[t], rew)
    np.testing.assert_array_equal(buf.lengths.cpu().numpy(), st[1])
    np.testing.assert_array_equal(env.board.cpu().numpy().view(np.uint64), st[0])
    np.testing.assert_array_equal(buf.policies[:steps].cpu().numpy(), np.stack(fed_p))
    np.testing.assert_array_equal(buf.values[:steps].cpu().numpy(), np.stack(fed_v))
    trajs = buf.to_trajectories(env)
    assert len(trajs) == n and all(isinstance(tr, Trajectory) for tr in trajs)
    assert Trajectory._fields == ('observations', 'actions', 'rewards', 'policies', 'values', 'done',
                                  'total_reward', 'max_tile')        # replay_buffer.py:58-65
    for i in (0, 5, n - 1):
        tr, k = trajs[i], int(st[1][i])
        assert len(tr) == k and tr.observations.shape == (k, 16) and tr.policies.shape == (k, 4)
        np.testing.assert_array_equal(tr.actions, acts[:k, i].astype(np.int32))
        np.testing.assert_array_equal(tr.rewards, rew_rec[:k, i])
        assert tr.done and tr.total_reward == float(st[3][i]) == float(tr.rewards.sum())
        assert tr.max_tile == int(oracle.unpack(st[0][i:i + 1]).max())
    # n-step targets straight from the recorded values
    assert tuple(buf.returns(None, 0.999, 10).shape) == (steps, n)


def test_normalised_reward_output(oracle):
    """norm_rewards = log2(reward) / log2(2^16), 0 for reward 0 (utils/normalize.py:6), f32."""
    import torch
    from simulate_2048_b200 import ops
    n = 30000
    boards = synthetic_packed(n, seed=13, max_exp=9)
    actions = np.random.default_rng(1).integers(0, 4, n).astype(np.uint8)
    d = lambda a: torch.from_numpy(a).cuda()
    raw = torch.empty(n, dtype=torch.float32, device='cuda')
    norm = torch.empty(n, dtype=torch.float32, device='cuda')
    ops.selfplay_step(d(boards.view(np.int64)), torch.zeros(n, dtype=torch.int32, device='cuda'),
                      torch.zeros(n, dtype=torch.uint8, device='cuda'), torch.zeros(n, dtype=torch.float32, device='cuda'),
                      None, actions=d(actions), seed=3, rewards=raw, norm_rewards=norm)
    r = raw.cpu().numpy()
    want = np.where(r > 0, np.log2(np.maximum(r, 1).astype(np.float64)) / 16.0, 0.0)
    np.testing.assert_allclose(norm.cpu().numpy(), want, rtol=0, atol=TOL)
    assert (r > 0).sum() > 1000
    np.testing.assert_array_equal(r, oracle.step(boards, actions, None, seed=3)[1])


def test_dlpack_round_trip_is_zero_copy():
    import torch
    from simulate_2048_b200 import game, interop
    st = game.batched_reset(3, 1000)
    caps = interop.state_to_dlpack(st)
    back = interop.state_from_dlpack(caps)
    assert back.board.data_ptr() == st.board.data_ptr() and back.done.dtype == torch.bool
    obs = game.batched_get_observation(st)
    alias = interop.from_dlpack(interop.to_dlpack(obs))
    assert alias.data_ptr() == obs.data_ptr() and alias.shape == obs.shape
    st2, _, _, _ = game.batched_step(back, torch.zeros(1000, dtype=torch.uint8, device='cuda'), 9, inplace=True)
    assert torch.equal(st.board, st2.board)           # stepping through the alias moved the original
    with pytest.raises(ValueError):
        interop.to_dlpack(torch.zeros(3))


def test_selfplay_step_injected_spawns_equal_env_step_oracle(oracle):
    """The parity mode (injected spawns, explicit actions): g2048_selfplay_step is bit-identical to
    the oracle's env.py:79-133 step over a whole trajectory — state, reward, done, observation, legal
    mask, max tile — and records exactly what it was given."""
    import torch
    from simulate_2048_b200 import ops
    from simulate_2048_b200.selfplay import TrajectoryBuffer
    from tests.helpers import valid_spawn_bytes
    n, steps = 2048, 150
    boards = oracle.reset(n, seed=77)
    st = (boards.copy(), np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    d_b = torch.from_numpy(boards.view(np.int64)).cuda()
    d_sc = torch.zeros(n, dtype=torch.int32, device='cuda')
    d_dn = torch.zeros(n, dtype=torch.uint8, device='cuda')
    d_tr = torch.zeros(n, dtype=torch.float32, device='cuda')
    rew = torch.empty(n, dtype=torch.float32, device='cuda')
    obs = torch.empty((n, 16), dtype=torch.float32, device='cuda')
    legal = torch.empty((n, 4), dtype=torch.bool, device='cuda')
    mt = torch.empty(n, dtype=torch.int32, device='cuda')
    buf = TrajectoryBuffer(steps, n)
    rng = np.random.default_rng(8)
    for t in range(steps):
        actions = rng.integers(0, 4, n).astype(np.uint8)
        after = oracle.expand(st[0])['after'][np.arange(n), actions]
        spawns = valid_spawn_bytes(after, seed=t)
        st, o_rew, o_fl, o_obs, o_mt = oracle.env_step(*st, actions, spawns)
        ops.selfplay_step(d_b, d_sc, d_dn, d_tr, None, actions=torch.from_numpy(actions).cuda(),
                          spawns=torch.from_numpy(spawns).cuda(), record=buf, rewards=rew, obs=obs, legal=legal,
                          max_tile=mt)
        if t % 30 == 29 or t == steps - 1:
            np.testing.assert_array_equal(d_b.cpu().numpy().view(np.uint64), st[0])
            np.testing.assert_array_equal(d_sc.cpu().numpy(), st[1])
            np.testing.assert_array_equal(d_dn.cpu().numpy(), st[2])
            np.testing.assert_array_equal(d_tr.cpu().numpy(), st[3])
            np.testing.assert_array_equal(rew.cpu().numpy(), o_rew)
            np.testing.assert_array_equal(obs.cpu().numpy(), o_obs)
            np.testing.assert_array_equal(mt.cpu().numpy(), o_mt)
            np.testing.assert_array_equal(legal.cpu().numpy(), ((o_fl[:, None] >> np.arange(4)) & 1).astype(bool))
        np.testing.assert_array_equal(buf.actions[t].cpu().numpy(), actions)
    assert int(buf.t_dev.item()) == steps and st[2].sum() > 0
    np.testing.assert_array_equal(buf.lengths.cpu().numpy(), st[1])
