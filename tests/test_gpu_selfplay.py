"""BatchedEnv (one g2048_env_step launch per self-play iteration, optionally graph-replayed)
against the oracle's env.py:79-133 semantics: state, reward, done, next observation, next legal
mask, max tile — with Philox spawns and with injected spawns, graph on and off."""

import numpy as np
import pytest

from tests.helpers import valid_spawn_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('graph', [False, True])
@pytest.mark.parametrize('injected', [False, True])
def test_batched_env_matches_oracle(oracle, graph, injected):
    import torch
    from simulate_2048_b200.selfplay import BatchedEnv
    n, steps = 4096, 200  # config 5: N = 4096, max_trajectory_length = 200 (config.py:51)
    env = BatchedEnv(n, seed=11, graph=graph)
    first = env.reset()
    key = env._key
    boards = oracle.reset(n, seed=key)
    np.testing.assert_array_equal(env.board.cpu().numpy().view(np.uint64), boards)
    o_obs, o_mx = oracle.observe(boards)
    np.testing.assert_array_equal(first.observation.cpu().numpy(), o_obs)
    np.testing.assert_array_equal(first.legal_mask.cpu().numpy(),
                                  ((oracle.mask(boards)[:, None] >> np.arange(4)) & 1).astype(bool))
    st = (boards, np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    rng = np.random.default_rng(2)
    for t in range(steps):
        actions = rng.integers(0, 4, n).astype(np.uint8)
        spawns = None
        if injected:
            after = oracle.expand(st[0])['after'][np.arange(n), actions]
            spawns = valid_spawn_bytes(after, seed=t)
        st, rew, fl, obs, mt = oracle.env_step(*st, actions, spawns, seed=key)
        res = env.step(torch.from_numpy(actions).cuda(),
                       None if spawns is None else torch.from_numpy(spawns).cuda())
        if t % 25 == 24 or t == steps - 1:
            np.testing.assert_array_equal(env.board.cpu().numpy().view(np.uint64), st[0])
            np.testing.assert_array_equal(env.step_count.cpu().numpy(), st[1])
            np.testing.assert_array_equal(res.done.cpu().numpy(), st[2].astype(bool))
            np.testing.assert_array_equal(env.total_reward.cpu().numpy(), st[3])
            np.testing.assert_array_equal(res.reward.cpu().numpy(), rew)
            np.testing.assert_array_equal(res.observation.cpu().numpy(), obs)
            np.testing.assert_array_equal(res.max_tile.cpu().numpy(), mt)
            np.testing.assert_array_equal(res.legal_mask.cpu().numpy(), ((fl[:, None] >> np.arange(4)) & 1).astype(bool))
    assert st[2].sum() > 0


def test_recorded_self_play_replays_through_oracle(oracle):
    """play_recorded keeps the whole trajectory on the device; replaying the recorded actions
    through the oracle (same Philox key) must reproduce every recorded observation, legal mask and
    reward, lengths must equal the oracle's step counts, and the n-step targets computed from the
    buffers must match the reference formula per env."""
    import torch
    from oracle import np_engine as npe
    from simulate_2048_b200.selfplay import BatchedEnv, TrajectoryBuffer, play_recorded, uniform_legal_policy
    n, cap = 512, 400
    env = BatchedEnv(n, seed=3, graph=True)
    env.reset()
    buf = TrajectoryBuffer(cap, n)
    steps = play_recorded(env, uniform_legal_policy(seed=1), buf)
    assert 100 < steps <= cap and bool(env.state.done.all().item())
    key = env._key
    st = (oracle.reset(n, seed=key), np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.float32))
    obs_rec, legal_rec = buf.obs[:steps].cpu().numpy(), buf.legal[:steps].cpu().numpy()
    act_rec, rew_rec = buf.actions[:steps].cpu().numpy(), buf.rewards[:steps].cpu().numpy()
    active_rec = buf.active[:steps].cpu().numpy()
    for t in range(steps):
        o_obs, _ = oracle.observe(st[0])
        np.testing.assert_array_equal(obs_rec[t], o_obs)
        np.testing.assert_array_equal(legal_rec[t], ((oracle.mask(st[0])[:, None] >> np.arange(4)) & 1).astype(bool))
        np.testing.assert_array_equal(active_rec[t], ~st[2].astype(bool))
        # the policy only picks legal actions while the game is alive
        alive = ~st[2].astype(bool)
        assert legal_rec[t][alive, act_rec[t][alive]].all()
        st, rew, _, _, _ = oracle.env_step(*st, act_rec[t], None, seed=key)
        np.testing.assert_array_equal(rew_rec[t], rew)
    lengths = buf.lengths.cpu().numpy()
    np.testing.assert_array_equal(lengths, st[1])
    assert lengths.min() > 20 and abs(lengths.mean() - 118) < 15  # random play: ~118 moves per game
    values = torch.from_numpy(np.random.default_rng(0).normal(300, 50, (steps, n)).astype(np.float32)).cuda()
    targets = buf.returns(values, 0.999, 10, 0.5).cpu().numpy()
    for e in (0, 7, n - 1):
        L = int(lengths[e])
        want = npe.n_step_returns(rew_rec[:L, e].astype(np.float64), values[:L, e].cpu().numpy().astype(np.float64), 0.999, 10, 0.5)
        np.testing.assert_allclose(targets[:L, e], want, rtol=4e-7, atol=1e-5)
