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
