"""Self-play stepping on the device (BASELINE config 5).

``reinforce/training/self_play.py:314-368`` makes three calls per loop iteration —
``batched_get_observation``, ``batched_get_legal_actions``, ``batched_step`` — plus two host syncs.
``BatchedEnv`` gives the same information from ONE launch of ``g2048_env_step`` per iteration:
the kernel that applies the actions also writes the observation and legal mask of the NEW state,
which are exactly what the next iteration reads.  All buffers are allocated once; with
``graph=True`` the launch (and the copy of the caller's actions into the static input buffer) is
replayed from a CUDA graph, so an iteration costs one graph launch instead of ~10 torch ops.
"""

from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np  # noqa: F401  (type names in Trajectory)
import torch

from . import ops
from .game.env import GameState


class StepResult(NamedTuple):
    observation: torch.Tensor   # f32[N,16]  (core.py:346) of the new state
    legal_mask: torch.Tensor    # bool[N,4]  (core.py:275) of the new state
    reward: torch.Tensor        # f32[N]
    done: torch.Tensor          # bool[N]
    max_tile: torch.Tensor      # i32[N]


class BatchedEnv:
    """N parallel games with static device buffers.  The returned tensors are views of those
    buffers: they are overwritten by the next ``step`` (clone what must be kept)."""

    def __init__(self, n: int, seed: int = 0, device=None, graph: bool = True):
        self.n = n
        self.seed = seed
        self.device = torch.device('cuda' if device is None else device)
        dev = self.device
        self.board = torch.empty(n, dtype=torch.int64, device=dev)
        self.step_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self._done8 = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.total_reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._actions = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._spawns: Optional[torch.Tensor] = None
        self._reward = torch.empty(n, dtype=torch.float32, device=dev)
        self._obs = torch.empty((n, 16), dtype=torch.float32, device=dev)
        self._legal = torch.empty((n, 4), dtype=torch.bool, device=dev)
        self._max_tile = torch.empty(n, dtype=torch.int32, device=dev)
        self.policy_in = torch.full((n, 4), 0.25, dtype=torch.float32, device=dev)   # static inputs of step_policy
        self.values_in = torch.zeros(n, dtype=torch.float32, device=dev)
        self._use_graph = graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._policy_graphs: dict = {}
        self._episode = 0

    @property
    def state(self) -> GameState:
        return GameState(self.board, self.step_count, self._done8.view(torch.bool), self.total_reward)

    def reset(self) -> StepResult:
        """New games (env.py:176): two tiles each, counters cleared."""
        self._episode += 1
        key = (self.seed + 0x9E3779B97F4A7C15 * self._episode) & (2**64 - 1)
        self._key = key
        self.board.copy_(ops.reset(self.n, seed=key, device=self.device))
        self.step_count.zero_()
        self._done8.zero_()
        self.total_reward.zero_()
        self._reward.zero_()
        obs, mx = ops.observe(self.board)
        self._obs.copy_(obs)
        self._max_tile.copy_(torch.where(mx > 0, torch.ones_like(mx, dtype=torch.int32) << mx.to(torch.int32),
                                         torch.zeros_like(mx, dtype=torch.int32)))
        flags = ops.legal_flags(self.board)
        bits = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=self.device)
        self._legal.copy_((flags.unsqueeze(1) & bits) != 0)
        self._graph = None  # the Philox key changed
        self._policy_graphs = {}
        return self._result()

    def _result(self) -> StepResult:
        return StepResult(self._obs, self._legal, self._reward, self._done8.view(torch.bool), self._max_tile)

    def _launch(self) -> None:
        ops.env_step(self.board, self.step_count, self._done8, self.total_reward, self._actions,
                     self._spawns, seed=self._key, rewards=self._reward, obs=self._obs,
                     max_tile=self._max_tile, legal=self._legal)

    def step_policy(self, policy: torch.Tensor, values: Optional[torch.Tensor] = None, *,
                    temperature: float = 1.0, record: Optional['TrajectoryBuffer'] = None,
                    normalize: bool = False) -> StepResult:
        """One whole self-play iteration as ONE graph launch (``g2048_selfplay_step``): draw an
        action per env from ``policy`` f32[N,4] (core.py:381 sample_action: legal mask, renormalise,
        temperature), record the step into ``record`` at its device-side cursor (observation and
        legal mask of the current state, ``policy``, ``values``, action, reward), apply it, and
        return the next observation / legal mask / reward / done.  ``policy`` and ``values`` are
        copied into static buffers (pass ``env.policy_in`` / ``env.values_in`` themselves, filled in
        place by the network, to skip the copy).  ``normalize``: ``StepResult.reward`` is the
        reference's log-normalised reward (utils/normalize.py:6)."""
        if policy.data_ptr() != self.policy_in.data_ptr():
            self.policy_in.copy_(policy.reshape(self.n, 4), non_blocking=True)
        if values is not None and values.data_ptr() != self.values_in.data_ptr():
            self.values_in.copy_(values.reshape(self.n), non_blocking=True)
        key = (id(record), float(temperature), values is not None, bool(normalize))
        if record is not None:   # host mirror of the device cursor (rows beyond the capacity are not recorded)
            record.t = min(record.t + 1, record.capacity)

        def launch():
            ops.selfplay_step(self.board, self.step_count, self._done8, self.total_reward, self.policy_in,
                              values=self.values_in if values is not None else None,
                              temperature=temperature, seed=self._key, record=record,
                              rewards=None if normalize else self._reward,
                              norm_rewards=self._reward if normalize else None, obs=self._obs,
                              legal=self._legal, max_tile=self._max_tile, actions_out=self._actions)
        if not self._use_graph:
            launch()
            return self._result()
        entry = self._policy_graphs.get(key)
        if entry is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):   # capture records the launches, it does not run them
                launch()
            # the entry keeps `record` alive: the graph holds its device pointers, and a freed
            # buffer's id() could be handed to a different object
            entry = self._policy_graphs[key] = (g, record)
        entry[0].replay()
        return self._result()

    def step(self, actions: torch.Tensor, spawns: Optional[torch.Tensor] = None) -> StepResult:
        """env.py:211 batched_step + the next observation / legal mask, one launch.  ``spawns``
        (uint8[N]) injects the tiles; otherwise Philox keyed by (episode key, env, step_count)."""
        self._actions.copy_(actions.reshape(-1), non_blocking=True)
        if spawns is not None:
            if self._spawns is None:
                self._spawns = torch.empty(self.n, dtype=torch.uint8, device=self.device)
                self._graph = None
            self._spawns.copy_(spawns.reshape(-1), non_blocking=True)
        elif self._spawns is not None:
            self._spawns = None
            self._graph = None
        if not self._use_graph:
            self._launch()
            return self._result()
        if self._graph is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # capture changes no state: the captured launch is only recorded, not executed
            with torch.cuda.graph(g):
                self._launch()
            self._graph = g
        self._graph.replay()
        return self._result()


class Trajectory(NamedTuple):
    """The reference's per-game record (reinforce/training/replay_buffer.py:35-68): same fields, same
    order, NumPy arrays on the host — what ``ReplayBuffer.add`` takes."""

    observations: 'np.ndarray'   # float32 [T, 16]
    actions: 'np.ndarray'        # int32   [T]
    rewards: 'np.ndarray'        # float32 [T]
    policies: 'np.ndarray'       # float32 [T, 4]
    values: 'np.ndarray'         # float32 [T]
    done: bool = True
    total_reward: float = 0.0
    max_tile: int = 0

    def __len__(self) -> int:
        return len(self.observations)


class TrajectoryBuffer:
    """Device-side trajectory storage for N parallel games, time-major (the layout
    ``ops.nstep_returns`` reads): replaces the per-step ``np.array(...)`` host copies and the O(N)
    Python bookkeeping of ``self_play.py:346-368``.  Step t holds the observation and legal mask the
    policy saw, the search policy and root value it produced (self_play.py:349-353 ``policies`` /
    ``values``), the action drawn, the reward of that action and whether the game was still running
    (finished games carry reward 0 and do not count towards ``lengths``).  ``t_dev`` is the write
    cursor on the DEVICE: ``g2048_selfplay_step`` records at row ``t_dev`` and advances it, so a
    recorded iteration needs no host-side index."""

    def __init__(self, capacity: int, n: int, device=None):
        dev = torch.device('cuda' if device is None else device)
        self.capacity, self.n, self.t = capacity, n, 0
        self.obs = torch.zeros((capacity, n, 16), dtype=torch.float32, device=dev)
        self.legal = torch.zeros((capacity, n, 4), dtype=torch.bool, device=dev)
        self.actions = torch.zeros((capacity, n), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((capacity, n), dtype=torch.float32, device=dev)
        self.policies = torch.zeros((capacity, n, 4), dtype=torch.float32, device=dev)
        self.values = torch.zeros((capacity, n), dtype=torch.float32, device=dev)
        self.active = torch.zeros((capacity, n), dtype=torch.uint8, device=dev)
        self.lengths = torch.zeros(n, dtype=torch.int32, device=dev)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=dev)

    def clear(self) -> None:
        self.t = 0
        self.lengths.zero_()
        self.t_dev.zero_()

    def returns(self, values: Optional[torch.Tensor], discount: float, n_steps: int, td_lambda: float = 0.5) -> torch.Tensor:
        """n-step TD(lambda) targets float32[t,N] for the recorded steps (self_play.py:524);
        ``values`` float32[t,N] are the caller's value estimates (default: the recorded root values)."""
        v = self.values[:self.t] if values is None else values
        return ops.nstep_returns(self.rewards[:self.t], v.contiguous(), discount, n_steps, td_lambda,
                                 lengths=self.lengths)

    def to_trajectories(self, env: Optional['BatchedEnv'] = None) -> list:
        """One ``Trajectory`` per game (replay_buffer.py:35-68), cut at each game's length: ONE
        device-to-host copy per field, then NumPy slicing.  ``env`` supplies done / total_reward /
        max_tile of the final states (self_play.py:377-386)."""
        import numpy as np
        t = self.t
        obs = self.obs[:t].cpu().numpy()
        act = self.actions[:t].cpu().numpy().astype(np.int32)
        rew = self.rewards[:t].cpu().numpy()
        pol = self.policies[:t].cpu().numpy()
        val = self.values[:t].cpu().numpy()
        lengths = self.lengths.cpu().numpy()
        if env is not None:
            done = env._done8.cpu().numpy().astype(bool)
            total = env.total_reward.cpu().numpy()
            top = env._max_tile.cpu().numpy()
        out = []
        for i in range(self.n):
            k = int(lengths[i])
            out.append(Trajectory(obs[:k, i].copy(), act[:k, i].copy(), rew[:k, i].copy(), pol[:k, i].copy(),
                                  val[:k, i].copy(),
                                  bool(done[i]) if env is not None else True,
                                  float(total[i]) if env is not None else float(rew[:k, i].sum()),
                                  int(top[i]) if env is not None else 0))
        return out


def play_recorded(env: BatchedEnv, policy, buf: TrajectoryBuffer, max_steps: Optional[int] = None,
                  temperature: float = 1.0) -> int:
    """Self-play loop (self_play.py:314-368) without a host copy per step.  ``policy(StepResult)``
    runs on the device and returns either ``(probs f32[N,4], values f32[N] | None)`` — the search
    output; the action is then drawn inside ``g2048_selfplay_step`` — or ``uint8[N]`` actions (legacy
    callers).  Every iteration is one graph launch that samples, records and steps; the only host
    sync is one ``all done`` check every 16 steps.  Returns the number of recorded steps."""
    res = env._result()
    limit = buf.capacity if max_steps is None else min(max_steps, buf.capacity)
    buf.clear()
    while buf.t < limit:
        out = policy(res)
        if isinstance(out, tuple):
            probs, values = out
            res = env.step_policy(probs, values, temperature=temperature, record=buf)
        else:   # explicit actions: recorded as a one-hot policy
            probs = torch.nn.functional.one_hot(out.reshape(-1).to(torch.int64), 4).to(torch.float32)
            res = env.step_policy(probs, None, temperature=0.0, record=buf)
        if buf.t % 16 == 0 and bool(res.done.all().item()):
            break
    return buf.t


def uniform_legal_policy(seed: int = 0):
    """Random policy over the legal actions (the reference's evaluation baseline): returns the legal
    mask as an unnormalised policy, so the draw happens inside ``g2048_selfplay_step`` (Philox tag 4)
    and an iteration stays one graph launch.  ``seed`` is kept for callers of the earlier signature;
    the stream is keyed by the env's own Philox key."""
    def policy(res: StepResult):
        return res.legal_mask.to(torch.float32), None
    return policy


def expectimax(boards: torch.Tensor, value_fn, depth: int = 1, discount: float = 1.0,
               terminal_value: float = 0.0, widen: Optional[int] = None, seed: int = 0, _level: int = 0):
    """Batched depth-``depth`` expectimax on real boards:
        Q_d(b, a) = r(b, a) + discount * E_children[V_{d-1}(child)]
        V_d(b)    = max over the legal actions of Q_d(b, a), ``terminal_value`` if there is none
        V_0(b)    = value_fn(b)
    ``widen=None``: the exact expectation over all 2 E chance children of every afterstate (one
    ``expand`` per level + ``g2048_expect``; the number of boards grows ~60x per level).
    ``widen=k``: progressive widening (gameboard.py:174-244, docs/project.md:49) — k children per
    afterstate drawn from the spawn distribution (``g2048_sample_children``) and averaged
    (``g2048_expect_sampled``): an unbiased estimate whose tree grows 4k per level.
    ``value_fn(int64[M]) -> float32[M]`` runs once on all leaves (e.g. a value network; boards of
    illegal branches are 0 and their values are ignored).
    Returns (q float32[N,4] with -inf for illegal actions, best action int64[N], ExpandOut)."""
    if depth < 1:
        raise ValueError('depth must be >= 1')
    flat = boards.reshape(-1)

    def level_values(children: torch.Tensor) -> torch.Tensor:
        if depth == 1:
            return value_fn(children).to(torch.float32).contiguous()
        q_child, _, _ = expectimax(children, value_fn, depth - 1, discount, terminal_value, widen, seed, _level + 1)
        best = q_child.max(dim=1).values
        return torch.where(torch.isinf(best), torch.full_like(best, terminal_value), best).contiguous()

    if widen is None:
        ex = ops.expand(flat, children='compact')
        q = ops.expect(level_values(ex.children), ex, discount=discount)
    else:
        k = int(widen)
        ex = ops.expand(flat, children='none')
        kids = ops.sample_children(ex, k, seed=seed, draw_base=_level * 4 * k)
        q = ops.expect_sampled(level_values(kids.reshape(-1)), ex, k, discount=discount)
    return q, q.argmax(dim=1), ex


def expectimax_depth1(boards: torch.Tensor, value_fn, discount: float = 1.0):
    """``expectimax`` with one level: reward + discount * E[value_fn(chance child)]."""
    return expectimax(boards, value_fn, 1, discount)
