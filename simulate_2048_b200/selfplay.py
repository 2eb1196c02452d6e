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
        self._use_graph = graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
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
        return self._result()

    def _result(self) -> StepResult:
        return StepResult(self._obs, self._legal, self._reward, self._done8.view(torch.bool), self._max_tile)

    def _launch(self) -> None:
        ops.env_step(self.board, self.step_count, self._done8, self.total_reward, self._actions,
                     self._spawns, seed=self._key, rewards=self._reward, obs=self._obs,
                     max_tile=self._max_tile, legal=self._legal)

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


class TrajectoryBuffer:
    """Device-side trajectory storage for N parallel games, time-major (the layout
    ``ops.nstep_returns`` reads): replaces the per-step ``np.array(...)`` host copies and the O(N)
    Python bookkeeping of ``self_play.py:346-368``.  Step t holds the observation and legal mask
    the policy saw, the action it chose, the reward of that action and whether the game was
    already over (such steps carry reward 0 and do not count towards ``lengths``)."""

    def __init__(self, capacity: int, n: int, device=None):
        dev = torch.device('cuda' if device is None else device)
        self.capacity, self.n, self.t = capacity, n, 0
        self.obs = torch.zeros((capacity, n, 16), dtype=torch.float32, device=dev)
        self.legal = torch.zeros((capacity, n, 4), dtype=torch.bool, device=dev)
        self.actions = torch.zeros((capacity, n), dtype=torch.uint8, device=dev)
        self.rewards = torch.zeros((capacity, n), dtype=torch.float32, device=dev)
        self.active = torch.zeros((capacity, n), dtype=torch.bool, device=dev)
        self.lengths = torch.zeros(n, dtype=torch.int32, device=dev)

    def clear(self) -> None:
        self.t = 0
        self.lengths.zero_()

    def returns(self, values: torch.Tensor, discount: float, n_steps: int, td_lambda: float = 0.5) -> torch.Tensor:
        """n-step TD(lambda) targets float32[t,N] for the recorded steps (self_play.py:524);
        ``values`` float32[t,N] are the caller's value estimates (search root values)."""
        return ops.nstep_returns(self.rewards[:self.t], values.contiguous(), discount, n_steps, td_lambda,
                                 lengths=self.lengths)


def play_recorded(env: BatchedEnv, policy, buf: TrajectoryBuffer, max_steps: Optional[int] = None) -> int:
    """Self-play loop without a host copy per step: ``policy(StepResult) -> uint8[N] actions`` runs
    on the device, every step is appended to ``buf``.  The only host sync is one ``all done``
    check every 16 steps.  Returns the number of recorded steps."""
    res = env._result()
    limit = buf.capacity if max_steps is None else min(max_steps, buf.capacity)
    buf.clear()
    while buf.t < limit:
        t = buf.t
        alive = ~res.done
        buf.obs[t].copy_(res.observation)
        buf.legal[t].copy_(res.legal_mask)
        buf.active[t].copy_(alive)
        buf.lengths.add_(alive.to(torch.int32))
        actions = policy(res)
        buf.actions[t].copy_(actions.reshape(-1))
        res = env.step(actions)
        buf.rewards[t].copy_(res.reward)
        buf.t = t + 1
        if buf.t % 16 == 0 and bool(res.done.all().item()):
            break
    return buf.t


def uniform_legal_policy(seed: int = 0):
    """Random policy over the legal actions, on the device (the reference's evaluation baseline)."""
    gen: list = []

    def policy(res: StepResult) -> torch.Tensor:
        if not gen:
            g = torch.Generator(device=res.legal_mask.device)
            g.manual_seed(seed)
            gen.append(g)
        w = res.legal_mask.to(torch.float32)
        w[:, 0] += (w.sum(dim=1) == 0).to(torch.float32)  # finished games: any action (a no-op)
        return torch.multinomial(w, 1, generator=gen[0]).reshape(-1).to(torch.uint8)
    return policy


def expectimax(boards: torch.Tensor, value_fn, depth: int = 1, discount: float = 1.0,
               terminal_value: float = 0.0):
    """Batched depth-``depth`` expectimax on real boards:
        Q_d(b, a) = r(b, a) + discount * sum_children p * V_{d-1}(child)     (g2048_expect)
        V_d(b)    = max over the legal actions of Q_d(b, a), ``terminal_value`` if there is none
        V_0(b)    = value_fn(b)
    One ``expand`` per level (mask, four afterstates, every chance child), ``value_fn(int64[M]) ->
    float32[M]`` once on all leaves (e.g. a value network).  The number of boards grows ~60x per
    level.  Returns (q float32[N,4] with -inf for illegal actions, best action int64[N], ExpandOut)."""
    if depth < 1:
        raise ValueError('depth must be >= 1')
    ex = ops.expand(boards, children='compact')
    if depth == 1:
        values = value_fn(ex.children).to(torch.float32).contiguous()
    else:
        q_child, _, _ = expectimax(ex.children, value_fn, depth - 1, discount, terminal_value)
        best = q_child.max(dim=1).values
        values = torch.where(torch.isinf(best), torch.full_like(best, terminal_value), best).contiguous()
    q = ops.expect(values, ex, discount=discount)
    return q, q.argmax(dim=1), ex


def expectimax_depth1(boards: torch.Tensor, value_fn, discount: float = 1.0):
    """``expectimax`` with one level: reward + discount * E[value_fn(chance child)]."""
    return expectimax(boards, value_fn, 1, discount)
