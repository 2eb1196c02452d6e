/*
 * g2048.h — C ABI of the B200-native batched 2048 transition library (libg2048.so).
 *
 * This is the drop-in boundary for the game-transition hot path of
 * schalappe/simulate_2048.  The reference has no FFI of its own: its boundary is the
 * Python function surface of `twentyfortyeight/core/{gameboard,gamemove}.py`,
 * `twentyfortyeight/envs/twentyfortyeight.py` and `reinforce/game/{core,env}.py`.  Each
 * entry point below names the reference function(s) it replaces (paths relative to the
 * reference repository root); INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - Plain C, caller-owned buffers, no allocation on the device-pointer entry points,
 *     no exceptions: every function returns 0 (G2048_OK) or a negative G2048_E* code.
 *   - Pointers are DEVICE pointers unless the function name ends in `_host`.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *     All device entry points are asynchronous with respect to the host.
 *   - Re-entrant across host threads, streams and devices.
 *   State.  The library keeps (1) a per-device, read-only row table, built once under a lock at
 *     the first call on that device (or by g2048_init); (2) a pool of host-pipeline resources
 *     (4 streams + 2 Mi-env staging buffers, ~190 MB of device memory each) used by
 *     g2048_step_host — a call borrows one and returns it, concurrent callers get their own;
 *     (3) the table of blocks handed out by g2048_host_alloc; (4) two process-wide MEASUREMENT
 *     knobs, g2048_set_impl / g2048_set_tuning (atomics, read once per call; meant for benchmarks
 *     and tests, not for production callers).  g2048_shutdown releases (1) and (2).
 *
 * Board format (`uint64_t`, one board per word)
 *   cell (r,c), flat index i = 4r + c (row-major, the order of `argwhere`/`flatten` in
 *   gameboard.py:154,275 and core.py:206-213); nibble i (bits [4i,4i+4)) holds
 *   log2(tile), 0 = empty.  Row r is bits [16r,16r+16), column 0 is the low nibble.
 *   Tiles above 32768 (exponent > 15) are not representable: a 15+15 merge raises
 *   G2048_FLAG_OVERFLOW for that env (the stored nibble wraps to 0, the reward is exact).
 *
 * Actions: 0 left, 1 up, 2 right, 3 down (twentyfortyeight.py:19, gameboard.py:114,
 *   core.py:142); only the low two bits are used (rot90(k) is modulo 4).
 *
 * Spawn byte (injected-spawn mode), one per env:
 *   bit 7 clear: low nibble = rank of the target cell among the empty cells of the
 *                afterstate in row-major order (what both engines reduce to:
 *                `available_cells[chosen]` gameboard.py:284-287, `pos_idx` core.py:210-213);
 *   bit 7 set  : low nibble = absolute flat cell index;
 *   bit 4      : spawned tile is a 4 (else a 2).
 *   A rank >= number of empties, or an occupied absolute cell, spawns nothing and raises
 *   G2048_FLAG_BADSPAWN.
 *
 * Philox mode (spawns == NULL), stream layout v2: Philox4x32-10, key = (seed_lo, seed_hi) or the
 *   per-env key.  ONE block serves the four consecutive envs of a quad: counter =
 *   (quad_lo, quad_hi, step, tag) with quad = GLOBAL env id (env_offset + i) >> 2, and env e reads
 *   word e & 3 — results do not depend on grid size, GPU count or fusion, and a thread that owns
 *   four consecutive envs pays for one block.  The word w is read as a mixed-radix fraction:
 *     rank  = hi32(w  * n_empty)   spawn cell among the empties,   f1 = lo32(w  * n_empty)
 *     tile 4 iff hi32(f1 * 10) == 0 (p = 0.1),                      f2 = lo32(f1 * 10)
 *     rollout policy: the NEXT action is the hi32(f2 * n_legal)-th legal action of the new board.
 *   Tags: 0 = play step, 3 = first action of a rollout that starts from a given board,
 *   4 = the action draw of g2048_selfplay_step (step = the env's step_count),
 *   5 = a sampled chance child of g2048_sample_children (step = draw index),
 *   1 / 2 = free for callers of g2048_spawn.
 *
 * Game stream (g2048_reset and g2048_rollout), stream layout v3: the draws of ONE game are
 *   consecutive words of that env's own blocks — block b = Philox4x32-10 of counter
 *   (env_lo, env_hi, b, tag 6) under the same key, and word p of the stream is word p & 3 of block
 *   p >> 2 (env = GLOBAL env id).  Four consecutive draws of a game cost one block: a rollout lane,
 *   whose games are all at different steps, cannot share the quad layout's block with anyone.
 *     p = 0, 1 : first / second reset tile (digits 1, 2 as above); the first action of a rollout
 *                from reset is digit 3 of word 1;
 *     rollout from reset: play step k reads word k + 2;  rollout from given boards: word k
 *     (its first action is digit 3 of the quad-layout word of (env, step 0, tag 3)).
 *   The lock-step entry points (g2048_step, g2048_env_step, g2048_selfplay_step, g2048_spawn) use
 *   the quad layout; a rollout is therefore NOT the Philox-mode g2048_step sequence of the same
 *   seed — its trace (action, spawn cell, tile per step) replays through g2048_step with injected
 *   spawns and through the reference engine.
 *
 * Flags byte (per env):
 */
#ifndef G2048_H
#define G2048_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2048_ABI_VERSION 3

#define G2048_FLAG_LEFT 0x01u     /* legal-action mask of the NEW board, bit a = action a */
#define G2048_FLAG_UP 0x02u
#define G2048_FLAG_RIGHT 0x04u
#define G2048_FLAG_DOWN 0x08u
#define G2048_FLAG_LEGAL_MASK 0x0Fu
#define G2048_FLAG_DONE 0x10u     /* no legal action on the new board (core.py:305-321) */
#define G2048_FLAG_MOVED 0x20u    /* the action changed the board (core.py:258) */
#define G2048_FLAG_OVERFLOW 0x40u /* a 32768+32768 merge happened (not representable) */
#define G2048_FLAG_BADSPAWN 0x80u /* injected spawn byte was invalid; nothing spawned */

#define G2048_OK 0
#define G2048_EINVAL (-1)   /* NULL/misaligned pointer, bad enum */
#define G2048_ECUDA (-2)    /* a CUDA runtime call failed; see g2048_last_cuda_error */
#define G2048_ENODEV (-3)   /* no CUDA device / not an sm_100 device */
#define G2048_ERANGE (-4)   /* tile value not representable (pack) */

#define G2048_PHILOX_TAG_STEP 0u
#define G2048_PHILOX_TAG_START 3u
#define G2048_PHILOX_TAG_POLICY 4u
#define G2048_PHILOX_TAG_CHANCE 5u
#define G2048_PHILOX_TAG_GAME 6u    /* the game stream: counter (env_lo, env_hi, block, 6) */

/* children layout for g2048_expand */
#define G2048_CHILDREN_NONE 0
#define G2048_CHILDREN_DENSE 1   /* children[((i*4 + a)*16 + k)*2 + v], k-th empty, v: 0 -> 2, 1 -> 4;
                                    every slot is written: slots beyond child_count hold 0 */
#define G2048_CHILDREN_COMPACT 2 /* children[child_offsets[i] + ...], action-major, cell-major, value-minor */

/* which row-move implementation the step kernel uses (for the ncu comparison) */
#define G2048_IMPL_AUTO 0
#define G2048_IMPL_LUT 1      /* 57344-entry row table staged in shared memory by TMA bulk copy */
#define G2048_IMPL_REGISTER 2 /* pure register arithmetic, no table */
#define G2048_IMPL_GATHER 3   /* full table gathered from global memory through L1/L2 (no staging) */

int g2048_abi_version(void);
const char* g2048_error_string(int code);
const char* g2048_last_cuda_error(void); /* thread-local text of the last CUDA failure */

/* Build the per-device row table now (otherwise done lazily). device < 0 = current. */
int g2048_init(int device);
/* Free the row tables and pooled host pipelines of every device.  Must not run concurrently with
 * other calls; the library re-initialises itself lazily if it is used again. */
int g2048_shutdown(void);
/* Select the row-move implementation used by g2048_step (process-wide measurement knob). */
int g2048_set_impl(int impl);
/* Tuning knobs of the smem-table step kernel: threads per CTA (0 = automatic = 640, or 512, 640,
 * 768, 1024; one CTA per SM) and consecutive envs per thread (4, or 8 with 512 / 768 / 1024
 * threads).  Process-wide measurement knob, like g2048_set_impl. */
int g2048_set_tuning(int step_threads, int envs_per_thread);

/*
 * Raw transition: replaces gameboard.py:291 next_state (+ is_done :328, gamemove.py:45
 * legal_actions_mask on the result) and core.py:234 next_state / :275 / :305.
 *   boards_in[n], actions[n], spawns[n] or NULL (Philox), step_ctr[n] or NULL (then every
 *   env uses step_base), env_keys[n] or NULL (per-env 64-bit Philox key replacing seed),
 *   done_in[n] or NULL: a non-zero entry passes the env through (env.py:103-104):
 *   board unchanged, reward 0, DONE set.
 *   Outputs: boards_out[n] (may be the SAME pointer as boards_in: every thread reads its envs before
 *   it writes them, and the loads carry no read-only (.nc) hint; partial overlap is not allowed), rewards[n] (sum of merged tile values,
 *   0 for an illegal move), flags[n].
 */
int g2048_step(const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns,
               uint64_t seed, uint64_t env_offset, const uint32_t* step_ctr, uint32_t step_base,
               const uint64_t* env_keys, const uint8_t* done_in, uint64_t* boards_out,
               int32_t* rewards, uint8_t* flags, size_t n, void* stream);

/*
 * Stateful env step on a GameState (env.py:33 GameState, :79 step, :211 batched_step):
 * updates boards / step_count / done / total_reward in place; step_count increments even
 * for an illegal action (env.py:116); done envs pass through with reward 0.
 *   Philox counter step = step_count before the increment.
 *   Optional outputs (NULL to skip): rewards f32[n], flags u8[n], obs f32[n*16]
 *   (core.py:346 encode_observation of the new board), max_tile i32[n] (core.py:427,
 *   the tile VALUE), legal u8[n*4] (core.py:275 legal_actions_mask of the new board as one
 *   0/1 byte per action, i.e. a bool[n,4] array) — so that one launch yields everything the
 *   self-play loop reads before its next step (self_play.py:316-317,354).
 */
int g2048_env_step(uint64_t* boards, int32_t* step_count, uint8_t* done, float* total_reward,
                   const uint8_t* actions, const uint8_t* spawns, uint64_t seed,
                   uint64_t env_offset, const uint64_t* env_keys, float* rewards,
                   uint8_t* flags, float* obs, int32_t* max_tile, uint8_t* legal, size_t n,
                   void* stream);

/*
 * core.py:381 sample_action for a batch: the distribution built from `policy` f32[n*4] and
 * `legal` u8[n*4] (0/1 bytes, a bool[n,4] array) exactly as core.py:401-421 (mask, renormalise, uniform
 * over the legal actions below 1e-8 of mass, softmax(log(p + 1e-8) / temperature), argmax below
 * temperature 0.01; f32, tolerance 1e-6 against a NumPy evaluation), and one draw per env: the
 * inverse CDF at the top 24 bits of the Philox word of (seed, env_offset + i, step, tag 4), with
 * step = step_ctr[i] (i32[n], e.g. the GameState step_count) or step_base.  actions_out u8[n] and
 * probs_out f32[n*4] are each optional.  An action of probability 0 is never returned.
 */
int g2048_sample_action(const float* policy, const uint8_t* legal, float temperature, uint64_t seed,
                        uint64_t env_offset, const int32_t* step_ctr, uint32_t step_base,
                        uint8_t* actions_out, float* probs_out, size_t n, void* stream);

/*
 * One self-play iteration in ONE call (self_play.py:314-368: observation, legal mask, sample_action,
 * batched_step and the per-step trajectory appends), in place on a GameState:
 *   1. action: drawn from `policy` f32[n*4] exactly as core.py:381-424 sample_action builds its
 *      distribution (illegal actions masked with the legal mask of the CURRENT board, renormalised,
 *      uniform over the legal actions when the mass is below 1e-8, softmax(log(p + 1e-8) /
 *      temperature), argmax when temperature < 0.01), f32 arithmetic; the draw is the inverse CDF at
 *      the top 24 bits of the env's Philox word of (seed, env, step_count, tag 4).  `actions_in`
 *      u8[n] (optional) replaces the draw.
 *   2. record (optional, row t = *t_dev of time-major buffers [capacity][n], skipped when
 *      t >= capacity): tr_obs f32[.,16] and tr_legal u8[.,4] of the board the action was chosen on,
 *      tr_actions, tr_policies f32[.,4] (the `policy` input), tr_values (`values` f32[n] or 0),
 *      tr_active (0 for an env that was already done), lengths i32[n] += active, tr_rewards after
 *      the step.  *t_dev is incremented by a trailing one-thread kernel (same stream).
 *   3. step as g2048_env_step (spawns u8[n] injected or NULL for Philox tag 0), with the optional
 *      outputs rewards f32[n], norm_rewards f32[n] (twentyfortyeight/utils/normalize.py:6:
 *      log2(reward)/16, 0 for reward 0), obs f32[n*16], legal u8[n*4], max_tile i32[n] of the NEW
 *      state, actions_out u8[n], probs_out f32[n*4] (the distribution of step 1).
 * Floating point: the distribution agrees with an f32 NumPy evaluation to 1e-6 (logf / expf).
 */
int g2048_selfplay_step(uint64_t* boards, int32_t* step_count, uint8_t* done, float* total_reward,
                        const float* policy, const float* values, const uint8_t* actions_in,
                        const uint8_t* spawns, float temperature, uint64_t seed, uint64_t env_offset,
                        int32_t* t_dev, uint32_t capacity, float* tr_obs, uint8_t* tr_legal,
                        uint8_t* tr_actions, float* tr_rewards, float* tr_policies, float* tr_values,
                        uint8_t* tr_active, int32_t* lengths, float* rewards, float* norm_rewards,
                        float* obs, uint8_t* legal, int32_t* max_tile, uint8_t* actions_out,
                        float* probs_out, size_t n, void* stream);

/* Legal mask + done of boards (gamemove.py:45, core.py:275/:305); flags has the LEGAL/DONE bits. */
int g2048_mask(const uint64_t* boards, uint8_t* flags, size_t n, void* stream);
/* The same as arrays the reference's callers index directly: legal u8[n*4], one 0/1 byte per action
 * (core.py:275 legal_actions_mask / env.py:245 batched_get_legal_actions as bool[n,4]) and done u8[n]
 * (core.py:305 is_done: no legal action).  Either may be NULL. */
int g2048_legal_actions(const uint64_t* boards, uint8_t* legal, uint8_t* done, size_t n, void* stream);

/*
 * MCTS leaf expansion (gameboard.py:105 latent_state for the four actions, :132
 * after_state for every legal afterstate, gamemove.py:45):
 *   after[n*4] afterstates, rewards[n*4], mask[n] (LEGAL|DONE bits, bit 6 = any overflow),
 *   num_empty[n*4] (empties of each afterstate; the spawn probability of a child is
 *   {0.9, 0.1}/num_empty, gameboard.py:166), child_count[n*4] (2*num_empty for a legal
 *   action with empties, 1 for a legal action whose afterstate is full, 0 for an illegal one),
 *   children per `child_layout`; child_offsets[n] (exclusive prefix sum over envs of the
 *   per-env total child_count) is required for G2048_CHILDREN_COMPACT.
 *   Any output pointer may be NULL to skip it.
 */
int g2048_expand(const uint64_t* boards, uint64_t* after, int32_t* rewards, uint8_t* mask,
                 uint8_t* num_empty, uint8_t* child_count, uint64_t* children,
                 const int64_t* child_offsets, int child_layout, size_t n, void* stream);

/* The chance children alone, from the afterstates and child counts an earlier g2048_expand call
 * wrote (after[n*4], child_count[n*4]): gameboard.py:132 after_state of every legal afterstate.
 * A caller that sizes the compact children buffer from child_count (count, prefix sum, allocate)
 * thus pays for the four moves once.  child_layout is DENSE or COMPACT; child_offsets as above. */
int g2048_expand_children(const uint64_t* after, const uint8_t* child_count,
                          const int64_t* child_offsets, uint64_t* children, int child_layout,
                          size_t n, void* stream);

/*
 * g2048_expand with the compact children layout as ONE call without a host round trip: the
 * per-board child totals are scanned on the device (n < 2^17: a single kernel launch, tiles
 * chained by ticket + aggregate look-back; larger batches: moves with per-group totals, one prefix kernel, children that derive the offsets).
 *   child_offsets[n] is an OUTPUT (exclusive prefix sum of the per-board totals, what a caller of
 *   g2048_expand computes on the host side); children has room for `capacity` boards (> 0): a
 *   board whose children would run past it writes none; total_out[2] (device, optional) receives
 *   {total number of children, 1 if total > capacity else 0}.  The worst case is 120 per board
 *   (4 actions x 15 empty cells x 2 values).  Other outputs as g2048_expand (NULL to skip).
 */
int g2048_expand_compact(const uint64_t* boards, uint64_t* after, int32_t* rewards, uint8_t* mask,
                         uint8_t* num_empty, uint8_t* child_count, int64_t* child_offsets,
                         uint64_t* children, size_t capacity, uint64_t* total_out, size_t n,
                         void* stream);

/* Chance children of given boards (gameboard.py:132 after_state, :174 after_state_lazy):
 * children[n*32] dense (slot 2k + v = k-th empty cell in row-major order, v: 0 -> tile 2,
 * 1 -> tile 4; a full board yields children[0] = the board itself; every remaining slot is
 * written as 0, so the buffer needs no memset), num_empty[n], and
 * empty_cells[n] (optional): 16-bit mask, bit c = flat cell c is empty. */
int g2048_after_state(const uint64_t* boards, uint64_t* children, uint8_t* num_empty,
                      uint16_t* empty_cells, size_t n, void* stream);

/* New games: two spawns on the empty board (core.py:324 create_initial_board,
 * twentyfortyeight.py:84 reset).  spawns[n*2] injected or NULL for words 0, 1 of the env's game stream. */
int g2048_reset(uint64_t seed, uint64_t env_offset, const uint64_t* env_keys,
                const uint8_t* spawns, uint64_t* boards_out, size_t n, void* stream);

/* One tile on each board (gameboard.py:247 fill_cells with number_tile = 1, core.py:173
 * fill_cells): a full board is returned unchanged.  spawns[n] injected, or NULL for Philox with
 * counter (env, step_base, tag).  flags (optional) receives G2048_FLAG_BADSPAWN per env. */
int g2048_spawn(uint64_t* boards_io, const uint8_t* spawns, uint64_t seed, uint64_t env_offset,
                const uint64_t* env_keys, uint32_t step_base, uint32_t tag, uint8_t* flags,
                size_t n, void* stream);

/* Observation (core.py:346: exponent/16 as f32[16]) and max tile exponent (core.py:427). */
int g2048_observe(const uint64_t* boards, float* obs, uint8_t* max_exp, size_t n, void* stream);

/* One-hot observation (twentyfortyeight/utils/binary.py:52 encode_flatten with
 * encodage_size = width, used by envs/twentyfortyeight.py:66-67 with width 31):
 * out[(i*16 + c)*width + e] = 1 if log2(tile of cell c) == e (empty cell: e = 0) else 0.
 * 16 <= width <= 64.  out is u8[n*16*width]. */
int g2048_observe_onehot(const uint64_t* boards, uint8_t* out, int width, size_t n, void* stream);

/* Expectation over the chance children of every (board, action) — the value backup of an
 * expectimax / MCTS leaf expansion.  Inputs are the compact-layout outputs of
 * g2048_expand plus one value per child (e.g. a network's estimate, in child order):
 *   q[i*4 + a] = rewards[i*4 + a] + discount * sum_k (0.9 * v[2k] + 0.1 * v[2k+1]) / num_empty
 * with the probabilities of gameboard.py:13,132-171 (after_state: P(2) = 0.9, P(4) = 0.1, uniform
 * over the empty cells); an illegal action (child_count 0) gets `illegal_value`.  A count of 1 (an
 * afterstate without an empty cell, probability 1: gameboard.py:158-159) is honoured.
 * rewards may be NULL (treated as 0).  f32 arithmetic. */
int g2048_expect(const float* child_values, const uint8_t* child_count, const int64_t* child_offsets,
                 const int32_t* rewards, float discount, float illegal_value, float* q_out, size_t n,
                 void* stream);

/*
 * Progressive widening (gameboard.py:174 after_state_lazy + :205 generate_outcome, docs/project.md:49):
 * instead of all 2 E chance children of an afterstate, k children DRAWN from the spawn distribution
 * (cell uniform over the empties, tile 2 w.p. 0.9 / 4 w.p. 0.1, with replacement).
 *   after[n*4], child_count[n*4] as written by g2048_expand (0 = illegal action: its k slots are 0);
 *   children[(i*4 + a)*k + j] = draw j of (board i, action a), from the Philox word of
 *   (seed, env_offset + i, draw_base + a*k + j, tag 5), digits 1 and 2 as for a spawn.
 *   An afterstate without an empty cell is its own child (gameboard.py:158-159).
 * g2048_expect_sampled is the matching backup:
 *   q[i*4 + a] = rewards[i*4 + a] + discount * mean_j child_values[(i*4 + a)*k + j]
 * (an unbiased estimate of g2048_expect's expectation), `illegal_value` for an illegal action.
 */
int g2048_sample_children(const uint64_t* after, const uint8_t* child_count, uint32_t k, uint64_t seed,
                          uint64_t env_offset, uint32_t draw_base, uint64_t* children, size_t n, void* stream);
int g2048_expect_sampled(const float* child_values, const uint8_t* child_count, const int32_t* rewards,
                         uint32_t k, float discount, float illegal_value, float* q_out, size_t n, void* stream);

/* n-step TD(lambda) returns over device trajectory buffers (reinforce/training/self_play.py:524
 * compute_n_step_returns, same weighting and normalisation), time-major layout [T][N]:
 *   G_n(t) = sum_{i<n} discount^i r[t+i] + discount^n v[t+n]   (the bootstrap term only if t+n < len)
 *   out[t] = sum_n w_n G_n(t) / sum_n w_n,  n = 1..min(n_steps, len - t),
 *   w_n = (1-lambda) lambda^(n-1) for n < n_steps, lambda^(n-1) for n = n_steps;
 *   out[t] = v[t] when the weights sum to 0.
 * lengths i32[N] gives the trajectory length of each env (NULL: T for all); steps t >= length are
 * written as 0.  Arithmetic is f64, inputs and output f32. */
int g2048_nstep_returns(const float* rewards, const float* values, const int32_t* lengths,
                        double discount, int n_steps, double td_lambda, float* out, size_t T,
                        size_t n, void* stream);

/*
 * Fused random-policy rollout to termination, boards in registers:
 *   policy = uniform over the legal actions (digit 3 of the previous step's word), spawns from
 *   digits 1 / 2 of the env's game-stream word of that step (see "Game stream"): results depend on
 *   (seed, global env id) only — not on the grid, the GPU count or which lane played the game.
 *   boards_io[n]: start boards if `from_reset` == 0 (updated to the final boards), else
 *   output only (games start from g2048_reset's Philox boards of the same seed).
 *   score_out i64[n] (sum of rewards), steps_out i32[n], maxexp_out u8[n] — any may be NULL.
 *   stats_out u64[G2048_STATS_WORDS] (device, accumulated atomically; zero it first):
 *     [0] episodes, [1] total steps, [2] total score, [3] max score, [4] max steps,
 *     [5] episodes stopped by max_steps, [6] overflow count, [7] reserved,
 *     [8 + e] episodes whose max tile exponent is e (e = 0..15).
 *   trace (optional, u8[n * trace_stride]): byte per step = action | spawn_cell << 2 |
 *     spawn_is_four << 6 | 0x80, zero-terminated — for replay through the reference.
 */
#define G2048_STATS_WORDS 24
int g2048_rollout(uint64_t seed, uint64_t env_offset, size_t n, uint32_t max_steps,
                  int from_reset, uint64_t* boards_io, int64_t* score_out, int32_t* steps_out,
                  uint8_t* maxexp_out, uint64_t* stats_out, uint8_t* trace,
                  uint32_t trace_stride, void* stream);

/* int32 tile values [n*16] <-> packed boards.  `status` (device int32, optional) is set
 * to G2048_ERANGE if any value is not 0 or a power of two in [2, 32768]. */
int g2048_pack(const int32_t* tiles, uint64_t* boards, int32_t* status, size_t n, void* stream);
int g2048_unpack(const uint64_t* boards, int32_t* tiles, size_t n, void* stream);

/* Copy the device row table (65536 x {result u16, reward u32, moved u8}) to HOST buffers. */
int g2048_row_table_host(uint16_t* result, uint32_t* reward, uint8_t* moved);

/*
 * HOST-buffer transition (the reference-facing call: host ndarray in, host ndarray out).
 * Same semantics as g2048_step; pointers are host memory (pinned memory from
 * g2048_host_alloc gives full PCIe bandwidth, pageable memory works).  The batch is cut
 * into chunks that move H2D, compute and move D2H on separate streams so both PCIe
 * directions and the kernel overlap.  Synchronous: returns when outputs are in place.
 */
int g2048_step_host(const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns,
                    uint64_t seed, uint64_t env_offset, uint32_t step_base,
                    uint64_t* boards_out, int32_t* rewards, uint8_t* flags, size_t n);
/* The copies of g2048_step_host alone (same stages, same streams, no kernel): the PCIe ceiling of
 * that call on this box, for reporting e2e against it. */
int g2048_step_host_copies(const uint64_t* boards_in, const uint8_t* actions, const uint8_t* spawns,
                           uint64_t* boards_out, int32_t* rewards, uint8_t* flags, size_t n);
/* Pinned host memory for the CURRENT device: placed on the NUMA node the GPU hangs off when the
 * process is allowed to (mmap + mbind(MPOL_PREFERRED) + first touch + cudaHostRegister; environment
 * G2048_HOST_NUMA=0 or a refusal falls back to cudaHostAlloc).  Free with g2048_host_free. */
int g2048_host_alloc(void** ptr, size_t bytes);
int g2048_host_free(void* ptr);
/* Diagnostics: NUMA node of the current GPU and of the page at `ptr` (-1 = unknown). */
int g2048_host_numa_info(const void* ptr, int* gpu_node, int* mem_node);

#ifdef __cplusplus
}
#endif
#endif /* G2048_H */
