#!/usr/bin/env python
"""bench.py — env-steps/s of the batched 2048 transition on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, N > 1)

A "step" is one pass of the hot path (g2048_step) over one batch of synthetic boards.
Workload: BASELINE.json configs[1] ("single batched step on 1M synthetic random boards") scaled
to 2^26 boards per GPU per launch — the 1M set tiled 64x with fresh actions/spawns — because a
1M batch (23 MB) sits in L2 and lasts about one launch latency; the 1M case
is reported beside it as `latency_1m`.  Each rank owns its own 2^26 envs ("weak" scaling, no
collective on the step path).

One JSON line on stdout (rank 0):
  value      whole-job env-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through g2048_step_host: pinned HOST buffers in, HOST buffers out,
             H2D and D2H copies inside the timed region
  roofline   algorithmic bytes per launch / mean launch time vs the measured HBM copy peak
  cpu_baseline  the CPU oracle port timed on this box's host cores (rank 0, N = 1)
plus, inside `roofline.other_kernels` (a key the driver's parse keeps), the Philox-mode step and
BASELINE configs[3] — 2^24 fused rollouts sharded over ALL ranks with the statistics collective in
the timed region, checked against an oracle-pinned vector — and secondary measurements (1M latency,
expand, self-play loop, ...).
`--impl reference` times the REAL reference engine (oracle/_ref, staged by oracle/make_ref.py; one
Python call per env as its users run it) on all host cores and prints the same line shape; the
NumPy port is the fallback when the reference was never staged (`kind` says which).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_INJECTED = 23  # read board 8 + action 1 + spawn 1; write board 8 + reward 4 + flags 1
BYTES_PHILOX = 22
ROLLOUT_INSTR_PER_STEP = 163.6   # thread instructions per env-step of rollout_kernel (ncu, profiles/r02_rollout_game_stream_ncu_summary.txt)
ROLLOUT_SLOTS_PER_STEP = 172.9   # warp issue slots x 32 per env-step (same capture; 30.28 of 32 lanes active)
METRIC = 'env_steps_per_sec'
UNIT = 'env-steps/s'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    """(GB/s, 'measured'|'fallback')"""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# --------------------------------------------------------------------------------------------
# synthetic workload (SURVEY.md §8d generator, seed 2048)
# --------------------------------------------------------------------------------------------

def synthetic_packed(n: int, seed: int = 2048) -> np.ndarray:
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 17, n)
    keys = rng.random((n, 16))
    ex = rng.integers(1, 12, (n, 16))
    order = np.argsort(keys, axis=1, kind='stable')
    ranks = np.empty_like(order)
    np.put_along_axis(ranks, order, np.arange(16)[None, :].repeat(n, 0), axis=1)
    vals = np.take_along_axis(ex, ranks, axis=1)
    exps = np.where(ranks < cnt[:, None], vals, 0).astype(np.uint64)
    return np.bitwise_or.reduce(exps << (4 * np.arange(16, dtype=np.uint64))[None, :], axis=1)


# --------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------

REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap',
           0x8: 'hw_slowdown', 0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown',
           0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}


class ClockSampler:
    def __init__(self, index: int, period: float = 0.004):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log('clock sampling unavailable:', e)
            self.nv = None

    def sample(self):
        """One reading of SM clock, throttle reasons and power (an NVML query takes milliseconds
        while the GPU is busy, so it must stay off the launching thread: sampling from inside the
        timed loop drained the launch queue and cost 40 % of the measured throughput)."""
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for b, name in REASONS.items():
                if bits & b and name != 'gpu_idle':
                    self.reasons.add(name)
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': int(statistics.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples),
                'power_w_max': round(max(self.power), 1) if self.power else None}


# --------------------------------------------------------------------------------------------
# CPU baselines (oracle; the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------

def cpu_engine():
    """('reference', oracle.ref_engine) when the real reference is staged under oracle/_ref, else
    ('port', oracle.np_engine)."""
    from oracle import make_ref
    if make_ref.available():
        from oracle import ref_engine
        return 'reference', ref_engine
    from oracle import np_engine
    return 'port', np_engine


def engine_label(kind):
    return ('the reference engine itself (oracle/_ref/twentyfortyeight, staged by oracle/make_ref.py)' if kind == 'reference'
            else 'NumPy port of the reference engine (oracle/np_engine.py; the reference was not staged)')


def _np_worker(args):
    tiles, actions, seed = args
    _, eng = cpu_engine()
    t0 = time.perf_counter()
    eng.step_boards(tiles, actions, seed)
    return time.perf_counter() - t0


def numpy_port_rate(packed: np.ndarray, actions: np.ndarray, per_worker: int, workers: int, pool=None):
    """env-steps/s of the NumPy port (one next_state call per env) over `workers` processes."""
    from oracle import c_oracle
    total = per_worker * workers
    idx = np.arange(total) % packed.size
    tiles = c_oracle.unpack(packed[idx]).astype(np.int64)
    jobs = [(tiles[w * per_worker:(w + 1) * per_worker], actions[idx][w * per_worker:(w + 1) * per_worker], w)
            for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _np_worker(jobs[0])
    else:
        pool.map(_np_worker, jobs)
    dt = time.perf_counter() - t0
    return total / dt, dt


def _np_expand_worker(tiles):
    """Config 3 on the CPU engine: mask, four afterstates + rewards, every chance child."""
    kind, eng = cpu_engine()
    if kind == 'reference':
        return eng.expand_boards(tiles)
    children = 0
    for b in tiles:
        mask = eng.legal_actions_mask(b)
        for a in range(4):
            after, _ = eng.latent_state(b, a)
            if mask[a]:
                children += len(eng.after_state(np.array(after)))
    return children


def numpy_expand_rate(packed: np.ndarray, per_worker: int, workers: int, pool):
    from oracle import c_oracle
    total = per_worker * workers
    tiles = c_oracle.unpack(packed[:total]).astype(np.int64).reshape(total, 4, 4)
    t0 = time.perf_counter()
    kids = sum(pool.map(_np_expand_worker, [tiles[w * per_worker:(w + 1) * per_worker] for w in range(workers)]))
    dt = time.perf_counter() - t0
    return {'boards_per_sec': total / dt, 'children_per_sec': kids / dt, 'cores': workers,
            'kind': cpu_engine()[0],
            'sample': f'{total} boards of the config-2 set: legal_actions_mask + latent_state x4 + after_state '
                      f'of every legal afterstate, one Python call each, {dt:.1f} s'}


def _games_worker(args):
    first, count = args
    return cpu_engine()[1].play_random_games(first, count)[:2]


def config1_games(cores, n_games=1024):
    """BASELINE configs[0]: 1024 random-policy games to game-over on the reference NumPy engine
    (TwentyFortyEight.reset(seed=g) / .step, uniform-legal policy), one
    process per core."""
    import multiprocessing as mp
    per = (n_games + cores - 1) // cores
    jobs = [(g, min(per, n_games - g)) for g in range(0, n_games, per)]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(cores) as pool:
        res = pool.map(_games_worker, jobs)
    dt = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    score = sum(r[1] for r in res)
    return {'games': n_games, 'env_steps': steps, 'seconds': dt, 'env_steps_per_sec': steps / dt,
            'mean_length': steps / n_games, 'mean_score': score / n_games, 'cores': cores,
            'kind': cpu_engine()[0],
            'note': 'play_random_games: TwentyFortyEight.reset(seed=g) + .step with a uniform policy over '
                    'legal_actions, ' + engine_label(cpu_engine()[0])}


def cpu_baseline_block(packed, actions, cores):
    """Bounded sample (~10-30 s of CPU work): NumPy port on all cores + the C oracle."""
    import multiprocessing as mp
    from oracle import c_oracle
    c_oracle.build()
    with mp.get_context('fork').Pool(cores) as pool:
        probe, _ = numpy_port_rate(packed, actions, 1 << 11, cores, pool)   # untimed sizing probe
        per_worker = int(max(1 << 12, min(1 << 18, probe / cores * 30.0)))   # ~10 s of work per core (the probe includes pool start-up)
        rate, dt = numpy_port_rate(packed, actions, per_worker, cores, pool)
        cfg3 = numpy_expand_rate(packed, 2048, cores, pool)
    rate1, _ = numpy_port_rate(packed, actions, 1 << 12, 1)
    n_c = min(packed.size, 1 << 22)
    t0 = time.perf_counter()
    c_oracle.step(packed[:n_c], actions[:n_c], None, seed=2048, threads=cores)
    c_rate = n_c / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    c_oracle.step(packed[: n_c // 8], actions[: n_c // 8], None, seed=2048, threads=1)
    c_rate1 = (n_c // 8) / (time.perf_counter() - t0)
    cfg1 = config1_games(cores)
    kind = cpu_engine()[0]
    return {
        'config1_numpy_engine_1024_games': cfg1,
        'value': rate, 'unit': UNIT, 'cores': cores, 'kind': kind,
        'sample': f'{engine_label(kind)}: next_state, one Python call '
                  f'per env, {per_worker} boards x {cores} processes of the config-2 set, {dt:.1f} s',
        'single_core_value': rate1,
        'config3_numpy_expand': cfg3,
        'jax_on_cpu': 'unavailable in this image (jax / jaxlib are not installed and there is no network); '
                      'the JAX engine is covered by a NumPy restatement of its semantics, not timed as "JAX"',
        'c_oracle': {'value': c_rate, 'single_core_value': c_rate1, 'cores': cores, 'unit': UNIT,
                     'sample': f'oracle/g2048_oracle.c orc_step, {n_c} boards, pthreads'},
    }


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from oracle import c_oracle
    c_oracle.build()
    cores = host_cores()
    packed = synthetic_packed(1 << 16)
    actions = np.random.default_rng(0).integers(0, 4, packed.size).astype(np.uint8)
    total_steps = args.steps + args.warmup
    with mp.get_context('fork').Pool(cores) as pool:
        probe, _ = numpy_port_rate(packed, actions, 512, cores, pool)   # untimed sizing probe
        budget = 90.0  # seconds for the whole run
        per_worker = int(max(256, min(1 << 14, probe * budget / total_steps / cores)))
        for _ in range(args.warmup):
            numpy_port_rate(packed, actions, per_worker, cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            numpy_port_rate(packed, actions, per_worker, cores, pool)
        dt = time.perf_counter() - t0
    envs = per_worker * cores * args.steps
    value = envs / dt
    kind = cpu_engine()[0]
    sample = (f'{per_worker * cores} boards per step ({per_worker} x {cores} processes) of the config-2 '
              f'synthetic set, one next_state call per board on {engine_label(kind)}')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int64',
        'data': 'synthetic',
        'config': {'workload': 'single batched 2048 step (BASELINE configs[1]) on the CPU engine; bounded sample of '
                               'the same workload as the GPU arm (which steps 2^26 boards per step): ' + sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist

    from simulate_2048_b200 import ops

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus and world > 1:
        log(f'warning: --gpus {args.gpus} but WORLD_SIZE {world}')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    host_group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        # ranks that wait for rank 0's rank-local measurements block on the HOST (gloo), not inside
        # a spinning NCCL kernel
        host_group = dist.new_group(backend='gloo')
    ops.init(local)
    if os.environ.get('G2048_STEP_THREADS'):
        ops.set_tuning(int(os.environ['G2048_STEP_THREADS']), int(os.environ.get('G2048_STEP_ENVS', '4')))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: the 1M config-2 set tiled to B boards, fresh actions / spawns per copy ----
    B = 1 << args.boards_log2
    base_np = synthetic_packed(1 << 20)
    base = torch.from_numpy(base_np.view(np.int64)).to(dev)
    boards = base.repeat(B // base.numel()).contiguous()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    actions = torch.randint(0, 4, (B,), dtype=torch.uint8, device=dev, generator=gen)
    # valid injected spawns: rank uniform over the empties of the afterstate, four w.p. 0.1
    spawns = torch.empty(B, dtype=torch.uint8, device=dev)
    chunk = 1 << 22
    for lo in range(0, B, chunk):
        ex = ops.expand(boards[lo:lo + chunk], children='none')
        ne = ex.num_empty.gather(1, actions[lo:lo + chunk].to(torch.int64).unsqueeze(1)).squeeze(1)
        r = torch.randint(0, 1 << 30, (chunk,), device=dev, generator=gen) % ne.clamp(min=1).to(torch.int64)
        four = (torch.rand(chunk, device=dev, generator=gen) < 0.1).to(torch.int64) << 4
        spawns[lo:lo + chunk] = (r | four).to(torch.uint8)
        del ex
    out = ops.StepOut(torch.empty_like(boards), torch.empty(B, dtype=torch.int32, device=dev),
                      torch.empty(B, dtype=torch.uint8, device=dev))
    env_offset = rank * B

    def timed(fn, steps, warmup, collective=True):
        """CUDA events on the launching (current) stream around every step; barrier + sync on
        both sides; returns (total seconds = max over ranks, mean launch seconds).  The
        rank-local secondary measurements pass collective=False: they run on rank 0 only, so a
        barrier there would pair with the other ranks' final barrier and hang the job."""
        sync = barrier if collective else torch.cuda.synchronize
        for _ in range(warmup):
            fn()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sync()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for a, b in ev:
            a.record()
            fn()
            b.record()
        t1.record()
        sync()
        total = t0.elapsed_time(t1) / 1e3
        if collective:
            total = max_over_ranks(total)
        per = [a.elapsed_time(b) / 1e3 for a, b in ev]
        timed.last = per
        return total, sum(per) / len(per), min(per)

    # ---- primary: injected spawns (the mode that is bit-exact against the reference) ----
    step_inj = lambda: ops.step(boards, actions, spawns, env_offset=env_offset, out=out)  # noqa: E731
    step_phx = lambda: ops.step(boards, actions, None, seed=2048, env_offset=env_offset, out=out)  # noqa: E731
    # the sampler runs through both timed regions of the step kernel (injected, then Philox): an NVML
    # query takes ~10 ms under load, so a short run yields only a few samples per region
    clocks = ClockSampler(local)
    clocks.__enter__()
    total, mean_launch, best_launch = timed(step_inj, args.steps, args.warmup)
    value = world * B * args.steps / total
    peak, peak_src = measured_peak()
    achieved = BYTES_INJECTED * B / mean_launch / 1e9
    traffic = args.ncu_traffic
    if traffic is None:
        try:  # dram bytes per env measured by `ncu --set full` (profiles/r02_traffic.json)
            with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
                traffic = json.load(f)['dram_bytes_per_env'] * B
        except Exception:
            traffic = None

    p_total, p_mean, _ = timed(step_phx, max(args.steps // 2, 5), 3)
    clocks.__exit__(None, None, None)
    philox = {'value': world * B * max(args.steps // 2, 5) / p_total, 'unit': UNIT,
              'bytes_per_env_step': BYTES_PHILOX, 'achieved_gbs': BYTES_PHILOX * B / p_mean / 1e9,
              'frac': BYTES_PHILOX * B / p_mean / 1e9 / peak, 'mean_launch_ms': p_mean * 1e3,
              'kernel': 'step_lut_kernel<true, 640, 1>',
              'stream': 'Philox4x32-10, one block per env quad (include/g2048.h "Philox mode")'}

    # ---- BASELINE configs[3]: fused rollouts to termination, 2^24 envs in total sharded over ALL
    # ranks (strong scaling), the statistics collective inside the timed region, checked against the
    # oracle-pinned vector of tests/golden/rollout_stats.json ----
    from simulate_2048_b200 import parallel
    n_roll = 1 << args.rollout_log2
    ROLL_SEED = 2048

    last = [None]

    def sharded():   # rollout kernel + statistics kernel + the collective; the reduced vector stays on the device
        last[0] = parallel.sharded_rollout(n_roll, seed=ROLL_SEED, rank=rank, world=world, device=dev, to_host=False)[1]
    sharded()
    r_stats = parallel.EpisodeStats.from_vector(last[0].cpu())
    try:
        with open(os.path.join(ROOT, 'tests', 'golden', 'rollout_stats.json')) as f:
            want_stats = json.load(f)['stats'].get(str(n_roll))
    except Exception:
        want_stats = None
    ro_total, ro_mean, ro_best = timed(sharded, 3, 1)
    r_stats = parallel.EpisodeStats.from_vector(last[0].cpu())    # of the LAST timed call, read after the timing
    got_stats = [int(x) & (2**64 - 1) for x in r_stats.vector]
    rollout = {
        'envs_total': n_roll, 'envs_per_gpu': n_roll // world, 'scaling': 'strong', 'seed': ROLL_SEED,
        'env_steps': r_stats.total_steps, 'mean_length': r_stats.total_steps / n_roll,
        'ms': ro_total / 3 * 1e3, 'best_ms_rank0': ro_best * 1e3, 'per_call_ms_rank0': [round(x * 1e3, 3) for x in timed.last],
        'env_steps_per_sec': r_stats.total_steps * 3 / ro_total, 'unit': UNIT,
        'collective': 'one all_gather_into_tensor of the 24-word statistics vector per rollout (NCCL), '
                      'inside the timed region' if world > 1 else 'none (1 GPU)',
        'stats_equal_oracle': (got_stats == want_stats) if want_stats is not None else None,
        'oracle_vector': 'tests/golden/rollout_stats.json (oracle/g2048_oracle.c orc_rollout_mt)',
        'summary': r_stats.summary(),
        'kernel': 'rollout_kernel<true, true, false, 640> + rollout_stats_kernel: two games per lane, one queue '
                  'per CTA, prefetched game ids, per-env game stream (one Philox4x32-10 block per four steps of a '
                  'game, include/g2048.h "Game stream")',
        # issue ceiling of one GPU: 148 SMs x 4 schedulers x 1 warp-instruction / clk x 32 lanes / instr per step
        'issue_ceiling_per_gpu': 148 * 4 * 1.965e9 * 32 / ROLLOUT_INSTR_PER_STEP,
        'issue_ceiling_per_gpu_warp_slots': 148 * 4 * 1.965e9 * 32 / ROLLOUT_SLOTS_PER_STEP,
        'instr_per_step': ROLLOUT_INSTR_PER_STEP, 'issue_slots_per_step': ROLLOUT_SLOTS_PER_STEP,
        'integer_pipe_utilisation': {'alu_pipe_pct': 71.8, 'fmaheavy_pipe_pct': 68.9, 'issue_active_pct': 71.6,
                                     'lanes_active_of_32': 30.28, 'kernel': 'rollout_kernel<1,1,0,640> on 2^22 envs'},
        'instr_per_step_source': 'stored: profiles/r02_rollout_game_stream_ncu_summary.txt (ncu --set full on 2^22 envs), not '
                                 'measured in this run',
    }
    rollout['frac_of_issue_ceiling'] = rollout['env_steps_per_sec'] / (world * rollout['issue_ceiling_per_gpu'])
    rollout['frac_of_issue_ceiling_warp_slots'] = rollout['env_steps_per_sec'] / (world * rollout['issue_ceiling_per_gpu_warp_slots'])

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': total / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64', 'data': 'synthetic',
        'config': {
            'workload': f'single batched 2048 step (BASELINE configs[1]) on 2^{args.boards_log2} synthetic '
                        'random boards per GPU per launch (the 1M-board set tiled, fresh actions and injected '
                        'spawns per copy)',
            'boards_per_gpu': B, 'spawn_mode': 'injected', 'bytes_per_env_step': BYTES_INJECTED,
            'l2': f'inputs {B * 10 / 1e6:.0f} MB + outputs {B * 13 / 1e6:.0f} MB per step, larger than the 126 MB L2; no flush needed',
            'parallelism': f'env-sharded x{world}, no collective on the step path',
        },
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'frac_of_spec_8000gbs': achieved / 8000.0,
                     'kernel': 'step_lut_kernel<false, %s, %s>' % (os.environ.get('G2048_STEP_THREADS', '640'),
                                                                    int(os.environ.get('G2048_STEP_ENVS', '4')) // 4),
                     'mean_launch_ms': mean_launch * 1e3,
                     'best_launch_ms': best_launch * 1e3,
                     'algorithmic_bytes_per_launch': BYTES_INJECTED * B,
                     'traffic_source': 'stored: ncu --set full DRAM bytes per env (profiles/r02_traffic.json) x envs; '
                                       'not measured in this run',
                     'other_kernels': {'philox_step': philox, 'rollout_sharded': rollout}},
        'clocks': clocks.summary(),
        'gpu_launches': args.steps,
    }

    # ---- e2e: pinned host buffers through g2048_step_host (H2D + kernel + D2H timed) ----
    Be = min(1 << args.e2e_log2, B)
    h_b = ops.pinned_empty(Be, torch.int64)     # g2048_host_alloc: pinned, on the GPU's NUMA node if allowed
    h_a = ops.pinned_empty(Be, torch.uint8)
    h_s = ops.pinned_empty(Be, torch.uint8)
    h_b[:] = boards[:Be].cpu().numpy()
    h_a[:] = actions[:Be].cpu().numpy()
    h_s[:] = spawns[:Be].cpu().numpy()
    h_out = (ops.pinned_empty(Be, torch.int64).view(np.uint64), ops.pinned_empty(Be, torch.int32),
             ops.pinned_empty(Be, torch.uint8))
    e_steps = max(3, min(args.steps, 10))

    def host_timed(fn):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            fn()  # synchronous
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)
    e_total = host_timed(lambda: ops.step_host(h_b, h_a, h_s, env_offset=env_offset, out=h_out))
    # the same copies on the same streams without the kernel: the PCIe / host-memory ceiling of this
    # call on this box at this number of ranks
    c_total = host_timed(lambda: ops.step_host_copies(h_b, h_a, h_s, h_out))
    gpu_node, mem_node = ops.host_numa_info(h_b)
    e_val, c_val = world * Be * e_steps / e_total, world * Be * e_steps / c_total
    line['e2e'] = {'value': e_val, 'unit': UNIT,
                   'h2d_bytes_per_step': Be * 10, 'd2h_bytes_per_step': Be * 13,
                   'boards_per_step': Be, 'steps': e_steps, 'ms_per_step': e_total / e_steps * 1e3,
                   'copy_ceiling': c_val, 'frac_of_copy_ceiling': e_val / c_val,
                   'copy_ceiling_gbs_per_gpu': Be * 23 * e_steps / c_total / 1e9,
                   'copy_ceiling_note': 'g2048_step_host_copies: the H2D and D2H copies of the same call on the same '
                                        'four streams, no kernel, all ranks at once (max over ranks)',
                   'host_buffers': {'alloc': 'g2048_host_alloc (mmap + mbind(MPOL_PREFERRED, GPU node) + cudaHostRegister)',
                                    'gpu_numa_node': gpu_node, 'buffer_numa_node': mem_node,
                                    'cpus_allowed': host_cores()},
                   'api': 'simulate_2048_b200.ops.step_host -> g2048_step_host (pinned host ndarrays in/out)'}
    del h_b, h_a, h_s, h_out

    # ---- secondary measurements (rank-local, short) ----
    if rank == 0 and not args.quick:
        extra = {}
        # 1M-board launches over rotating buffers (64 distinct 1M slices => cold in L2)
        m1 = 1 << 20
        views = [(boards[i * m1:(i + 1) * m1], actions[i * m1:(i + 1) * m1], spawns[i * m1:(i + 1) * m1],
                  ops.StepOut(out.boards[i * m1:(i + 1) * m1], out.rewards[i * m1:(i + 1) * m1],
                              out.flags[i * m1:(i + 1) * m1])) for i in range(B // m1)]
        k = [0]

        def one_m():
            b_, a_, s_, o_ = views[k[0] % len(views)]
            k[0] += 1
            ops.step(b_, a_, s_, out=o_)
        # back-to-back launches replayed from a CUDA graph, so the figure is device time per launch
        # (Python/ctypes call overhead is ~20 us per call and would otherwise dominate)
        for _ in range(len(views)):
            one_m()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(len(views)):
                    one_m()
        torch.cuda.synchronize()
        _, g_mean, g_min = timed(graph.replay, 20, 3, collective=False)
        lat_mean, lat_min = g_mean / len(views), g_min / len(views)
        extra['latency_1m'] = {'boards': m1, 'us_per_launch_mean': lat_mean * 1e6,
                               'us_per_launch_min': lat_min * 1e6, 'env_steps_per_sec': m1 / lat_mean,
                               'achieved_gbs': BYTES_INJECTED * m1 / lat_mean / 1e9,
                               'note': f'2^20-board launches (BASELINE configs[1] as written), {len(views)} rotating '
                                       'slices replayed back to back from one CUDA graph'}
        # LUT-in-shared-memory vs pure register arithmetic (north star: "the choice evidenced by ncu")
        ops.set_impl(2)
        r_total, r_mean, _ = timed(step_inj, 10, 3, collective=False)
        ops.set_impl(0)
        extra['register_arithmetic_variant'] = {'env_steps_per_sec': B / r_mean,
                                                'achieved_gbs': BYTES_INJECTED * B / r_mean / 1e9,
                                                'frac': BYTES_INJECTED * B / r_mean / 1e9 / peak}
        # config 3: MCTS leaf expansion (mask + 4 afterstates + all chance children), 65536 boards and 1M
        exp_out = {}
        for n_e in (65536, 1 << 20):
            eb = boards[:n_e].contiguous()
            ex_c = ops.expand(eb, children='compact')
            ex_d = ops.expand(eb, children='dense', zero_fill=False)
            n_children = int(ex_c.children.numel())
            reps = 40 if n_e == 65536 else 10

            def graph_time(fn, calls):
                # device time per call: `calls` back-to-back calls replayed from one CUDA graph
                # (a Python/ctypes call costs ~10 us, more than the kernels of a 65536-board call)
                fn()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                side_e = torch.cuda.Stream()
                with torch.cuda.stream(side_e):
                    with torch.cuda.graph(g, stream=side_e):
                        for _ in range(calls):
                            fn()
                torch.cuda.synchronize()
                _, mean_s, _ = timed(g.replay, 5, 3, collective=False)
                return mean_s / calls
            # compact: g2048_expand_compact — the child offsets are scanned on the device INSIDE the
            # timed call (no count pass, no host round trip)
            t_c = graph_time(lambda: ops.expand(eb, children='compact', out=ex_c), reps)
            t_d = graph_time(lambda: ops.expand(eb, children='dense', out=ex_d), reps)
            # the whole public call as a fresh caller makes it: allocation, launch, ONE host read of
            # the total to size the result (wall clock, synchronised)
            for _ in range(3):
                ops.expand(eb, children='compact')
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                ops.expand(eb, children='compact')
            torch.cuda.synchronize()
            t_call = (time.perf_counter() - t0) / 10
            bytes_c = n_e * (8 + 8 + 1 + 32 + 16 + 4 + 4 + 8) + 8 * n_children
            exp_out[str(n_e)] = {
                'children_per_board': n_children / n_e,
                'compact': {'us_per_launch': t_c * 1e6, 'boards_per_sec': n_e / t_c, 'children_per_sec': n_children / t_c,
                            'algorithmic_gbs': bytes_c / t_c / 1e9, 'frac_of_hbm_peak': bytes_c / t_c / 1e9 / peak,
                            'whole_call_us': t_call * 1e6},
                'dense': {'us_per_launch': t_d * 1e6, 'boards_per_sec': n_e / t_d},
            }
            del ex_c, ex_d
        exp_out['note'] = ('compact.us_per_launch: g2048_expand_compact into preallocated buffers — moves, device-side '
                           'scan of the child totals (offsets are an output) and children in one asynchronous call '
                           '(a single kernel below 2^17 boards), device time from a CUDA graph of back-to-back calls; '
                           'compact.whole_call_us: ops.expand(boards, children="compact") from scratch, wall clock incl. '
                           'allocation and the one host read that sizes the result; dense = [N,4,16,2] slots')
        extra['expand'] = exp_out
        # config 4 on ONE GPU at two batch sizes (the sharded run over all ranks is in
        # roofline.other_kernels.rollout_sharded)
        ro_out = {}
        for lg in (20, 21):
            n_r = 1 << lg
            ro = ops.rollout(n_r, seed=1, per_env=False)
            torch.cuda.synchronize()
            r_steps = int(ro.stats[1].item())
            _, ro_m, ro_b = timed(lambda: ops.rollout(n_r, seed=1, per_env=False), 3, 1, collective=False)
            ro_out[str(n_r)] = {'env_steps': r_steps, 'ms': ro_m * 1e3, 'env_steps_per_sec': r_steps / ro_m,
                                'best_env_steps_per_sec': r_steps / ro_b}
        extra['rollout_one_gpu'] = ro_out
        # config 5: self-play stepping loop, N = 4096, 200 steps, three API calls per iteration
        from simulate_2048_b200 import game
        st = game.batched_reset(7, 4096, device=dev)
        acts = torch.randint(0, 4, (200, 4096), dtype=torch.uint8, device=dev, generator=gen)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(200):
            game.batched_get_observation(st)
            game.batched_get_legal_actions(st)
            st, _, _, _ = game.batched_step(st, acts[t], 5, inplace=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        extra['selfplay_4096'] = {'us_per_iteration': dt / 200 * 1e6, 'env_steps_per_sec': 4096 * 200 / dt,
                                  'note': 'reference call pattern: observation + legal mask + batched_step per iteration'}
        # the same loop through BatchedEnv: one graph-replayed g2048_env_step launch per iteration
        from simulate_2048_b200.selfplay import BatchedEnv
        benv = BatchedEnv(4096, seed=7, device=dev, graph=True)
        benv.reset()
        for t in range(3):
            benv.step(acts[t])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(200):
            benv.step(acts[t])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        extra['selfplay_4096_fused'] = {'us_per_iteration': dt / 200 * 1e6, 'env_steps_per_sec': 4096 * 200 / dt,
                                        'note': 'BatchedEnv: next observation, legal mask, reward, done from ONE '
                                                'graph-replayed launch per iteration'}
        # the whole self-play iteration (sample from a policy, record, step) as ONE graph launch
        from simulate_2048_b200.selfplay import TrajectoryBuffer
        tbuf = TrajectoryBuffer(256, 4096, device=dev)
        benv.reset()
        pol = torch.rand((4096, 4), device=dev, generator=gen)
        val = torch.rand(4096, device=dev, generator=gen)
        benv.policy_in.copy_(pol)
        benv.values_in.copy_(val)
        for t in range(3):
            benv.step_policy(benv.policy_in, benv.values_in, record=tbuf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(200):
            benv.step_policy(benv.policy_in, benv.values_in, record=tbuf)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        extra['selfplay_4096_policy_recorded'] = {
            'us_per_iteration': dt / 200 * 1e6, 'env_steps_per_sec': 4096 * 200 / dt,
            'note': 'BatchedEnv.step_policy: g2048_selfplay_step (sample_action semantics, trajectory record with '
                    'policies / values, env step, next observation + legal mask) replayed from one CUDA graph'}
        # BASELINE configs[0] through the drop-in: 1024 random-policy games to game-over, lock-step
        from simulate_2048_b200.twentyfortyeight.envs import TwentyFortyEight, VectorTwentyFortyEight
        venv = VectorTwentyFortyEight(1024, seed=0)
        rng = np.random.default_rng(0)
        best = None
        for rep in range(3):
            venv.reset(seed=rep)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_steps = 0
            while True:
                fin = venv.is_finished
                if fin.all():
                    break
                mask = venv.legal_actions_mask
                acts = (rng.random((1024, 4)) * mask).argmax(axis=1)
                venv.step(acts)
                n_steps += int((~fin).sum())
            dt = time.perf_counter() - t0
            if best is None or n_steps / dt > best[0]:
                best = (n_steps / dt, n_steps, dt)
        single = TwentyFortyEight()
        single.reset(seed=0)
        t0 = time.perf_counter()
        k1 = 0
        while not single.is_finished and k1 < 200:
            single.step(int(rng.integers(4)))
            k1 += 1
        dt1 = time.perf_counter() - t0
        extra['config1_through_dropin'] = {
            'games': 1024, 'env_steps': best[1], 'seconds': best[2], 'env_steps_per_sec': best[0],
            'api': 'simulate_2048_b200.twentyfortyeight.envs.VectorTwentyFortyEight (lock-step games, host policy '
                   'over the returned legal masks, one env_step launch + pinned copies per step), best of 3',
            'single_env_class_steps_per_sec': k1 / dt1,
            'single_env_note': 'the one-game TwentyFortyEight mirror pays a host/device round trip per call; the '
                               'vector env is the fast form of the drop-in'}
        # row (f2): batched expectimax (expand + value function on every leaf + probability-weighted backup)
        from simulate_2048_b200 import selfplay

        def leaf_value(children):  # stand-in for a value network: empty cells + max exponent / 4
            o, mx = ops.observe(children)
            return (o == 0).sum(dim=1).to(torch.float32) + 0.25 * mx.to(torch.float32)
        xm = {}
        for depth, n_x in ((1, 65536), (2, 4096)):
            xb = boards[:n_x].contiguous()
            _, _, ex_x = selfplay.expectimax(xb, leaf_value, depth=depth)
            _, t_x, _ = timed(lambda: selfplay.expectimax(xb, leaf_value, depth=depth), 5, 2, collective=False)
            xm[f'depth{depth}_{n_x}'] = {'ms': t_x * 1e3, 'root_boards_per_sec': n_x / t_x,
                                        'first_level_children': int(ex_x.children.numel())}
            del ex_x
        wide_fn = lambda: selfplay.expectimax(boards[:4096].contiguous(), leaf_value, depth=2, widen=4)  # noqa: E731
        _, t_w, _ = timed(wide_fn, 5, 2, collective=False)
        xm['depth2_4096_widen4'] = {'ms': t_w * 1e3, 'root_boards_per_sec': 4096 / t_w,
                                    'note': 'progressive widening: 4 sampled chance children per afterstate and level'}
        xm['note'] = 'selfplay.expectimax: one g2048_expand per level, leaf values from a torch function, g2048_expect backup'
        extra['expectimax'] = xm
        # row (f4): n-step TD(lambda) targets over device trajectory buffers [T=200, N=4096]
        tr = torch.rand((200, 4096), device=dev) * 64
        tv = torch.rand((200, 4096), device=dev) * 500
        tout = torch.empty_like(tr)
        _, t_ns, _ = timed(lambda: ops.nstep_returns(tr, tv, 0.999, 10, 0.5, out=tout), 20, 3, collective=False)
        extra['nstep_returns_200x4096'] = {'us_per_call': t_ns * 1e6, 'targets_per_sec': 200 * 4096 / t_ns,
                                           'note': 'g2048_nstep_returns, n_steps 10, f64 arithmetic'}
        line['secondary'] = extra

    if rank == 0 and world == 1 and not args.no_cpu:
        line['cpu_baseline'] = cpu_baseline_block(base_np, actions[: base_np.size].cpu().numpy(), host_cores())

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)   # host-side wait (gloo): no kernel spins on the idle GPUs
        dist.destroy_process_group()
    return 0


def _quiet_stdout():
    """Route fd 1 to stderr (NCCL prints a version banner on stdout) and return a file object on
    the real stdout for the single JSON line."""
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


def main():
    global print
    real_stdout = _quiet_stdout()
    import builtins

    def print(*a, **kw):  # noqa: A001 - only the JSON line reaches the real stdout
        if kw.get('file') is None:
            kw['file'] = real_stdout
        builtins.print(*a, **kw)

    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--boards-log2', type=int, default=26)
    ap.add_argument('--e2e-log2', type=int, default=26,
                    help='boards per end-to-end step (default: the same 2^26 batch as the device-resident measurement)')
    ap.add_argument('--rollout-log2', type=int, default=24,
                    help='total envs of the sharded rollout (BASELINE configs[3]: 2^24 over all ranks)')
    ap.add_argument('--quick', action='store_true', help='skip the secondary measurements')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline block')
    ap.add_argument('--ncu-traffic', type=float, default=None,
                    help='dram bytes per launch from the committed ncu capture (profiles/)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    return run_reference(args) if args.impl == 'reference' else run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
