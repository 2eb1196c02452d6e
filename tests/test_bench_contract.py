"""The bench.py output contract, checked on the committed lines of the last GPU runs
(profiles/r02_bench_final.json, r02_bench_reference_arm.json, r02_bench_2gpu.json and the round-1
8-GPU line): every key the driver reads is present and consistent.  No GPU needed."""

import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.load(f)


BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
             'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches'}


@pytest.mark.parametrize('name,n', [('r02_bench_final.json', 1), ('r02_bench_2gpu.json', 2), ('r02_bench_8gpu.json', 8),
                                    ('r01_bench_8gpu.json', 8)])
def test_our_arm_line(name, n):
    d = _load(name)
    assert BASE_KEYS <= set(d) and {'roofline', 'clocks'} <= set(d)
    assert d['metric'] == 'env_steps_per_sec' and d['unit'] == 'env-steps/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == n and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['warmup'] >= 3 and d['gpu_launches'] == d['steps'] > 0
    assert 'workload' in d['config'] and 'l2' in d['config'] and 'model' not in d['config']
    # value = envs of all ranks x steps / time
    per_step = d['config']['boards_per_gpu'] * n
    assert abs(d['value'] - per_step / (d['ms_per_step'] * 1e-3)) / d['value'] < 1e-6
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['traffic'] is None or r['traffic'] < 1.1 * r['algorithmic_bytes_per_launch']
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert e['value'] < d['value']            # host buffers: PCIe-bound, far below the device-resident rate
    c = d['clocks']
    assert not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    assert c['sm_mhz'] >= 0.9 * c['sm_max_mhz']
    if n == 1:
        b = d['cpu_baseline']
        assert b['kind'] in ('reference', 'port') and b['cores'] >= 1 and b['value'] > 0 and 'sample' in b


@pytest.mark.parametrize('name,n', [('r02_bench_final.json', 1), ('r02_bench_2gpu.json', 2), ('r02_bench_8gpu.json', 8)])
def test_round2_keys(name, n):
    """Philox-mode step and the sharded config-4 rollout live inside `roofline` (a key the driver's
    parse keeps); e2e carries the copy ceiling; the CPU baseline is the real reference engine."""
    d = _load(name)
    other = d['roofline']['other_kernels']
    p = other['philox_step']
    assert p['bytes_per_env_step'] == 22 and 0.5 < p['frac'] < 1 and p['value'] > 1.5e11 * n
    ro = other['rollout_sharded']
    assert ro['envs_total'] == 1 << 24 and ro['envs_per_gpu'] * n == ro['envs_total'] and ro['scaling'] == 'strong'
    assert ro['stats_equal_oracle'] is True                                # all 24 words equal the CPU oracle's
    assert ro['env_steps'] == 1986340051                                   # tests/golden/rollout_stats.json [1]
    assert abs(ro['env_steps_per_sec'] - ro['env_steps'] / (ro['ms'] * 1e-3)) / ro['env_steps_per_sec'] < 1e-6
    assert ('all_gather' in ro['collective']) == (n > 1)
    assert ro['summary']['episodes'] == 1 << 24 and 110 < ro['summary']['mean_length'] < 127
    e = d['e2e']
    assert e['copy_ceiling'] > 0 and 0.8 < e['frac_of_copy_ceiling'] < 1.1
    if n == 1:
        assert d['cpu_baseline']['kind'] == 'reference'
        assert d['cpu_baseline']['config1_numpy_engine_1024_games']['kind'] == 'reference'
        x = d['secondary']['expand']['65536']['compact']
        assert x['whole_call_us'] > x['us_per_launch'] > 0                  # the whole public call is timed too


def test_reference_arm_line():
    d = _load('r02_bench_reference_arm.json')
    assert d['impl'] == 'reference' and BASE_KEYS <= set(d)
    assert d['metric'] == 'env_steps_per_sec' and d['unit'] == 'env-steps/s'
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    b = d['cpu_baseline']
    assert b['kind'] == 'reference' and b['value'] == d['value'] and b['cores'] >= 1 and 'sample' in b
    assert d['gpu_launches'] == 0
