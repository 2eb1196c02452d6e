"""CPU check of the branch-free fast transition (simulate_2048_b200/csrc/g2048_fast.cuh): the
header is host-compilable, so the exact dataflow the step kernel runs (orientation algebra, row
hash, table-entry format, prefix-count spawn, legal mask) is compared with the oracle on a few
hundred thousand random boards, injected and Philox mode, without a GPU."""

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_transition_matches_oracle_on_host(oracle, tmp_path):
    exe = str(tmp_path / 'fast_host_check')
    so = os.path.join(ROOT, 'oracle', 'libg2048_oracle.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(ROOT, 'tests', 'fast_host_check.cpp'),
                           so, '-Wl,-rpath,' + os.path.dirname(so)])
    out = subprocess.run([exe, '400000'], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'mismatches: 0' in out.stdout
