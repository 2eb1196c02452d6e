"""Stages the UNMODIFIED reference NumPy engine under ``oracle/_ref/`` (test infrastructure).

The reference (schalappe/simulate_2048) is pure Python + NumPy, so there is nothing to compile: the
"build" of ``oracle/_ref`` is a verbatim copy of its ``twentyfortyeight/`` package — the files the
hot path lives in — made where the reference is mounted (``/root/reference``, the
build container; ``__graft_entry__.build()`` calls this).  ``oracle/_ref/`` is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so the copy travels to the GPU box,
where ``/root/reference`` does not exist.  Used only by tests/, smoke() and bench.py's
``cpu_baseline`` / ``--impl reference`` legs; the product never imports it.

    python -m oracle.make_ref          # stage (or re-stage) and print the manifest
"""

from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get('G2048_REFERENCE', '/root/reference')
SRC = os.path.join(REF_ROOT, 'twentyfortyeight')
DST_ROOT = os.path.join(HERE, '_ref')
DST = os.path.join(DST_ROOT, 'twentyfortyeight')
MANIFEST = os.path.join(DST_ROOT, 'MANIFEST.json')
REFERENCE_TESTS = ('test_board.py', 'test_move.py', 'test_environment.py')   # the NumPy engine's own tests
TESTS_DST = os.path.join(DST_ROOT, 'reference_tests')


def _sha(path: str) -> str:
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(force: bool = False) -> Optional[str]:
    """Copy the reference package if it is mounted; return the staged path (None if neither the
    reference nor an earlier copy exists)."""
    if os.path.isdir(SRC) and (force or not os.path.exists(MANIFEST) or not os.path.isdir(TESTS_DST)):
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        # the reference's own tests of this path (run against the device-backed mirror by
        # tests/test_gpu_reference_suite.py)
        tests_dst = os.path.join(DST_ROOT, 'reference_tests')
        if os.path.isdir(tests_dst):
            shutil.rmtree(tests_dst)
        os.makedirs(tests_dst)
        for name in REFERENCE_TESTS:
            src = os.path.join(REF_ROOT, 'tests', name)
            if os.path.exists(src):
                shutil.copy(src, os.path.join(tests_dst, name))
        files = {}
        for root, _, names in list(os.walk(DST)) + list(os.walk(tests_dst)):
            for name in sorted(names):
                p = os.path.join(root, name)
                files[os.path.relpath(p, DST_ROOT)] = _sha(p)
        with open(MANIFEST, 'w') as f:
            json.dump({'source': SRC, 'files': files}, f, indent=1, sort_keys=True)
    return DST if os.path.exists(MANIFEST) else None


def available() -> bool:
    return stage() is not None


def load():
    """Import the staged reference: returns (gameboard, gamemove, TwentyFortyEight).  Raises
    ImportError when it was never staged."""
    if stage() is None:
        raise ImportError('oracle/_ref is not staged (run `python -m oracle.make_ref` where /root/reference is mounted)')
    if DST_ROOT not in sys.path:
        sys.path.insert(0, DST_ROOT)
    gameboard = importlib.import_module('twentyfortyeight.core.gameboard')
    gamemove = importlib.import_module('twentyfortyeight.core.gamemove')
    env = importlib.import_module('twentyfortyeight.envs.twentyfortyeight')
    if not os.path.abspath(gameboard.__file__).startswith(DST_ROOT):
        raise ImportError(f'twentyfortyeight resolved to {gameboard.__file__}, not to oracle/_ref')
    return gameboard, gamemove, env.TwentyFortyEight


if __name__ == '__main__':
    print(stage(force=True))
    if os.path.exists(MANIFEST):
        print(open(MANIFEST).read())
