"""Builds libg2048.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The shared library lives next to this file so that it travels to the GPU box with the
repository snapshot (it is git-ignored, not gpurun-ignored).
"""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libg2048.so')
SOURCES = ['g2048_kernels.cu']
HEADERS = ['g2048_device.cuh', 'g2048_fast.cuh', os.path.join('..', '..', 'include', 'g2048.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-shared',
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources are newer than the library.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libg2048.so')
    cmd = [nvcc, *NVCC_FLAGS, '-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
