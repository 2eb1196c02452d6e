"""CPU-side checks of the drop-in boundary: libg2048.so loads without a GPU, exports every
symbol include/g2048.h declares, and refuses to compute without a device (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cabi():
    from simulate_2048_b200 import _cabi, build
    build.build()
    return _cabi


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'g2048.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(g2048_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(cabi):
    names = declared_symbols()
    assert len(names) >= 18
    handle = ctypes.CDLL(cabi.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f'{name} declared in include/g2048.h but not exported'
    assert sorted(cabi.PROTOTYPES) == names, 'python prototypes out of sync with the header'


def test_abi_version_and_error_strings(cabi):
    lib = cabi.lib()
    assert lib.g2048_abi_version() == 3  # v3: quad-block Philox layout + per-env game stream (reset, rollout)
    for code in (0, -1, -2, -3, -4):
        assert lib.g2048_error_string(code)
    assert lib.g2048_set_impl(99) == cabi.EINVAL


def test_flag_constants_match_header(cabi):
    text = open(os.path.join(ROOT, 'include', 'g2048.h')).read()
    want = {m[0]: int(m[1], 16) for m in re.findall(r'#define G2048_FLAG_([A-Z_]+) (0x[0-9A-F]+)u', text)}
    assert want['DONE'] == cabi.FLAG_DONE and want['MOVED'] == cabi.FLAG_MOVED
    assert want['OVERFLOW'] == cabi.FLAG_OVERFLOW and want['BADSPAWN'] == cabi.FLAG_BADSPAWN
    assert want['LEGAL_MASK'] == cabi.FLAG_LEGAL_MASK


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, 'simulate_2048_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('the oracle', ''), f'{f} mentions the oracle'


def test_no_compute_without_gpu(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from simulate_2048_b200 import ops
    with pytest.raises(ValueError):
        ops.step(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8))
    with pytest.raises(cabi.G2048Error):
        ops.step_host(np.zeros(4, np.uint64), np.zeros(4, np.uint8))


def test_missing_library_fails_loudly(cabi, monkeypatch):
    monkeypatch.setattr(cabi, '_lib', None)
    monkeypatch.setattr(cabi, 'LIB_PATH', '/nonexistent/libg2048.so')
    with pytest.raises(ImportError):
        cabi.lib()
