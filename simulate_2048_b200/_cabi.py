"""ctypes binding of libg2048.so — the only way the Python layer reaches the GPU.

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  Signatures mirror include/g2048.h one to one.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libg2048.so')

OK, EINVAL, ECUDA, ENODEV, ERANGE = 0, -1, -2, -3, -4

FLAG_LEFT, FLAG_UP, FLAG_RIGHT, FLAG_DOWN = 1, 2, 4, 8
FLAG_LEGAL_MASK = 0x0F
FLAG_DONE, FLAG_MOVED, FLAG_OVERFLOW, FLAG_BADSPAWN = 0x10, 0x20, 0x40, 0x80
CHILDREN_NONE, CHILDREN_DENSE, CHILDREN_COMPACT = 0, 1, 2
IMPL_AUTO, IMPL_LUT, IMPL_REGISTER, IMPL_GATHER = 0, 1, 2, 3
STATS_WORDS = 24

# every symbol include/g2048.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
_U64 = C.c_uint64
_U32 = C.c_uint32
PROTOTYPES = {
    'g2048_abi_version': (C.c_int, []),
    'g2048_error_string': (C.c_char_p, [C.c_int]),
    'g2048_last_cuda_error': (C.c_char_p, []),
    'g2048_init': (C.c_int, [C.c_int]),
    'g2048_shutdown': (C.c_int, []),
    'g2048_set_impl': (C.c_int, [C.c_int]),
    'g2048_set_tuning': (C.c_int, [C.c_int, C.c_int]),
    'g2048_step': (C.c_int, [_P, _P, _P, _U64, _U64, _P, _U32, _P, _P, _P, _P, _P, _SZ, _P]),
    'g2048_env_step': (C.c_int, [_P, _P, _P, _P, _P, _P, _U64, _U64, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'g2048_sample_action': (C.c_int, [_P, _P, C.c_float, _U64, _U64, _P, _U32, _P, _P, _SZ, _P]),
    'g2048_selfplay_step': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_float, _U64, _U64, _P, _U32,
                                      _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'g2048_mask': (C.c_int, [_P, _P, _SZ, _P]),
    'g2048_legal_actions': (C.c_int, [_P, _P, _P, _SZ, _P]),
    'g2048_expand': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _SZ, _P]),
    'g2048_expand_children': (C.c_int, [_P, _P, _P, _P, C.c_int, _SZ, _P]),
    'g2048_expand_compact': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P, _SZ, _P]),
    'g2048_after_state': (C.c_int, [_P, _P, _P, _P, _SZ, _P]),
    'g2048_reset': (C.c_int, [_U64, _U64, _P, _P, _P, _SZ, _P]),
    'g2048_spawn': (C.c_int, [_P, _P, _U64, _U64, _P, _U32, _U32, _P, _SZ, _P]),
    'g2048_observe': (C.c_int, [_P, _P, _P, _SZ, _P]),
    'g2048_observe_onehot': (C.c_int, [_P, _P, C.c_int, _SZ, _P]),
    'g2048_expect': (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_float, _P, _SZ, _P]),
    'g2048_sample_children': (C.c_int, [_P, _P, _U32, _U64, _U64, _U32, _P, _SZ, _P]),
    'g2048_expect_sampled': (C.c_int, [_P, _P, _P, _U32, C.c_float, C.c_float, _P, _SZ, _P]),
    'g2048_nstep_returns': (C.c_int, [_P, _P, _P, C.c_double, C.c_int, C.c_double, _P, _SZ, _SZ, _P]),
    'g2048_rollout': (C.c_int, [_U64, _U64, _SZ, _U32, C.c_int, _P, _P, _P, _P, _P, _P, _U32, _P]),
    'g2048_pack': (C.c_int, [_P, _P, _P, _SZ, _P]),
    'g2048_unpack': (C.c_int, [_P, _P, _SZ, _P]),
    'g2048_row_table_host': (C.c_int, [_P, _P, _P]),
    'g2048_step_host': (C.c_int, [_P, _P, _P, _U64, _U64, _U32, _P, _P, _P, _SZ]),
    'g2048_step_host_copies': (C.c_int, [_P, _P, _P, _P, _P, _P, _SZ]),
    'g2048_host_alloc': (C.c_int, [C.POINTER(C.c_void_p), _SZ]),
    'g2048_host_numa_info': (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'g2048_host_free': (C.c_int, [_P]),
}


class G2048Error(RuntimeError):
    """A g2048_* call returned a negative code."""

    def __init__(self, code: int, text: str):
        super().__init__(f'g2048 error {code}: {text}')
        self.code = code


_lib = None


def lib() -> C.CDLL:
    """Load libg2048.so (raises if it has not been built — there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -m simulate_2048_b200.build` '
                '(nvcc, sm_100a). simulate_2048_b200 has no CPU fallback.'
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.g2048_abi_version() != 3:
            raise ImportError('libg2048.so ABI version mismatch; rebuild it')
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != OK:
        h = lib()
        text = h.g2048_error_string(code).decode()
        if code in (ECUDA, ENODEV):
            text += ' — ' + h.g2048_last_cuda_error().decode()
        raise G2048Error(code, text)
