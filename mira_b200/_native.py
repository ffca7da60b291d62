"""ctypes binding of the C ABI in include/mira_b200.h (libmira_b200.so, built in-tree by
__graft_entry__.build()).  There is no Python/CPU fallback: if the shared library is missing or
no sm_100 device is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmira_b200.so")

MIRA_OK = 0
MIRA_ERR_TOO_LONG_INPUT = -1
MIRA_ERR_CUDA = -2
MIRA_ERR_INVALID = -3
MIRA_ERR_NOT_ON_CURVE = -4

# every symbol include/mira_b200.h declares (tests/test_capi_symbols.py checks the .so exports them all)
SYMBOLS = [
    "mira_last_error", "mira_msm_ctx_create", "mira_msm_ctx_destroy", "mira_msm_ctx_len",
    "mira_msm_ctx_check_on_curve", "mira_msm_ctx_prepare", "mira_msm_commit", "mira_msm_commit_device",
    "mira_msm_partial", "mira_msm_combine", "mira_msm_get_stats", "mira_msm_set_profiling",
    "mira_msm_set_window", "mira_gen_scalars", "mira_gen_bases", "mira_test_field_op", "mira_test_point_op",
]


class MsmStats(C.Structure):
    _fields_ = [
        ("window_bits", C.c_int), ("windows", C.c_int), ("entries", C.c_uint64), ("buckets", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("ms_digits", C.c_float), ("ms_sort", C.c_float),
        ("ms_accumulate", C.c_float), ("ms_reduce", C.c_float), ("ms_total", C.c_float),
    ]


_LIB = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    """Load libmira_b200.so.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). mira_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, i, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.mira_last_error.restype = C.c_char_p
    L.mira_msm_ctx_create.argtypes = [i, vp, sz, i, i, C.POINTER(vp)]
    L.mira_msm_ctx_destroy.argtypes = [vp]
    L.mira_msm_ctx_destroy.restype = None
    L.mira_msm_ctx_len.argtypes = [vp]
    L.mira_msm_ctx_len.restype = sz
    L.mira_msm_ctx_check_on_curve.argtypes = [vp]
    L.mira_msm_ctx_prepare.argtypes = [vp, sz]
    L.mira_msm_commit.argtypes = [vp, vp, sz, vp]
    L.mira_msm_commit_device.argtypes = [vp, vp, sz, vp, vp]
    L.mira_msm_partial.argtypes = [vp, vp, sz, i, vp, vp]
    L.mira_msm_combine.argtypes = [i, vp, sz, i, vp]
    L.mira_msm_get_stats.argtypes = [vp, C.POINTER(MsmStats)]
    L.mira_msm_set_profiling.argtypes = [vp, i]
    L.mira_msm_set_window.argtypes = [vp, i]
    L.mira_gen_scalars.argtypes = [i, u64, sz, sz, i, i, vp]
    L.mira_gen_bases.argtypes = [i, u64, sz, sz, i, vp]
    L.mira_test_field_op.argtypes = [i, i, vp, vp, sz, i, vp]
    L.mira_test_point_op.argtypes = [i, i, vp, vp, sz, i, vp]
    for name in SYMBOLS:
        if name not in ("mira_last_error", "mira_msm_ctx_destroy", "mira_msm_ctx_len"):
            getattr(L, name).restype = i
    _LIB = L
    return L


def last_error() -> str:
    return lib().mira_last_error().decode()
