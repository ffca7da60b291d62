"""ctypes binding of the C ABI in include/mira_b200.h (libmira_b200.so, built in-tree by
__graft_entry__.build()).  There is no Python/CPU fallback: if the shared library is missing or
no sm_100 device is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MIRA_B200_LIB selects an A/B build of the same sources (mira_b200/csrc/Makefile VARIANT=...); never a fallback
LIB_PATH = os.environ.get("MIRA_B200_LIB") or os.path.join(_HERE, "libmira_b200.so")

MIRA_OK = 0
MIRA_ERR_TOO_LONG_INPUT = -1
MIRA_ERR_CUDA = -2
MIRA_ERR_INVALID = -3
MIRA_ERR_NOT_ON_CURVE = -4
MIRA_ERR_EVAL_CHALLENGE = -11
MIRA_ERR_EVAL_COLUMN = -12
MIRA_ERR_EVAL_ROW = -13
MIRA_ERR_EVAL_WITNESS_INDEX = -14
MIRA_ERR_EVAL_PROGRAM = -15
MIRA_FQ, MIRA_FR = 0, 1
MIRA_EVAL_LOOKUP_DOMAIN = 1

# every symbol include/mira_b200.h declares (tests/test_capi_symbols.py checks the .so exports them all)
SYMBOLS = [
    "mira_last_error", "mira_msm_ctx_create", "mira_msm_ctx_create_sharded", "mira_msm_ctx_num_devices", "mira_msm_ctx_destroy", "mira_msm_ctx_len",
    "mira_msm_ctx_check_on_curve", "mira_msm_ctx_prepare", "mira_msm_ctx_prepare_for", "mira_msm_commit", "mira_msm_commit_device", "mira_msm_commit_batch", "mira_msm_scalars_device",
    "mira_msm_partial", "mira_msm_combine", "mira_msm_get_stats", "mira_msm_set_profiling",
    "mira_msm_set_window", "mira_msm_set_adaptive_window", "mira_msm_set_slice_min", "mira_msm_set_affine_levels", "mira_msm_set_pipeline", "mira_msm_partial_batch_dev", "mira_msm_combine_dev", "mira_dev_alloc", "mira_dev_free", "mira_dev_upload", "mira_dev_download", "mira_dev_sync", "mira_host_register", "mira_host_unregister", "mira_gen_scalars", "mira_gen_bases", "mira_test_field_op", "mira_test_point_op",
    "mira_fold_w", "mira_fold_e", "mira_concat_pad", "mira_eval_program_create", "mira_eval_program_destroy",
    "mira_eval_rows", "mira_eval_rows_range", "mira_eval_program_stats", "mira_lookup_m", "mira_lookup_h_g", "mira_fft", "mira_fft_std", "mira_test_eval_link_multi", "mira_eval_rows_multi",
]
_VOID = ("mira_last_error", "mira_msm_ctx_destroy", "mira_msm_ctx_len", "mira_msm_ctx_num_devices", "mira_eval_program_destroy", "mira_msm_scalars_device")


class EvalDomain(C.Structure):
    """mira_eval_domain (include/mira_b200.h) == PlonkEvalDomain (src/plonk/eval.rs:93-106)."""
    _fields_ = [("row_size", C.c_uint64), ("num_selectors", C.c_uint32), ("num_fixed", C.c_uint32),
                ("num_advice", C.c_uint32), ("num_lookup", C.c_uint32), ("num_challenges", C.c_uint32),
                ("num_w1", C.c_uint32), ("num_w2", C.c_uint32), ("flags", C.c_uint32),
                ("selectors", C.c_void_p), ("fixed", C.c_void_p), ("w1", C.c_void_p), ("w1_len", C.c_void_p),
                ("w2", C.c_void_p), ("w2_len", C.c_void_p), ("challenges", C.c_void_p)]


class EvalStats(C.Structure):
    _fields_ = [("instructions", C.c_uint32), ("slots", C.c_uint32), ("accesses", C.c_uint32), ("uniforms", C.c_uint32),
                ("muls", C.c_uint32), ("adds", C.c_uint32), ("loads", C.c_uint32), ("fused", C.c_uint32)]



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
    L.mira_msm_ctx_create_sharded.argtypes = [i, vp, sz, C.POINTER(C.c_int), sz, C.POINTER(vp)]
    L.mira_msm_ctx_num_devices.argtypes = [vp]
    L.mira_msm_ctx_num_devices.restype = sz
    L.mira_msm_ctx_destroy.argtypes = [vp]
    L.mira_msm_ctx_destroy.restype = None
    L.mira_msm_ctx_len.argtypes = [vp]
    L.mira_msm_ctx_len.restype = sz
    L.mira_msm_ctx_check_on_curve.argtypes = [vp]
    L.mira_msm_ctx_prepare.argtypes = [vp, sz]
    L.mira_msm_ctx_prepare_for.argtypes = [vp, vp, sz, i]
    L.mira_msm_commit.argtypes = [vp, vp, sz, vp]
    L.mira_msm_commit_device.argtypes = [vp, vp, sz, vp, vp]
    L.mira_msm_scalars_device.argtypes = [vp, C.POINTER(sz)]
    L.mira_msm_scalars_device.restype = vp
    L.mira_msm_commit_batch.argtypes = [vp, vp, sz, sz, vp, vp]
    L.mira_msm_partial.argtypes = [vp, vp, sz, i, vp, vp]
    L.mira_msm_combine.argtypes = [i, vp, sz, i, vp]
    L.mira_msm_get_stats.argtypes = [vp, C.POINTER(MsmStats)]
    L.mira_msm_set_profiling.argtypes = [vp, i]
    L.mira_msm_set_window.argtypes = [vp, i]
    L.mira_msm_set_slice_min.argtypes = [vp, sz]
    L.mira_msm_set_affine_levels.argtypes = [vp, i]
    L.mira_msm_set_pipeline.argtypes = [vp, i, sz]
    L.mira_msm_partial_batch_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.mira_msm_combine_dev.argtypes = [i, vp, sz, sz, sz, i, vp, vp]
    L.mira_msm_set_adaptive_window.argtypes = [vp, i]
    L.mira_dev_alloc.argtypes = [i, sz, C.POINTER(vp)]
    L.mira_dev_free.argtypes = [i, vp]
    L.mira_dev_upload.argtypes = [i, vp, vp, sz, vp]
    L.mira_dev_download.argtypes = [i, vp, vp, sz, vp]
    L.mira_dev_sync.argtypes = [i, vp]
    L.mira_host_register.argtypes = [vp, sz]
    L.mira_host_unregister.argtypes = [vp]
    L.mira_gen_scalars.argtypes = [i, u64, sz, sz, i, i, vp]
    L.mira_gen_bases.argtypes = [i, u64, sz, sz, i, vp]
    L.mira_test_field_op.argtypes = [i, i, vp, vp, sz, i, vp]
    L.mira_test_point_op.argtypes = [i, i, vp, vp, sz, i, vp]
    L.mira_fold_w.argtypes = [i, vp, vp, sz, vp, vp, i, vp]
    L.mira_fold_e.argtypes = [i, vp, vp, sz, sz, vp, vp, i, vp]
    L.mira_concat_pad.argtypes = [vp, vp, sz, sz, vp, sz, C.POINTER(sz), i, vp]
    L.mira_eval_program_create.argtypes = [i, vp, sz, vp, sz, vp, sz, C.c_uint32, C.POINTER(vp)]
    L.mira_eval_program_destroy.argtypes = [vp]
    L.mira_eval_program_destroy.restype = None
    L.mira_eval_rows.argtypes = [vp, C.POINTER(EvalDomain), vp, i, vp]
    L.mira_eval_rows_range.argtypes = [vp, C.POINTER(EvalDomain), C.c_uint64, C.c_uint64, vp, i, vp]
    L.mira_eval_program_stats.argtypes = [vp, C.POINTER(EvalStats)]
    u32p = C.POINTER(C.c_uint32)
    L.mira_test_eval_link_multi.argtypes = [vp, sz, C.POINTER(EvalDomain), vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz), vp, sz,
                                            C.POINTER(sz), u32p]
    L.mira_eval_rows_multi.argtypes = [vp, sz, C.POINTER(EvalDomain), C.c_uint64, C.c_uint64, vp, i, vp]
    L.mira_lookup_m.argtypes = [i, vp, sz, vp, sz, vp, i, vp]
    L.mira_lookup_h_g.argtypes = [i, vp, vp, vp, sz, vp, vp, vp, i, vp]
    L.mira_fft.argtypes = [i, vp, C.c_uint32, vp, i, vp]
    L.mira_fft_std.argtypes = [i, vp, C.c_uint32, i, i, vp]
    for name in SYMBOLS:
        if name not in _VOID:
            getattr(L, name).restype = i
    _LIB = L
    return L


def last_error() -> str:
    return lib().mira_last_error().decode()
