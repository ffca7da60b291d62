"""ctypes binding of the CPU oracle (oracle/libmira_oracle.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ORACLE_DIR, "libmira_oracle.so")
    src = os.path.join(ORACLE_DIR, "mira_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    L = C.CDLL(so)
    vp, sz, i, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.oracle_fe_from_u64.argtypes = [i, u64, vp]
    for name in ("oracle_fe_from_canonical", "oracle_fe_to_canonical", "oracle_fe_inv"):
        getattr(L, name).argtypes = [i, vp, vp]
    for name in ("oracle_fe_add", "oracle_fe_sub", "oracle_fe_mul"):
        getattr(L, name).argtypes = [i, vp, vp, vp]
    L.oracle_fe_mul_many.argtypes = [i, vp, vp, sz, vp]
    L.oracle_generator.argtypes = [i, vp]
    L.oracle_is_on_curve.argtypes = [i, vp]
    L.oracle_is_on_curve.restype = i
    L.oracle_point_add_affine.argtypes = [i, vp, vp, vp]
    L.oracle_point_neg_affine.argtypes = [i, vp, vp]
    L.oracle_scalar_mul.argtypes = [i, vp, vp, vp]
    L.oracle_commit.argtypes = [i, vp, sz, vp, sz, i, vp]
    L.oracle_commit.restype = i
    L.oracle_commit_naive.argtypes = [i, vp, vp, sz, vp]
    L.oracle_gen_scalars.argtypes = [i, u64, sz, sz, i, vp]
    L.oracle_gen_bases.argtypes = [i, u64, sz, sz, i, vp]
    L.oracle_num_cores.restype = i
    _LIB = L
    return L


def _buf(n):
    return C.create_string_buffer(n)


def fe_from_u64(field, v):
    o = _buf(32); lib().oracle_fe_from_u64(field, v, o); return o.raw


def fe_from_canonical(field, b):
    o = _buf(32); lib().oracle_fe_from_canonical(field, b, o); return o.raw


def fe_to_canonical(field, b):
    o = _buf(32); lib().oracle_fe_to_canonical(field, b, o); return o.raw


def fe_op(name, field, a, b):
    o = _buf(32); getattr(lib(), "oracle_fe_" + name)(field, a, b, o); return o.raw


def fe_inv(field, a):
    o = _buf(32); lib().oracle_fe_inv(field, a, o); return o.raw


def fe_mul_many(field, a: bytes, b: bytes) -> bytes:
    n = len(a) // 32
    o = _buf(len(a)); lib().oracle_fe_mul_many(field, a, b, n, o); return o.raw


def generator(curve):
    o = _buf(64); lib().oracle_generator(curve, o); return o.raw


def is_on_curve(curve, p):
    return bool(lib().oracle_is_on_curve(curve, p))


def point_add(curve, p, q):
    o = _buf(64); lib().oracle_point_add_affine(curve, p, q, o); return o.raw


def point_neg(curve, p):
    o = _buf(64); lib().oracle_point_neg_affine(curve, p, o); return o.raw


def scalar_mul(curve, base, scalar):
    o = _buf(64); lib().oracle_scalar_mul(curve, base, scalar, o); return o.raw


class TooLongInput(Exception):
    pass


def commit(curve, bases: bytes, scalars: bytes, threads=0) -> bytes:
    o = _buf(64)
    rc = lib().oracle_commit(curve, bases, len(bases) // 64, scalars, len(scalars) // 32, threads, o)
    if rc != 0:
        raise TooLongInput(f"input len {len(scalars)//32}, limit {len(bases)//64}")
    return o.raw


def commit_naive(curve, bases: bytes, scalars: bytes) -> bytes:
    o = _buf(64); lib().oracle_commit_naive(curve, bases, scalars, len(scalars) // 32, o); return o.raw


def gen_scalars(curve, seed, n, dist=0, first=0) -> bytes:
    o = _buf(32 * max(n, 1)); lib().oracle_gen_scalars(curve, seed, first, n, dist, o); return o.raw[:32 * n]


def gen_bases(curve, seed, n, first=0, threads=0) -> bytes:
    o = _buf(64 * max(n, 1)); lib().oracle_gen_bases(curve, seed, first, n, threads, o); return o.raw[:64 * n]


def num_cores():
    return lib().oracle_num_cores()
