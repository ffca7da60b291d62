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
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("mira_oracle.c", "mira_oracle_witness.c", "oracle_field.h", "mira_oracle.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
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
    L.oracle_fold_w.argtypes = [i, vp, vp, sz, vp, vp]
    L.oracle_fold_e.argtypes = [i, vp, vp, sz, sz, vp, vp]
    L.oracle_concat_pad.argtypes = [vp, vp, sz, sz, vp]
    L.oracle_concat_pad.restype = sz
    L.oracle_eval_rows.argtypes = [i, vp, sz, vp, sz, vp, sz, C.c_uint32, vp, sz, sz, vp]
    L.oracle_eval_rows.restype = i
    L.oracle_lookup_m.argtypes = [i, vp, sz, vp, sz, vp]
    L.oracle_lookup_h_g.argtypes = [i, vp, vp, vp, sz, vp, vp, vp]
    L.oracle_fft.argtypes = [i, vp, C.c_uint32, vp]
    L.oracle_fft_omega.argtypes = [i, C.c_uint32, i, vp]
    L.oracle_fft_omega.restype = i
    L.oracle_fft_divisor.argtypes = [i, C.c_uint32, vp]
    L.oracle_fft_forward.argtypes = [i, vp, C.c_uint32]
    L.oracle_fft_forward.restype = i
    L.oracle_fft_inverse.argtypes = [i, vp, C.c_uint32]
    L.oracle_fft_inverse.restype = i
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


# ---------------------------------------------------------------------------------- witness side
def fold_w(field, w1: bytes, w2: bytes, r: bytes) -> bytes:
    o = _buf(max(len(w1), 1)); lib().oracle_fold_w(field, w1, w2, len(w1) // 32, r, o); return o.raw[:len(w1)]


def _ptr_array(bufs):
    keep = [C.create_string_buffer(bytes(b), max(len(b), 1)) for b in bufs]
    arr = (C.c_void_p * max(len(bufs), 1))(*[C.cast(k, C.c_void_p).value for k in keep])
    return arr, keep


def fold_e(field, e: bytes, terms, r: bytes) -> bytes:
    arr, keep = _ptr_array(terms)
    o = _buf(max(len(e), 1)); lib().oracle_fold_e(field, e, arr, len(terms), len(e) // 32, r, o); return o.raw[:len(e)]


def concat_pad(cols, pad_size: int) -> bytes:
    arr, keep = _ptr_array(cols)
    lens = (C.c_size_t * max(len(cols), 1))(*[len(c) // 32 for c in cols])
    n = lib().oracle_concat_pad(arr, lens, len(cols), pad_size, None)
    o = _buf(max(32 * n, 1)); lib().oracle_concat_pad(arr, lens, len(cols), pad_size, o); return o.raw[:32 * n]


class EvalDomainStruct(C.Structure):
    """oracle_eval_domain == mira_eval_domain (same field order)."""
    _fields_ = [("row_size", C.c_uint64), ("num_selectors", C.c_uint32), ("num_fixed", C.c_uint32),
                ("num_advice", C.c_uint32), ("num_lookup", C.c_uint32), ("num_challenges", C.c_uint32),
                ("num_w1", C.c_uint32), ("num_w2", C.c_uint32), ("flags", C.c_uint32),
                ("selectors", C.c_void_p), ("fixed", C.c_void_p), ("w1", C.c_void_p), ("w1_len", C.c_void_p),
                ("w2", C.c_void_p), ("w2_len", C.c_void_p), ("challenges", C.c_void_p)]


class EvalError(Exception):
    def __init__(self, rc):
        super().__init__(f"eval error {rc}")
        self.rc = rc


def eval_rows(field, prog: dict, dom: dict, row_begin=0, row_end=None) -> bytes:
    """prog: GraphEvaluator.encode() with `constants` as Montgomery bytes; dom: dict with row_size,
    selectors [bytes of 0/1], fixed [bytes], w1 [bytes], w2 [bytes], challenges bytes, num_advice, num_lookup."""
    row_end = dom["row_size"] if row_end is None else row_end
    code = (C.c_uint32 * max(len(prog["code"]), 1))(*prog["code"])
    rots = (C.c_int32 * max(len(prog["rotations"]), 1))(*prog["rotations"])
    consts = prog["constants"]
    sel, k1 = _ptr_array(dom.get("selectors", []))
    fx, k2 = _ptr_array(dom.get("fixed", []))
    w1, k3 = _ptr_array(dom.get("w1", []))
    w2, k4 = _ptr_array(dom.get("w2", []))
    l1 = (C.c_uint64 * max(len(dom.get("w1", [])), 1))(*[len(b) // 32 for b in dom.get("w1", [])])
    l2 = (C.c_uint64 * max(len(dom.get("w2", [])), 1))(*[len(b) // 32 for b in dom.get("w2", [])])
    ch = dom.get("challenges", b"")
    chb = C.create_string_buffer(ch, max(len(ch), 1))
    d = EvalDomainStruct(dom["row_size"], len(dom.get("selectors", [])), len(dom.get("fixed", [])), dom.get("num_advice", 0),
                         dom.get("num_lookup", 0), len(ch) // 32, len(dom.get("w1", [])), len(dom.get("w2", [])), dom.get("flags", 0),
                         C.cast(sel, C.c_void_p), C.cast(fx, C.c_void_p), C.cast(w1, C.c_void_p), C.cast(l1, C.c_void_p),
                         C.cast(w2, C.c_void_p), C.cast(l2, C.c_void_p), C.cast(chb, C.c_void_p))
    n = max(row_end - row_begin, 0)
    o = _buf(max(32 * n, 1))
    rc = lib().oracle_eval_rows(field, code, len(prog["code"]), consts, len(consts) // 32, rots, len(prog["rotations"]),
                                prog["num_intermediates"], C.byref(d), row_begin, row_end, o)
    if rc:
        raise EvalError(rc)
    return o.raw[:32 * n]


def fft_omega(field, k, inverse=False) -> bytes:
    o = _buf(32)
    if lib().oracle_fft_omega(field, k, 1 if inverse else 0, o):
        raise ValueError("k exceeds the field's two-adicity")
    return o.raw


def best_fft(field, a: bytes, log_n: int, omega: bytes) -> bytes:
    o = C.create_string_buffer(a, len(a)); lib().oracle_fft(field, o, log_n, omega); return o.raw[:len(a)]


def fft(field, a: bytes, log_n: int) -> bytes:
    o = C.create_string_buffer(a, len(a)); assert lib().oracle_fft_forward(field, o, log_n) == 0; return o.raw[:len(a)]


def ifft(field, a: bytes, log_n: int) -> bytes:
    o = C.create_string_buffer(a, len(a)); assert lib().oracle_fft_inverse(field, o, log_n) == 0; return o.raw[:len(a)]


def lookup_m(field, l: bytes, t: bytes) -> bytes:
    o = _buf(max(len(t), 1)); lib().oracle_lookup_m(field, l, len(l) // 32, t, len(t) // 32, o); return o.raw[:len(t)]


def lookup_h_g(field, l: bytes, t: bytes, m: bytes, r: bytes):
    h, g = _buf(max(len(l), 1)), _buf(max(len(l), 1))
    lib().oracle_lookup_h_g(field, l, t, m, len(l) // 32, r, h, g)
    return h.raw[:len(l)], g.raw[:len(l)]
