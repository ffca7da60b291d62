"""Helpers shared by the GPU parity tests (torch is used only for device memory)."""
import ctypes as C

import torch

from mira_b200 import _native as N
from mira_b200.commitment import _check


def gen_scalars_dev(curve, seed, n, dist=0, first=0, device=0):
    buf = torch.empty(max(n, 1) * 32, dtype=torch.uint8, device=f"cuda:{device}")
    _check(N.lib().mira_gen_scalars(curve, seed, first, n, dist, device, buf.data_ptr()))
    return buf[: n * 32]


def gen_bases_dev(curve, seed, n, first=0, device=0):
    buf = torch.empty(max(n, 1) * 64, dtype=torch.uint8, device=f"cuda:{device}")
    _check(N.lib().mira_gen_bases(curve, seed, first, n, device, buf.data_ptr()))
    return buf[: n * 64]


def to_bytes(t):
    return t.cpu().numpy().tobytes()


def field_op(field, op, a: bytes, b: bytes = None, device=0) -> bytes:
    n = len(a) // 32
    out = C.create_string_buffer(len(a))
    _check(N.lib().mira_test_field_op(field, op, a, b, n, device, out))
    return out.raw


def point_op(curve, op, p: bytes, q: bytes = None, device=0) -> bytes:
    n = len(p) // 64
    out = C.create_string_buffer(len(p))
    _check(N.lib().mira_test_point_op(curve, op, p, q, n, device, out))
    return out.raw
