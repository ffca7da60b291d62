"""Host-side mirror of the reference's witness-side hot path over the C ABI: cross-term row evaluation,
witness folding, column concatenation and FFT, all on vectors that live in GPU memory.

Mirrors (names, argument meaning, error behaviour):
  GraphEvaluator::evaluate over a PlonkEvalDomain   /root/reference/src/polynomial/graph_evaluator.rs:361-388,
                                                     src/plonk/eval.rs:93-228, src/nifs/vanilla/mod.rs:100-121
  RelaxedPlonkWitness::fold                          src/plonk/mod.rs:1097-1134
  util::concatenate_with_padding                     src/util.rs:189-193
  fft::{best_fft, fft, ifft}                         src/fft.rs:51-115,160-175

Vectors are torch CUDA uint8 tensors (32 bytes per element, Montgomery form) — torch is used for device
memory only; every computation is a kernel in libmira_b200.so.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _native as N

FQ, FR = N.MIRA_FQ, N.MIRA_FR
ELEM = 32


class EvalError(ValueError):
    """plonk::eval::Error (src/plonk/eval.rs:3-25)."""

    KINDS = {N.MIRA_ERR_EVAL_CHALLENGE: "ChallengeIndexOutOfBoundary", N.MIRA_ERR_EVAL_COLUMN: "ColumnVariableIndexOutOfBoundary",
             N.MIRA_ERR_EVAL_ROW: "RowIndexOutOfBoundary", N.MIRA_ERR_EVAL_WITNESS_INDEX: "InvalidWitnessIndex",
             N.MIRA_ERR_EVAL_PROGRAM: "InvalidProgram"}

    def __init__(self, rc: int, msg: str):
        super().__init__(f"{self.KINDS.get(rc, rc)}: {msg}")
        self.rc = rc
        self.kind = self.KINDS.get(rc, str(rc))


def _check(rc: int):
    if rc == N.MIRA_OK:
        return
    msg = N.last_error()
    if rc in EvalError.KINDS:
        raise EvalError(rc, msg)
    if rc == N.MIRA_ERR_INVALID:
        raise ValueError(msg)
    from .commitment import CudaError
    raise CudaError(msg)


def _dev(t) -> int:
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor (witness vectors live in HBM)")
    return t.device.index or 0


def _stream(stream) -> Optional[int]:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream or None
    return stream or None


def _host32(b) -> bytes:
    b = bytes(b)
    if len(b) != ELEM:
        raise ValueError("expected one 32-byte field element")
    return b


# ------------------------------------------------------------------------------------------- fold
def fold_w(field: int, w1, w2, r: bytes, out=None, stream=None):
    """`*w1 + *r * *w2` element-wise (src/plonk/mod.rs:1100-1110).  Returns `out` (default: a new tensor)."""
    import torch
    if w1.numel() != w2.numel():
        raise ValueError("zip_eq: W1 and W2 lengths differ")
    n = w1.numel() // ELEM
    out = torch.empty_like(w1) if out is None else out
    _check(N.lib().mira_fold_w(field, w1.data_ptr(), w2.data_ptr(), n, _host32(r), out.data_ptr(), _dev(w1), _stream(stream)))
    return out


def fold_e(field: int, e, cross_terms: Sequence, r: bytes, out=None, stream=None):
    """`E[i] + sum_k r^(k+1) * T_k[i]` (src/plonk/mod.rs:1118-1131)."""
    import torch
    n = e.numel() // ELEM
    for t in cross_terms:
        if t.numel() != e.numel():
            raise ValueError("cross term length differs from E")
    out = torch.empty_like(e) if out is None else out
    ptrs = (C.c_void_p * max(len(cross_terms), 1))(*[t.data_ptr() for t in cross_terms])
    _check(N.lib().mira_fold_e(field, e.data_ptr(), ptrs, len(cross_terms), n, _host32(r), out.data_ptr(), _dev(e), _stream(stream)))
    return out


def concatenate_with_padding(cols: Sequence, pad_size: int, device: int = 0, stream=None):
    """src/util.rs:189-193 on device columns; returns one tensor of sum(max(len, pad_size)) elements."""
    import torch
    lens = [c.numel() // ELEM for c in cols]
    total = sum(max(l, pad_size) for l in lens)
    dev = _dev(cols[0]) if cols else device
    out = torch.empty(max(total, 1) * ELEM, dtype=torch.uint8, device=f"cuda:{dev}")
    ptrs = (C.c_void_p * max(len(cols), 1))(*[c.data_ptr() for c in cols])
    clens = (C.c_size_t * max(len(cols), 1))(*lens)
    got = C.c_size_t(0)
    _check(N.lib().mira_concat_pad(ptrs, clens, len(cols), pad_size, out.data_ptr(), total, C.byref(got), dev, _stream(stream)))
    assert got.value == total
    return out[: total * ELEM]


# ------------------------------------------------------------------------------------------- evaluator
class PlonkEvalDomain:
    """`PlonkEvalDomain` (src/plonk/eval.rs:93-106): columns are CUDA tensors, challenges are host bytes."""

    flags = 0

    def __init__(self, num_advice: int, num_lookup: int, challenges: bytes, selectors: Sequence, fixed: Sequence,
                 W1s: Sequence, W2s: Sequence, row_size: Optional[int] = None):
        self.num_advice, self.num_lookup = num_advice, num_lookup
        self.challenges = bytes(challenges)
        self.selectors, self.fixed, self.W1s, self.W2s = list(selectors), list(fixed), list(W1s), list(W2s)
        if row_size is None:          # GetDataForEval::row_size (src/plonk/eval.rs:47-54)
            if self.fixed:
                row_size = self.fixed[0].numel() // ELEM
            elif self.selectors:
                row_size = self.selectors[0].numel()
            else:
                raise ValueError("Fixed & Selectors can't be empty in one time")
        self.row_size = row_size

    def device(self) -> int:
        for group in (self.fixed, self.selectors, self.W1s, self.W2s):
            for t in group:
                return _dev(t)
        return 0

    def _struct(self):
        def arr(ts):
            return (C.c_void_p * max(len(ts), 1))(*[t.data_ptr() for t in ts])
        keep = [arr(self.selectors), arr(self.fixed), arr(self.W1s), arr(self.W2s),
                (C.c_uint64 * max(len(self.W1s), 1))(*[t.numel() // ELEM for t in self.W1s]),
                (C.c_uint64 * max(len(self.W2s), 1))(*[t.numel() // ELEM for t in self.W2s]),
                C.create_string_buffer(self.challenges, max(len(self.challenges), 1))]
        d = N.EvalDomain(self.row_size, len(self.selectors), len(self.fixed), self.num_advice, self.num_lookup,
                         len(self.challenges) // ELEM, len(self.W1s), len(self.W2s), self.flags,
                         C.cast(keep[0], C.c_void_p), C.cast(keep[1], C.c_void_p), C.cast(keep[2], C.c_void_p),
                         C.cast(keep[4], C.c_void_p), C.cast(keep[3], C.c_void_p), C.cast(keep[5], C.c_void_p),
                         C.cast(keep[6], C.c_void_p))
        return d, keep


class LookupEvalDomain(PlonkEvalDomain):
    """`LookupEvalDomain` (src/plonk/eval.rs:84-135): `advice` is a list of separate columns (`&[Vec<F>]`), as
    `evaluate_ls` / `evaluate_ts` build it (src/plonk/lookup.rs:212-276)."""
    flags = N.MIRA_EVAL_LOOKUP_DOMAIN

    def __init__(self, num_lookup: int, challenges: bytes, selectors: Sequence, fixed: Sequence, advice: Sequence,
                 row_size: Optional[int] = None):
        super().__init__(len(advice), num_lookup, challenges, selectors, fixed, advice, [], row_size)


class GraphEvaluator:
    """A serialised `GraphEvaluator` (src/polynomial/graph_evaluator.rs:163-178) bound to the device interpreter.
    Build it from the words `GraphEvaluator::to_bytecode()` emits on the Rust side."""

    def __init__(self, field: int, code: Sequence[int], constants: bytes, rotations: Sequence[int], num_intermediates: int):
        self.field = field
        self._h = C.c_void_p()
        code_arr = (C.c_uint32 * max(len(code), 1))(*code)
        rot_arr = (C.c_int32 * max(len(rotations), 1))(*rotations)
        _check(N.lib().mira_eval_program_create(field, code_arr, len(code), bytes(constants), len(constants) // ELEM, rot_arr,
                                                len(rotations), num_intermediates, C.byref(self._h)))

    def evaluate_rows(self, domain: PlonkEvalDomain, out=None, stream=None, rows=None):
        """`(0..row_size).map(|row| evaluator.evaluate(&domain, row))` (src/nifs/vanilla/mod.rs:109-116), on the GPU.
        rows=(begin, end) evaluates that row range only (the row-range shard of a multi-GPU run)."""
        import torch
        dev = domain.device()
        begin, end = (0, domain.row_size) if rows is None else rows
        count = max(end - begin, 0)
        if out is None:
            out = torch.empty(max(count, 1) * ELEM, dtype=torch.uint8, device=f"cuda:{dev}")
        d, keep = domain._struct()
        _check(N.lib().mira_eval_rows_range(self._h, C.byref(d), begin, end, out.data_ptr(), dev, _stream(stream)))
        return out[: count * ELEM]

    def stats(self) -> dict:
        st = N.EvalStats()
        _check(N.lib().mira_eval_program_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_ if not f.startswith("_")}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().mira_eval_program_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def evaluate_rows_multi(programs: Sequence["GraphEvaluator"], domain: PlonkEvalDomain, outs=None, stream=None, rows=None):
    """All cross terms of a fold in one launch (`mira_eval_rows_multi`): the programs are merged by value numbering so
    that sub-products shared between terms are computed once per row.  Returns one tensor per program, each
    bit-identical to `program.evaluate_rows(domain)`."""
    import torch
    dev = domain.device()
    begin, end = (0, domain.row_size) if rows is None else rows
    count = max(end - begin, 0)
    if outs is None:
        outs = [torch.empty(max(count, 1) * ELEM, dtype=torch.uint8, device=f"cuda:{dev}") for _ in programs]
    d, keep = domain._struct()
    ph = (C.c_void_p * len(programs))(*[p._h.value for p in programs])
    po = (C.c_void_p * len(programs))(*[o.data_ptr() for o in outs])
    _check(N.lib().mira_eval_rows_multi(ph, len(programs), C.byref(d), begin, end, po, dev, _stream(stream)))
    return [o[: count * ELEM] for o in outs]


# ------------------------------------------------------------------------------------------- lookup argument
def evaluate_m(field: int, l, t, out=None, stream=None):
    """`Arguments::evaluate_m` (src/plonk/lookup.rs:278-305): multiplicity of every table value among the lookup
    values, reported at the first occurrence of each distinct table value."""
    import torch
    out = torch.empty_like(t) if out is None else out
    _check(N.lib().mira_lookup_m(field, l.data_ptr(), l.numel() // ELEM, t.data_ptr(), t.numel() // ELEM, out.data_ptr(), _dev(t),
                                 _stream(stream)))
    return out


def evaluate_h_g(field: int, l, t, r: bytes, m, stream=None):
    """`Arguments::evaluate_h_g` (src/plonk/lookup.rs:307-319): h = 1/(l + r), g = m/(t + r), zero denominators -> 0."""
    import torch
    if not (l.numel() == t.numel() == m.numel()):
        raise ValueError("zip_eq: l, t and m lengths differ")
    h, g = torch.empty_like(l), torch.empty_like(t)
    _check(N.lib().mira_lookup_h_g(field, l.data_ptr(), t.data_ptr(), m.data_ptr(), l.numel() // ELEM, _host32(r), h.data_ptr(),
                                   g.data_ptr(), _dev(l), _stream(stream)))
    return h, g


# ------------------------------------------------------------------------------------------- FFT
def best_fft(field: int, a, omega: bytes, log_n: int, stream=None):
    """`best_fft(a, omega, log_n)` in place (src/fft.rs:51-115)."""
    if a.numel() != ELEM << log_n:
        raise ValueError("assert_eq!(n, 1 << log_n)")
    _check(N.lib().mira_fft(field, a.data_ptr(), log_n, _host32(omega), _dev(a), _stream(stream)))
    return a


def fft(field: int, a, log_n: int, stream=None):
    """`fft(a, log_n)` in place (src/fft.rs:160-162)."""
    if a.numel() != ELEM << log_n:
        raise ValueError("assert_eq!(n, 1 << log_n)")
    _check(N.lib().mira_fft_std(field, a.data_ptr(), log_n, 0, _dev(a), _stream(stream)))
    return a


def ifft(field: int, a, log_n: int, stream=None):
    """`ifft(a, log_n)` in place, including the division by 2^log_n (src/fft.rs:165-175)."""
    if a.numel() != ELEM << log_n:
        raise ValueError("assert_eq!(n, 1 << log_n)")
    _check(N.lib().mira_fft_std(field, a.data_ptr(), log_n, 1, _dev(a), _stream(stream)))
    return a
