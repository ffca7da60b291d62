"""Host-side mirror of the reference's `CommitmentKey<C>` (/root/reference/src/commitment.rs:26-167)
over the C ABI.  Same method names, argument meaning and error behaviour, so parity tests read like
the reference's own: `CommitmentKey.commit(v)` == `ck.commit(&v)`.

The reference is Rust; this image has no Rust toolchain, so this Python class (and the C++ header
include/mira_commitment.hpp) stand where the Rust shim of INTEGRATION.md would.  All arithmetic is
in libmira_b200.so (CUDA, sm_100a); nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _native as N

BN254_G1 = 0       # halo2curves::bn256::G1Affine
GRUMPKIN_G1 = 1    # halo2curves::grumpkin::G1Affine
SCALAR_BYTES = 32
POINT_BYTES = 64
IDENTITY = bytes(POINT_BYTES)   # CommitmentKey::default_value() == C::identity() == (0, 0)


class TooLongInput(ValueError):
    """commitment::Error::TooLongInput { input_len, limit } (src/commitment.rs:20-24)."""

    def __init__(self, input_len: int, limit: int):
        super().__init__(f"Can't commit too long input: input len: {input_len}, but limit is {limit}")
        self.input_len = input_len
        self.limit = limit


class CudaError(RuntimeError):
    """Any CUDA failure.  The reference has no such error path, so callers abort (no CPU fallback)."""


class NotOnCurve(ValueError):
    """io::ErrorKind::InvalidData "Wrong file in cache, some ptr out of curve" (src/commitment.rs:145-153)."""


def _check(rc: int, n: int = 0, limit: int = 0):
    if rc == N.MIRA_OK:
        return
    msg = N.last_error()
    if rc == N.MIRA_ERR_TOO_LONG_INPUT:
        raise TooLongInput(n, limit)
    if rc == N.MIRA_ERR_NOT_ON_CURVE:
        raise NotOnCurve(msg)
    if rc == N.MIRA_ERR_INVALID:
        raise ValueError(msg)
    raise CudaError(msg)


def _as_ptr(buf):
    """bytes / bytearray / memoryview / numpy array / torch CPU tensor -> (address, nbytes, keepalive)."""
    if hasattr(buf, "data_ptr"):          # torch tensor (CPU or CUDA)
        return buf.data_ptr(), buf.numel() * buf.element_size(), buf
    if hasattr(buf, "ctypes") and hasattr(buf, "nbytes"):   # numpy
        return buf.ctypes.data, buf.nbytes, buf
    if isinstance(buf, bytes):
        return C.cast(C.c_char_p(buf), C.c_void_p).value, len(buf), buf
    mv = memoryview(buf)
    arr = (C.c_char * mv.nbytes).from_buffer(mv)
    return C.addressof(arr), mv.nbytes, (arr, mv)


class CommitmentKey:
    """`CommitmentKey<C>`: an immutable vector of affine generators resident in GPU memory."""

    def __init__(self, curve: int, bases, device: int = 0, on_device: bool = False, n: Optional[int] = None):
        L = N.lib()
        ptr, nbytes, keep = _as_ptr(bases) if bases is not None else (None, 0, None)
        if n is None:
            n = nbytes // POINT_BYTES
        self.curve = curve
        self.device = device
        self._n = n
        self._image = bases          # the memory image of `ck: Box<[C]>` this key was made from (for save_to_file)
        self._image_on_device = on_device
        self._ctx = C.c_void_p()
        _check(L.mira_msm_ctx_create(curve, ptr, n, 1 if on_device else 0, device, C.byref(self._ctx)))

    @classmethod
    def sharded(cls, curve: int, bases, devices) -> "CommitmentKey":
        """The key spread over several GPUs of this process (mira_msm_ctx_create_sharded): `bases` in HOST memory,
        contiguous point ranges on `devices`; `commit(host_vector)` then runs on all of them and returns the same 64
        bytes a single-device key returns.  Device-vector methods are not available on such a key."""
        L = N.lib()
        ptr, nbytes, keep = _as_ptr(bases)
        self = cls.__new__(cls)
        self.curve, self.device = curve, int(devices[0])
        self._n = nbytes // POINT_BYTES
        self._image, self._image_on_device = bases, False
        self._ctx = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        _check(L.mira_msm_ctx_create_sharded(curve, ptr, self._n, devs, len(devices), C.byref(self._ctx)))
        return self

    def num_devices(self) -> int:
        return int(N.lib().mira_msm_ctx_num_devices(self._ctx))

    # -- reference API -------------------------------------------------------------------------
    @staticmethod
    def default_value() -> bytes:
        return IDENTITY

    def len(self) -> int:
        return int(N.lib().mira_msm_ctx_len(self._ctx))

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def commit(self, v) -> bytes:
        """`ck.commit(&v)`: v = n x 32 B Montgomery scalars in HOST memory -> 64 B affine point."""
        ptr, nbytes, keep = _as_ptr(v)
        n = nbytes // SCALAR_BYTES
        out = C.create_string_buffer(POINT_BYTES)
        _check(N.lib().mira_msm_commit(self._ctx, ptr, n, out), n, self._n)
        return out.raw

    @classmethod
    def load_from_file(cls, file_path: str, k: int, curve: int, device: int = 0) -> "CommitmentKey":
        """Raw memory-image key file, 64 * 2^k bytes (src/commitment.rs:109-124)."""
        want = POINT_BYTES << k
        with open(file_path, "rb") as f:
            data = f.read(want)
        if len(data) != want:
            raise IOError(f"{file_path}: expected {want} bytes, got {len(data)}")   # read_exact failure
        return cls(curve, data, device)

    @classmethod
    def load_or_setup_cache(cls, cache_folder: str, label: str, k: int, curve: int, device: int = 0,
                            setup_seed: int = 0x4D495241) -> "CommitmentKey":
        """{cache_folder}/{label}/{k}.bin if present (with the on-curve check, src/commitment.rs:134-156),
        else generate, save and return a key (src/commitment.rs:157-166)."""
        path = os.path.join(cache_folder, label, f"{k}.bin")
        if os.path.exists(path):
            key = cls.load_from_file(path, k, curve, device)
            key.check_on_curve()
            return key
        key = cls.setup(k, label.encode(), curve, device, seed=setup_seed)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        key.save_to_file(path)
        return key

    @classmethod
    def setup(cls, k: int, label: bytes, curve: int, device: int = 0, seed: int = 0x4D495241) -> "CommitmentKey":
        """Deterministic synthetic key of 2^k points generated on the GPU.

        NOT bit-compatible with the reference's `setup` (src/commitment.rs:52-76): that one hashes
        Shake256(label) blocks to the curve with halo2curves' `hash_to_curve` (un-vendored) in
        `par_bridge` order (unordered), so keys are only ever compared through their cached bytes
        (SURVEY.md §3.4).  Use load_from_file / load_or_setup_cache for reference-made keys."""
        assert k < 32
        import torch
        n = 1 << k
        mix = seed ^ int.from_bytes(label[:8].ljust(8, b"\0"), "little")
        buf = torch.empty(n * POINT_BYTES, dtype=torch.uint8, device=f"cuda:{device}")
        _check(N.lib().mira_gen_bases(curve, mix & 0xFFFFFFFFFFFFFFFF, 0, n, device, buf.data_ptr()))
        return cls(curve, buf, device, on_device=True)

    def save_to_file(self, file_path: str):
        """Writes the key as the raw memory image of `[C]`, 64 bytes per point (src/commitment.rs:96-107)."""
        if self._image is None:
            data = b""
        elif self._image_on_device:
            data = self._image.cpu().numpy().tobytes()[: self._n * POINT_BYTES]
        else:
            ptr, nbytes, keep = _as_ptr(self._image)
            data = C.string_at(ptr, self._n * POINT_BYTES)
        with open(file_path, "wb") as f:
            f.write(data)

    # -- extensions used by the GPU pipeline ------------------------------------------------------
    def check_on_curve(self):
        _check(N.lib().mira_msm_ctx_check_on_curve(self._ctx))

    def prepare(self, n: int, like=None, like_on_device: bool = False):
        """Build the fixed-base table commits of length n use (otherwise built lazily on first commit).  `like`: a
        vector (host buffer, or device pointer with like_on_device) whose sampled density picks the window, as a
        commit of it would (sparse witness columns use a much narrower window than uniform scalars)."""
        if like is None:
            _check(N.lib().mira_msm_ctx_prepare(self._ctx, n), n, self._n)
            return
        ptr = like if like_on_device else _as_ptr(like)[0]
        _check(N.lib().mira_msm_ctx_prepare_for(self._ctx, ptr, n, 1 if like_on_device else 0), n, self._n)

    def commit_device(self, scalars_dev_ptr: int, n: int, stream: int = 0) -> bytes:
        out = C.create_string_buffer(POINT_BYTES)
        _check(N.lib().mira_msm_commit_device(self._ctx, scalars_dev_ptr, n, out, stream or None), n, self._n)
        return out.raw

    def scalars_device(self):
        """Device copy of the scalars of the last host-buffer commit, as a CUDA uint8 tensor view (no copy); None if
        there is none.  Valid until the next host-buffer commit on this key."""
        import torch
        n = C.c_size_t(0)
        ptr = N.lib().mira_msm_scalars_device(self._ctx, C.byref(n))
        if not ptr or not n.value:
            return None

        class _View:
            __cuda_array_interface__ = {"shape": (n.value * SCALAR_BYTES,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_View(), device=f"cuda:{self.device}")

    def commit_batch_device(self, scalar_dev_ptrs, n: int, stream: int = 0):
        """`vs.iter().map(|v| ck.commit(v))` for device vectors of equal length n in one call (one sort, one
        accumulation, one reduction for all of them).  Returns a list of 64-byte commitments."""
        k = len(scalar_dev_ptrs)
        ptrs = (C.c_void_p * max(k, 1))(*scalar_dev_ptrs)
        out = C.create_string_buffer(POINT_BYTES * max(k, 1))
        _check(N.lib().mira_msm_commit_batch(self._ctx, ptrs, k, n, out, stream or None), n, self._n)
        return [out.raw[POINT_BYTES * i:POINT_BYTES * (i + 1)] for i in range(k)]

    def partial(self, scalars, n: Optional[int] = None, on_device: bool = False, stream: int = 0) -> bytes:
        """Un-normalised XYZZ partial sum (128 B) of this rank's slice."""
        if on_device:
            ptr = scalars
        else:
            ptr, nbytes, keep = _as_ptr(scalars)
            n = nbytes // SCALAR_BYTES if n is None else n
        out = C.create_string_buffer(128)
        _check(N.lib().mira_msm_partial(self._ctx, ptr, n, 1 if on_device else 0, out, stream or None), n, self._n)
        return out.raw

    def partial_batch_device(self, scalar_dev_ptrs, n: int, out_dev_ptr: int, stream: int = 0):
        """XYZZ partial sums (128 B each) of `scalar_dev_ptrs` against this key shard, written to device memory at
        `out_dev_ptr` on `stream` without synchronising the host (row-sharded provers)."""
        k = len(scalar_dev_ptrs)
        ptrs = (C.c_void_p * max(k, 1))(*scalar_dev_ptrs)
        _check(N.lib().mira_msm_partial_batch_dev(self._ctx, ptrs, k, n, out_dev_ptr, stream or None), n, self._n)

    def stats(self) -> dict:
        st = N.MsmStats()
        _check(N.lib().mira_msm_get_stats(self._ctx, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def set_profiling(self, on: bool):
        _check(N.lib().mira_msm_set_profiling(self._ctx, 1 if on else 0))

    def set_window(self, c: int):
        _check(N.lib().mira_msm_set_window(self._ctx, c))

    def set_adaptive_window(self, on: bool):
        """Window picked from a sample of the scalars (default) or from the vector length alone."""
        _check(N.lib().mira_msm_set_adaptive_window(self._ctx, 1 if on else 0))

    def set_slice_min(self, n: int):
        """Host-buffer commits are pipelined in up to 4 slices of >= n scalars behind their H2D copies (0: off)."""
        _check(N.lib().mira_msm_set_slice_min(self._ctx, n))

    def set_pipeline(self, slices: int, min_scalars_per_slice: int = 0):
        """Slice pipeline (part k+1 sorted on a second stream while part k is accumulated).  Host-buffer commits use it
        on their H2D slices unless `slices` == 1; device-resident commits only when `slices` >= 2 (default 0: whole)."""
        _check(N.lib().mira_msm_set_pipeline(self._ctx, slices, min_scalars_per_slice))

    AFFINE_THREAD_LOCAL_PAIRS = -2

    def set_affine_levels(self, levels: int):
        """Experimental (default 0): batched-affine pre-reduction levels before the XYZZ accumulation;
        AFFINE_THREAD_LOCAL_PAIRS (-2) = the thread-local pair pre-addition of round 2 (measured slower)."""
        _check(N.lib().mira_msm_set_affine_levels(self._ctx, levels))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            N.lib().mira_msm_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def combine_partials_device(curve: int, partials_dev_ptr: int, n_ranks: int, n_commits: int, rank_stride: int, device: int = 0,
                            stream: int = 0):
    """Commitments from the gathered per-rank partial buffers (device memory, `[rank][commit]` x 128 B with
    `rank_stride` bytes between ranks): sums over the ranks, normalises, returns `n_commits` 64-byte points."""
    out = C.create_string_buffer(POINT_BYTES * max(n_commits, 1))
    _check(N.lib().mira_msm_combine_dev(curve, partials_dev_ptr, n_ranks, n_commits, rank_stride, device, out, stream or None))
    return [out.raw[POINT_BYTES * i:POINT_BYTES * (i + 1)] for i in range(n_commits)]


def combine_partials(curve: int, partials: bytes, device: int = 0) -> bytes:
    """Fold 128-byte XYZZ partials (one per rank) into the normalised affine commitment."""
    out = C.create_string_buffer(POINT_BYTES)
    _check(N.lib().mira_msm_combine(curve, partials, len(partials) // 128, device, out))
    return out.raw
