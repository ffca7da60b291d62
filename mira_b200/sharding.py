"""Point-range sharding of one commitment across ranks.

An MSM is a sum over independent (scalar, point) terms, so the key is split into `world` contiguous
index ranges; rank g keeps bases[lo_g:hi_g] resident on its GPU, commits the matching scalar slice to a
128-byte un-normalised XYZZ partial, and ONE collective — an all_gather of world x 128 bytes (NCCL over
NVLink on GPUs, gloo in the CPU tests) — brings the partials to every rank; rank 0 folds them and
normalises.  There is no other data-path exchange.

The reference has no distributed path (single process + rayon, Cargo.toml:36); this is the B200-native
replacement for rayon's chunk-per-thread split inside halo2's best_multiexp.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

PARTIAL_BYTES = 128


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of range(n) owned by `rank` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_of_prefix(n_commit: int, key_lo: int, key_hi: int) -> Tuple[int, int]:
    """CommitmentKey::commit uses the key prefix ck[..v.len()] (src/commitment.rs:80).  A rank that owns key
    indices [key_lo, key_hi) therefore handles v[key_lo : min(key_hi, len(v))] — possibly empty."""
    lo = min(key_lo, n_commit)
    hi = min(key_hi, n_commit)
    return lo, hi


def row_shard_key_ranges(rows: int, columns: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """Key index ranges of a ROW-sharded prover (evaluation and fold shard by row range).

    The witness is `columns` advice columns of `rows` rows, concatenated column by column
    (`concatenate_with_padding`, src/util.rs:189-193), so key point `c * rows + r` belongs to row r of column c.  A rank
    that owns rows [lo, hi) therefore needs `ck[c * rows + lo .. c * rows + hi)` for every column c, in this order:
    its local witness (the same rows of each column, concatenated) then commits against the whole local key, and a
    local cross-term vector (one value per owned row) against the first range — the prefix the unsharded
    `ck.commit(T)` uses."""
    lo, hi = shard_range(rows, world, rank)
    return [(c * rows + lo, c * rows + hi) for c in range(columns)]


class ShardedCommitmentKey:
    """`CommitmentKey<C>` spread over the ranks of a torch.distributed process group.

    partial_fn(local_scalar_bytes_or_tensor) -> 128 B and combine_fn(world*128 B) -> 64 B default to the
    CUDA implementations (CommitmentKey.partial / combine_partials); the CPU gloo tests inject
    oracle-backed functions to exercise exactly this host logic without a GPU.
    """

    def __init__(self, curve: int, n_total: int, local_key=None, group=None, device: int = 0,
                 partial_fn: Optional[Callable] = None, combine_fn: Optional[Callable] = None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.curve = curve
        self.n_total = n_total
        self.device = device
        self.key_lo, self.key_hi = shard_range(n_total, self.world, self.rank)
        self.local_key = local_key
        self._partial = partial_fn
        self._combine = combine_fn

    def len(self) -> int:
        return self.n_total

    def local_slice(self, n_commit: int) -> Tuple[int, int]:
        return shard_of_prefix(n_commit, self.key_lo, self.key_hi)

    def commit(self, local_scalars, n_commit: int) -> Optional[bytes]:
        """local_scalars: this rank's slice v[lo:hi] (see local_slice).  Returns the 64-byte commitment on
        rank 0 and None elsewhere.  Raises TooLongInput on every rank if n_commit > len (checked first)."""
        import torch
        from .commitment import TooLongInput, combine_partials
        if n_commit > self.n_total:
            raise TooLongInput(n_commit, self.n_total)
        if self._partial is not None:
            part = self._partial(local_scalars)
        else:
            part = self.local_key.partial(local_scalars)
        assert len(part) == PARTIAL_BYTES
        if self.world == 1:
            gathered = part
        else:
            backend = self.dist.get_backend(self.group)
            dev = torch.device("cuda", self.device) if backend == "nccl" else torch.device("cpu")
            mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
            buf = torch.empty(self.world * PARTIAL_BYTES, dtype=torch.uint8, device=dev)
            self.dist.all_gather_into_tensor(buf, mine, group=self.group)
            gathered = buf.cpu().numpy().tobytes()
        if self.rank != 0:
            return None
        if self._combine is not None:
            return self._combine(gathered)
        return combine_partials(self.curve, gathered, self.device)
