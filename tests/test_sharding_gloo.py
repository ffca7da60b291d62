"""World-size-2/3 gloo tests (CPU) of the multi-rank host logic: shard ranges, prefix-of-key slicing,
the 128-byte all_gather and the fold on rank 0.  The per-rank partial and the fold are injected from the
CPU oracle here (no GPU in this container); the GPU versions of the same two functions are covered by
tests/test_gpu_parity.py::test_sharded_partials_combine_to_full_commit."""
import os
import socket
import sys

import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mira_b200.sharding import shard_of_prefix, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 7, 8, 9, 1000, (1 << 26), (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_prefix_slicing():
    assert shard_of_prefix(5, 0, 10) == (0, 5)
    assert shard_of_prefix(5, 10, 20) == (5, 5)      # empty: commit shorter than this rank's range start
    assert shard_of_prefix(15, 10, 20) == (10, 15)
    assert shard_of_prefix(30, 10, 20) == (10, 20)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, curve, n_key, n_commit, q):
    import torch.distributed as dist
    import oracle_lib as O
    from mira_b200.sharding import ShardedCommitmentKey
    from mira_b200.commitment import TooLongInput
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bases = O.gen_bases(curve, 77, n_key)
        scalars = O.gen_scalars(curve, 78, n_commit)
        lo, hi = shard_range(n_key, world, rank)
        local_bases = bases[64 * lo:64 * hi]

        def partial_fn(local_scalars: bytes) -> bytes:
            # oracle stand-in for CommitmentKey.partial: affine sum as XYZZ with ZZ = ZZZ = 1 (or identity)
            k = len(local_scalars) // 32
            aff = O.commit(curve, local_bases, local_scalars) if k else bytes(64)
            one = O.fe_from_u64(0 if curve == 0 else 1, 1)
            return aff + (one + one if aff != bytes(64) else bytes(64))

        def combine_fn(parts: bytes) -> bytes:
            acc = bytes(64)
            for g in range(len(parts) // 128):
                acc = O.point_add(curve, acc, parts[128 * g:128 * g + 64])
            return acc

        sk = ShardedCommitmentKey(curve, n_key, partial_fn=partial_fn, combine_fn=combine_fn)
        s_lo, s_hi = sk.local_slice(n_commit)
        got = sk.commit(scalars[32 * s_lo:32 * s_hi], n_commit)
        if rank == 0:
            want = O.commit(curve, bases, scalars)
            q.put(("ok", got == want))
        else:
            assert got is None
        try:
            sk.commit(b"", n_key + 1)
            too_long = False
        except TooLongInput:
            too_long = True
        if rank == 0:
            q.put(("toolong", too_long))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,curve,n_key,n_commit", [(2, 0, 257, 257), (2, 1, 300, 120), (3, 0, 100, 33), (2, 0, 64, 0)])
def test_sharded_commit_over_gloo(world, curve, n_key, n_commit):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, curve, n_key, n_commit, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res["ok"] and res["toolong"]
