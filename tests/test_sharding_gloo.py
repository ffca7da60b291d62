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


# ---------------------------------------------------------------------------------- row-sharded fold step
def test_row_shard_key_ranges_cover_the_key_once():
    from mira_b200.sharding import row_shard_key_ranges
    for rows, cols, world in ((64, 7, 2), (64, 14, 3), (1 << 19, 14, 8), (5, 3, 4)):
        seen = []
        for r in range(world):
            rg = row_shard_key_ranges(rows, cols, world, r)
            assert len(rg) == cols
            assert len({hi - lo for lo, hi in rg}) == 1                      # the same rows of every column
            assert rg[0] == shard_range(rows, world, r)                      # first range = the cross terms' key prefix
            seen += [i for lo, hi in rg for i in range(lo, hi)] if rows <= 64 else []
        if rows <= 64:
            assert sorted(seen) == list(range(rows * cols))


def _fold_worker(rank, world, port, q):
    """One fold step of the secondary circuit (k = 6) with the rows cut across `world` ranks: local evaluation, local
    commitments and local fold through the CPU oracle, ONE all_gather of the partial commitments, sum on rank 0."""
    import torch.distributed as dist
    import graph_evaluator_model as G
    import oracle_lib as O
    import pyref as R
    from mira_b200.sharding import row_shard_key_ranges
    from witness_util import pack_program
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        curve, field, rows = R.GRUMPKIN, R.FQ, 64
        progs, meta = G.cross_term_programs(5, 1, R.P)
        progs = [pack_program(p) for p in progs]
        cols = meta["num_advice"]
        assert all(r == 0 for p in progs for r in p["rotations"])
        n_w = cols * rows
        bases = O.gen_bases(curve, 11, n_w)
        fixed = [O.gen_scalars(curve, 20 + i, rows, 1) for i in range(meta["num_fixed"])]
        w1, w2, e = O.gen_scalars(curve, 1, n_w, 0), O.gen_scalars(curve, 3, n_w, 1), O.gen_scalars(curve, 2, rows, 0)
        ch, r_ = O.gen_scalars(curve, 4, meta["num_challenges"], 0), O.gen_scalars(curve, 5, 1, 0)

        def step(key, fx, a1, a2, ee, n_rows):
            dom = {"row_size": n_rows, "fixed": fx, "w1": [a1], "w2": [a2], "challenges": ch, "num_advice": cols}
            ts = [O.eval_rows(field, p, dom) for p in progs]
            commits = [O.commit(curve, key, a2)] + [O.commit(curve, key[:64 * n_rows], t) for t in ts]
            return commits, O.fold_w(field, a1, a2, r_), O.fold_e(field, ee, ts, r_)

        ranges = row_shard_key_ranges(rows, cols, world, rank)
        lo, hi = ranges[0]
        cut = lambda v, size: b"".join(v[size * a:size * b] for a, b in ranges)            # the rank's rows of every column
        part, w_loc, e_loc = step(cut(bases, 64), [f[32 * lo:32 * hi] for f in fixed], cut(w1, 32), cut(w2, 32), e[32 * lo:32 * hi], hi - lo)
        mine = torch.frombuffer(bytearray(b"".join(part)), dtype=torch.uint8)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                                                  # the only exchange
        if rank == 0:
            want, w_all, e_all = step(bases, fixed, w1, w2, e, rows)
            got = []
            for j in range(len(part)):
                acc = bytes(64)
                for g in gathered:
                    acc = O.point_add(curve, acc, g.numpy().tobytes()[64 * j:64 * j + 64])
                got.append(acc)
            q.put(("commits", got == want))
        else:
            w_all = O.fold_w(field, w1, w2, r_)
            e_all = None
        q.put((f"fold{rank}", w_loc == cut(w_all, 32) and (e_all is None or e_loc == e_all[32 * lo:32 * hi])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_fold_step_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fold_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world + 1))
    assert res["commits"] and all(res[f"fold{r}"] for r in range(world))
