"""GPU parity at the sizes and window widths the headline metric is quoted on: the CUDA
commit (through the C ABI) against the CPU oracle on identical bytes, bit for bit, at 2^20 and 2^22 points, at the
window widths the 2^24 (c = 22) and 2^26 (c = 24) configurations use, on a lazily built table behind a caller's
stream, and on Ethereum's public alt_bn128 vectors.  Run on the B200 box: pytest -m gpu."""
import json
import os

import pytest

import oracle_lib as O
import pyref as R

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("curve,logn", [(R.BN254, 20), (R.GRUMPKIN, 20), (R.BN254, 22), (R.GRUMPKIN, 22)])
def test_commit_2_20_and_2_22_vs_oracle(gpu, curve, logn):
    """'every timed result compared bit-for-bit with the oracle at <= 2^20' and one oracle run above it.
    Uniform scalars use the size heuristic's window (c = 19 / 20 here), witness-like ones the sampled window; both from
    device memory and, for the uniform vector, through the host-buffer call with its H2D slices."""
    from mira_b200 import CommitmentKey
    n = 1 << logn
    bases_dev = gpu.gen_bases_dev(curve, 0x4D495241, n)
    bases = gpu.to_bytes(bases_dev)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    windows = {}
    for dist in (0, 1):
        sc = gpu.gen_scalars_dev(curve, 0x4D495242 + dist, n, dist)
        host_sc = gpu.to_bytes(sc)
        want = O.commit(curve, bases, host_sc)
        assert ck.commit_device(sc.data_ptr(), n) == want, (curve, logn, dist)
        windows[dist] = ck.stats()["window_bits"]
        if dist == 0:
            assert ck.commit(host_sc) == want
            # a ragged prefix of the same key (prefix-of-key semantics, src/commitment.rs:80)
            m = n - 12345
            assert ck.commit_device(sc.data_ptr(), m) == O.commit(curve, bases[: 64 * m], host_sc[: 32 * m])
    assert windows[0] >= 18 and windows[1] < windows[0]


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
@pytest.mark.parametrize("c", [22, 24])
def test_headline_window_widths_vs_oracle(gpu, curve, c):
    """c = 22 is what a 2^24-point commit runs with, c = 24 what 2^26 runs with (2^21 / 2^23 buckets, 12 / 11 windows):
    the same kernels, launch shapes and reduction depths, on 2^17 + 1 points so that the oracle finishes in a second."""
    from mira_b200 import CommitmentKey
    n = (1 << 17) + 1
    bases_dev = gpu.gen_bases_dev(curve, 2224, n)
    bases = gpu.to_bytes(bases_dev)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    ck.set_window(c)
    for dist in (0, 1):
        sc = gpu.gen_scalars_dev(curve, 2225 + dist, n, dist)
        host_sc = gpu.to_bytes(sc)
        want = O.commit(curve, bases, host_sc)
        assert ck.commit_device(sc.data_ptr(), n) == want
        st = ck.stats()
        assert st["window_bits"] == c and st["buckets"] == 1 << (c - 1)
        assert ck.commit(host_sc) == want
    # the largest window the ABI accepts
    ck.set_window(26)
    sc = gpu.gen_scalars_dev(curve, 2230, 4097, 0)
    assert ck.commit_device(sc.data_ptr(), 4097) == O.commit(curve, bases[: 64 * 4097], gpu.to_bytes(sc))


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_lazy_table_is_built_on_the_callers_stream(gpu, curve):
    """a fresh key, no prepare(), a non-default stream and >= 2^18 sparse scalars: the fixed-base table
    is built lazily inside the commit and must be complete before the accumulation reads it.  Then the case that was
    reachable after prepare(n): prepare builds the uniform-scalar table, the sampled window of a sparse vector needs
    another one, built on first use on the caller's stream."""
    from mira_b200 import CommitmentKey
    n = (1 << 18) + 3
    bases_dev = gpu.gen_bases_dev(curve, 4242, n)
    bases = gpu.to_bytes(bases_dev)
    s = torch.cuda.Stream()
    sparse = gpu.gen_scalars_dev(curve, 4243, n, 1)
    dense = gpu.gen_scalars_dev(curve, 4244, n, 0)
    want_sparse = O.commit(curve, bases, gpu.to_bytes(sparse))
    want_dense = O.commit(curve, bases, gpu.to_bytes(dense))
    torch.cuda.synchronize()
    ck = CommitmentKey(curve, bases_dev, on_device=True)            # fresh: no table yet
    assert ck.commit_device(sparse.data_ptr(), n, s.cuda_stream) == want_sparse
    ck2 = CommitmentKey(curve, bases_dev, on_device=True)
    ck2.prepare(n)                                                   # the table uniform scalars use
    c_uniform = None
    assert ck2.commit_device(dense.data_ptr(), n, s.cuda_stream) == want_dense
    c_uniform = ck2.stats()["window_bits"]
    assert ck2.commit_device(sparse.data_ptr(), n, s.cuda_stream) == want_sparse     # builds the narrow table now
    assert ck2.stats()["window_bits"] < c_uniform
    # prepare(like=...) builds the table the adaptive path picks, so the commit itself launches no k_precompute
    ck3 = CommitmentKey(curve, bases_dev, on_device=True)
    ck3.prepare(n, like=sparse.data_ptr(), like_on_device=True)
    assert ck3.commit_device(sparse.data_ptr(), n, s.cuda_stream) == want_sparse
    assert ck3.stats()["window_bits"] == ck2.stats()["window_bits"]


def test_null_stream_orders_producer_and_commit(gpu):
    """NULL means the legacy default stream for every entry point that reads device scalars, as it
    does for the witness kernels: fold_w(..., NULL) followed by commit_device(..., NULL) needs no synchronisation."""
    from mira_b200 import CommitmentKey
    from mira_b200 import witness as W
    curve, field = R.BN254, R.FR
    n = (1 << 19) + 7
    bases_dev = gpu.gen_bases_dev(curve, 77, n)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    ck.prepare(n)
    w1 = gpu.gen_scalars_dev(curve, 78, n, 0)
    w2 = gpu.gen_scalars_dev(curve, 79, n, 0)
    r = R.to_mont_bytes(R.gen_scalar(curve, 80, 0), R.scalar_mod(curve))
    want_vec = O.fold_w(field, gpu.to_bytes(w1), gpu.to_bytes(w2), r)
    want = O.commit(curve, gpu.to_bytes(bases_dev), want_vec)
    torch.cuda.synchronize()
    for _ in range(3):
        out = torch.zeros_like(w1)
        W.fold_w(field, w1, w2, r, out=out)                      # default (NULL) stream, asynchronous
        got = ck.commit_device(out.data_ptr(), n)                # default (NULL) stream: ordered behind the fold
        assert got == want


def test_eip196_vectors_on_the_gpu(gpu):
    """Ethereum's alt_bn128 ecAdd / ecMul vectors (tests/golden/eip196_alt_bn128.json, an authority outside this
    repository) through CommitmentKey::commit on the GPU: an MSM with unit scalars is ecAdd, an MSM of one term ecMul."""
    from mira_b200 import CommitmentKey
    with open(os.path.join(os.path.dirname(__file__), "golden", "eip196_alt_bn128.json")) as f:
        v = json.load(f)

    def pt(xy):
        return R.point_to_bytes((int(xy[0], 16), int(xy[1], 16)), R.BN254)
    one = R.to_mont_bytes(1, R.R_)
    for c in v["ec_add"]:
        ck = CommitmentKey(R.BN254, pt(c["a"]) + pt(c["b"]))
        ck.check_on_curve()
        assert ck.commit(one + one) == pt(c["sum"])
    for c in v["ec_mul"]:
        ck = CommitmentKey(R.BN254, pt(c["p"]))
        assert ck.commit(R.to_mont_bytes(int(c["k"], 16), R.R_)) == pt(c["product"])
    # the two chfast1 cases in one 3-term MSM, at every window width class (tiny, medium, headline)
    ca, cm = v["ec_add"][1], v["ec_mul"][1]
    ck = CommitmentKey(R.BN254, pt(cm["p"]) + pt(ca["a"]) + pt(ca["b"]))
    want = R.point_to_bytes(R.add(tuple(int(t, 16) for t in cm["product"]), tuple(int(t, 16) for t in ca["sum"]), R.BN254), R.BN254)
    for c in (0, 4, 13, 22):
        ck.set_window(c)
        assert ck.commit(R.to_mont_bytes(int(cm["k"], 16), R.R_) + one + one) == want


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_single_process_sharded_key_equals_single_device(gpu, curve):
    """mira_msm_ctx_create_sharded: one process, the key cut into point ranges over a device list,
    `commit` on it returns the bytes a single-device key returns (and the oracle's) for full, prefix, ragged, tiny and
    empty vectors.  With one GPU the list names it several times (several ranges on one device: same code path, same
    threads); with more, one range per GPU."""
    from mira_b200 import CommitmentKey, TooLongInput
    n = (1 << 18) + 11
    bases = gpu.to_bytes(gpu.gen_bases_dev(curve, 9001, n))
    ndev = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]] if ndev < 2 else [list(range(ndev)), [0, 1, 0]]
    single = CommitmentKey(curve, bases)
    for devices in lists:
        ck = CommitmentKey.sharded(curve, bases, devices)
        assert ck.len() == n and ck.num_devices() == len(devices)
        ck.check_on_curve()
        for dist, m in ((0, n), (1, n), (0, n - 70001), (0, min(n, n // len(devices) + 5)), (0, 3), (0, 1), (0, 0)):
            sc = O.gen_scalars(curve, 9002 + dist, m, dist)
            got = ck.commit(sc)
            assert got == single.commit(sc), (devices, dist, m)
            if m in (n, 3, 0):
                assert got == O.commit(curve, bases[: 64 * m], sc)
        assert ck.stats()["entries"] == 0                      # the empty commit
        ck.commit(O.gen_scalars(curve, 9010, n, 0))
        assert ck.stats()["entries"] >= 10 * n                 # pairs summed over the shards (each picks its own window)
        with pytest.raises(TooLongInput):
            ck.commit(bytes(32 * (n + 1)))
        with pytest.raises(ValueError):
            ck.commit_device(0x1000, 1)                        # a device vector lives on ONE device: not on a sharded key
        ck.close()


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_very_sparse_long_vectors(gpu, curve):
    """Vectors long enough for the sampled window and the exact-pair-count path (>= 2^18 scalars) with almost nothing in
    them: all zeros (commitment = identity), a single non-zero scalar where no sample chunk looks, non-zeros only inside
    the first sampled chunk, and 1 % ones — from device and from host memory."""
    import numpy as np
    from mira_b200 import CommitmentKey
    n = (1 << 18) + 77
    bases_dev = gpu.gen_bases_dev(curve, 606, n)
    bases = gpu.to_bytes(bases_dev)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    sm = R.scalar_mod(curve)
    one = np.frombuffer(R.to_mont_bytes(1, sm), dtype=np.uint8)
    big = np.frombuffer(R.to_mont_bytes(sm - 2, sm), dtype=np.uint8)
    rng = np.random.default_rng(17)
    cases = []
    z = np.zeros((n, 32), dtype=np.uint8)
    cases.append(("all zeros", z.copy()))
    v = z.copy(); v[n - 1] = big
    cases.append(("one scalar at the very end", v))
    v = z.copy(); v[:100] = one; v[7] = big
    cases.append(("non-zeros in the first chunk only", v))
    v = z.copy(); v[rng.choice(n, n // 100, replace=False)] = one
    cases.append(("1 % ones", v))
    for name, v in cases:
        host = v.tobytes()
        want = O.commit(curve, bases, host)
        dev = torch.frombuffer(bytearray(host), dtype=torch.uint8).cuda()
        assert ck.commit_device(dev.data_ptr(), n) == want, name
        assert ck.commit(host) == want, name
    assert O.commit(curve, bases, cases[0][1].tobytes()) == bytes(64)


def test_dense_vector_with_a_heavy_hitter_vs_oracle(gpu):
    """A dense vector in which one full-length value is repeated in a quarter of the scalars (a column filled with -1,
    say): every window then has ONE bucket holding a quarter of that window's pairs.  With the window chosen from the
    sample the heavy hitter is detected and the LSD sort is kept; with a forced window there is no sample, so the MSD
    partition runs and its per-group pass meets groups far beyond its shared-memory capacity.  Both must give the
    oracle's bytes (2^22 points: enough pairs for the MSD path to be chosen)."""
    from mira_b200 import CommitmentKey
    import numpy as np
    curve, n = R.BN254, 1 << 22
    bases_dev = gpu.gen_bases_dev(curve, 0x48454156, n)
    bases = gpu.to_bytes(bases_dev)
    sc = gpu.gen_scalars_dev(curve, 0x48454157, n, 0)
    host = np.frombuffer(bytearray(gpu.to_bytes(sc)), dtype=np.uint8).reshape(n, 32).copy()
    rep = host[5].copy()                                            # a random full-length scalar
    host[np.arange(0, n, 4)] = rep                                  # ... in every fourth position
    host_sc = host.tobytes()
    want = O.commit(curve, bases, host_sc)
    d = torch.frombuffer(bytearray(host_sc), dtype=torch.uint8).cuda()
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    assert ck.commit_device(d.data_ptr(), n) == want                # sampled: heavy hitter seen, LSD sort
    ck.set_window(20)
    assert ck.commit_device(d.data_ptr(), n) == want                # forced window: MSD partition, oversized groups
    assert ck.commit(host_sc) == want
