"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical bytes.
Bit-exact is the bar everywhere (integer arithmetic).  Run on the B200 box: pytest -m gpu."""
import os
import random

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


def _rand_fe(rng, m, n):
    edge = [0, 1, 2, m - 1, m - 2, 1 << 253, (1 << 254) % m, (1 << 32) - 1, 1 << 32, m >> 1]
    vals = edge + [rng.randrange(m) for _ in range(n - len(edge))]
    return vals


@pytest.mark.parametrize("field", [R.FQ, R.FR])
def test_field_kernels_vs_oracle(gpu, field):
    m = R.FIELD_MOD[field]
    rng = random.Random(5150 + field)
    n = 4096
    av, bv = _rand_fe(rng, m, n), _rand_fe(random.Random(77 + field), m, n)
    rng.shuffle(bv)
    a = b"".join(R.to_mont_bytes(v, m) for v in av)
    b = b"".join(R.to_mont_bytes(v, m) for v in bv)
    assert gpu.field_op(field, 0, a, b) == O.fe_mul_many(field, a, b)
    assert gpu.field_op(field, 3, a) == O.fe_mul_many(field, a, a)
    assert gpu.field_op(field, 1, a, b) == b"".join(R.to_mont_bytes(x + y, m) for x, y in zip(av, bv))
    assert gpu.field_op(field, 2, a, b) == b"".join(R.to_mont_bytes(x - y, m) for x, y in zip(av, bv))
    assert gpu.field_op(field, 5, a) == b"".join(v.to_bytes(32, "little") for v in av)
    canon = b"".join(v.to_bytes(32, "little") for v in av)
    assert gpu.field_op(field, 6, canon) == a
    k = 256
    inv = gpu.field_op(field, 4, a[: 32 * k])
    assert inv == b"".join(O.fe_inv(field, a[32 * i:32 * i + 32]) for i in range(k))


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_point_kernels_vs_oracle(gpu, curve):
    n = 512
    P = bytearray(O.gen_bases(curve, 11, n))
    Q = bytearray(O.gen_bases(curve, 12, n))
    # exceptional cases: identity operands, P == Q, P == -Q
    Q[0:64] = bytes(64)
    P[64:128] = bytes(64)
    Q[128:192] = P[128:192]
    Q[192:256] = O.point_neg(curve, bytes(P[192:256]))
    P[256:320] = bytes(64); Q[256:320] = bytes(64)
    P, Q = bytes(P), bytes(Q)
    want = b"".join(O.point_add(curve, P[64 * i:64 * i + 64], Q[64 * i:64 * i + 64]) for i in range(n))
    assert gpu.point_op(curve, 0, P, Q) == want          # mixed add
    assert gpu.point_op(curve, 1, P, Q) == want          # full XYZZ add path
    dbl = b"".join(O.point_add(curve, P[64 * i:64 * i + 64], P[64 * i:64 * i + 64]) for i in range(n))
    assert gpu.point_op(curve, 2, P) == dbl
    sm = R.scalar_mod(curve)
    ks = [0, 1, 2, 3, 0xFFFFFFFF, 0x80000000] + [random.Random(i).randrange(1 << 32) for i in range(n - 6)]
    K = b"".join(k.to_bytes(4, "little") + bytes(60) for k in ks)
    want = b"".join(O.scalar_mul(curve, P[64 * i:64 * i + 64], R.to_mont_bytes(ks[i], sm)) for i in range(n))
    assert gpu.point_op(curve, 3, P, K) == want


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_device_generators_match_oracle(gpu, curve):
    seed, n = 0x4D495241, 3000
    for dist in (0, 1):
        assert gpu.to_bytes(gpu.gen_scalars_dev(curve, seed, n, dist)) == O.gen_scalars(curve, seed, n, dist)
    assert gpu.to_bytes(gpu.gen_scalars_dev(curve, seed, 100, 0, first=12345)) == O.gen_scalars(curve, seed, 100, 0, first=12345)
    assert gpu.to_bytes(gpu.gen_bases_dev(curve, seed + 1, n)) == O.gen_bases(curve, seed + 1, n)
    assert gpu.to_bytes(gpu.gen_bases_dev(curve, seed + 1, 64, first=999)) == O.gen_bases(curve, seed + 1, 64, first=999)


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 31, 32, 33, 100, 1000, 4097])
def test_commit_small_vs_oracle(gpu, curve, n):
    from mira_b200 import CommitmentKey
    bases = O.gen_bases(curve, 100 + n, max(n, 1) + 5)
    ck = CommitmentKey(curve, bases)
    assert ck.len() == max(n, 1) + 5
    for dist in (0, 1):
        sc = O.gen_scalars(curve, 200 + n, n, dist)
        want = O.commit(curve, bases, sc)
        assert ck.commit(sc) == want
    if 0 < n <= 100:
        assert ck.commit(sc) == R.commit_bytes(curve, bases, sc)   # independent Python big-int model


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
@pytest.mark.parametrize("c", [2, 5, 8, 13, 16, 20])
def test_commit_all_window_widths(gpu, curve, c):
    from mira_b200 import CommitmentKey
    n = 3000
    bases = O.gen_bases(curve, 31, n)
    ck = CommitmentKey(curve, bases)
    ck.set_window(c)
    for dist in (0, 1):
        sc = O.gen_scalars(curve, 32 + c, n, dist)
        assert ck.commit(sc) == O.commit(curve, bases, sc)
    st = ck.stats()
    assert st["window_bits"] == c and st["windows"] == (255 + c - 1) // c


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_commit_edge_cases(gpu, curve):
    from mira_b200 import CommitmentKey, TooLongInput, IDENTITY
    sm = R.scalar_mod(curve)
    n = 600
    bases = bytearray(O.gen_bases(curve, 4242, n))
    bases[64 * 3:64 * 4] = bytes(64)                                   # identity generator
    bases[64 * 5:64 * 6] = bases[64 * 4:64 * 5]                        # duplicated generator
    bases[64 * 7:64 * 8] = O.point_neg(curve, bytes(bases[64 * 6:64 * 7]))  # a generator and its negation
    for i in range(100, 200):                                          # a long run of one repeated point
        bases[64 * i:64 * i + 64] = bases[64 * 99:64 * 100]
    bases = bytes(bases)
    vals = [R.gen_scalar(curve, 4243, i) for i in range(n)]
    vals[0] = 0; vals[1] = 1; vals[2] = sm - 1; vals[4] = vals[5] = 5; vals[6] = vals[7] = 9
    for i in range(100, 200):
        vals[i] = 7                                                    # same point, same scalar -> same bucket: doublings
    for i in range(300, 400):
        vals[i] = 1 << 200                                             # only one high window set
    sc = b"".join(R.to_mont_bytes(v, sm) for v in vals)
    ck = CommitmentKey(curve, bases)
    for c, levels in ((0, 0), (4, 0), (11, 0), (4, 3), (11, 6), (2, 2), (0, -2), (4, -2), (11, -2)):     # -2: thread-local pairs
        ck.set_window(c)
        ck.set_affine_levels(levels)      # the batched-affine levels meet the same duplicates / negations / identities
        assert ck.commit(sc) == O.commit(curve, bases, sc)
        assert ck.commit(bytes(32 * n)) == IDENTITY                    # zero vector -> (0,0)
        assert ck.commit(sc[: 32 * 10]) == O.commit(curve, bases, sc[: 32 * 10])   # prefix of the key
    two = CommitmentKey(curve, bases[:64] * 2)
    assert two.commit(R.to_mont_bytes(12345, sm) + R.to_mont_bytes(sm - 12345, sm)) == IDENTITY
    with pytest.raises(TooLongInput) as ei:
        two.commit(bytes(32 * 3))
    assert ei.value.input_len == 3 and ei.value.limit == 2
    assert two.commit(b"") == IDENTITY


@pytest.mark.parametrize("curve,logn", [(R.BN254, 14), (R.GRUMPKIN, 14), (R.BN254, 17), (R.GRUMPKIN, 16)])
def test_commit_medium_vs_oracle(gpu, curve, logn):
    from mira_b200 import CommitmentKey
    n = (1 << logn) + 17          # ragged, not a power of two
    bases_dev = gpu.gen_bases_dev(curve, 555, n)
    bases = gpu.to_bytes(bases_dev)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    ck.check_on_curve()
    for dist in (0, 1):
        sc = O.gen_scalars(curve, 556 + dist, n, dist)
        assert ck.commit(sc) == O.commit(curve, bases, sc)


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_on_curve_check_rejects_bad_key(gpu, curve):
    from mira_b200 import CommitmentKey, NotOnCurve
    bases = bytearray(O.gen_bases(curve, 9, 300))
    CommitmentKey(curve, bytes(bases)).check_on_curve()
    bases[64 * 123 + 5] ^= 1
    with pytest.raises(NotOnCurve):
        CommitmentKey(curve, bytes(bases)).check_on_curve()


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_sharded_partials_combine_to_full_commit(gpu, curve):
    """SURVEY.md §8e on one GPU: 4 point-range shards -> 4 XYZZ partials -> combine == full commit."""
    from mira_b200 import CommitmentKey, combine_partials
    n, shards = 10000, 4
    bases = O.gen_bases(curve, 71, n)
    sc = O.gen_scalars(curve, 72, n)
    want = O.commit(curve, bases, sc)
    parts = b""
    for g in range(shards):
        lo, hi = g * n // shards, (g + 1) * n // shards
        ck = CommitmentKey(curve, bases[64 * lo:64 * hi])
        parts += ck.partial(sc[32 * lo:32 * hi])
    assert combine_partials(curve, parts) == want
    assert combine_partials(curve, bytes(128) * 3) == bytes(64)


def test_homomorphism_at_2_20(gpu):
    """Size-independent property the reference's own tests rely on (is_sat_relaxed,
    src/plonk/mod.rs:547-557): commit(a + r*b) == commit(a) + r*commit(b), at a size the oracle
    would need minutes for; plus one oracle cross-check of commit(a) at 2^18."""
    from mira_b200 import CommitmentKey
    curve = R.BN254
    sm = R.scalar_mod(curve)
    n = 1 << 20
    bases_dev = gpu.gen_bases_dev(curve, 2020, n)
    ck = CommitmentKey(curve, bases_dev, on_device=True)
    a = O.gen_scalars(curve, 1, n, 1)
    b = O.gen_scalars(curve, 2, n, 0)
    r = R.gen_scalar(curve, 3, 0)
    rb = R.to_mont_bytes(r, sm)
    # a + r*b element-wise with the oracle's field ops (vectorised through fe_mul_many)
    rvec = rb * n
    rbv = O.fe_mul_many(R.FR, rvec, b)
    import numpy as np
    av = np.frombuffer(a, dtype="<u8").reshape(n, 4)
    # add via python ints would be slow; use the GPU-independent oracle adds per element in chunks
    folded = bytearray(32 * n)
    L = O.lib()
    import ctypes as C
    buf = C.create_string_buffer(32)
    for i in range(0, n, 1):
        L.oracle_fe_add(R.FR, a[32 * i:32 * i + 32], rbv[32 * i:32 * i + 32], buf)
        folded[32 * i:32 * i + 32] = buf.raw
    ca, cb, cf = ck.commit(a), ck.commit(b), ck.commit(bytes(folded))
    assert cf == O.point_add(curve, ca, O.scalar_mul(curve, cb, rb))
    k = 1 << 18
    assert ck.commit(a[: 32 * k]) == O.commit(curve, gpu.to_bytes(bases_dev[: 64 * k]), a[: 32 * k])


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_shim_startup_selfcheck(gpu, curve):
    """The three assertions INTEGRATION.md §3 asks the Rust shim to run once per curve (layout pinning)."""
    from mira_b200 import CommitmentKey
    m = R.scalar_mod(curve)
    g = O.generator(curve)
    one, zero, minus_one = R.to_mont_bytes(1, m), bytes(32), R.to_mont_bytes(m - 1, m)
    assert CommitmentKey(curve, g).commit(one) == g
    assert CommitmentKey(curve, g).commit(zero) == bytes(64)
    assert CommitmentKey(curve, g + g).commit(minus_one + one) == bytes(64)


def test_golden_commit_vectors(gpu):
    """tests/golden/commit_vectors.json (oracle-generated in the build container): same bytes on the GPU."""
    import hashlib
    import json
    import os
    from mira_b200 import CommitmentKey
    with open(os.path.join(os.path.dirname(__file__), "golden", "commit_vectors.json")) as f:
        vecs = json.load(f)
    for v in vecs:
        bases = gpu.to_bytes(gpu.gen_bases_dev(v["curve"], v["seed_bases"], v["n"]))
        sc = gpu.to_bytes(gpu.gen_scalars_dev(v["curve"], v["seed_scalars"], v["n"], v["dist"]))
        assert hashlib.sha256(bases).hexdigest() == v["bases_sha256"]
        assert hashlib.sha256(sc).hexdigest() == v["scalars_sha256"]
        assert CommitmentKey(v["curve"], bases).commit(sc).hex() == v["commit_hex"]


def test_key_file_roundtrip(gpu, tmp_path):
    """file_tests::consistency (src/commitment.rs:178-194): setup -> save -> load -> same key; plus the
    on-curve check of load_or_setup_cache (src/commitment.rs:134-156)."""
    from mira_b200 import BN254_G1, CommitmentKey, NotOnCurve
    k = 10
    key = CommitmentKey.load_or_setup_cache(str(tmp_path), "bn256", k, BN254_G1)
    path = tmp_path / "bn256" / f"{k}.bin"
    data = path.read_bytes()
    assert len(data) == 64 << k and all(O.is_on_curve(R.BN254, data[i:i + 64]) for i in range(0, 64 * 32, 64))
    again = CommitmentKey.load_or_setup_cache(str(tmp_path), "bn256", k, BN254_G1)
    sc = O.gen_scalars(R.BN254, 99, 1 << k)
    assert key.commit(sc) == again.commit(sc) == O.commit(R.BN254, data, sc)
    bad = bytearray(data)
    bad[64 * 5] ^= 1
    path.write_bytes(bytes(bad))
    with pytest.raises(NotOnCurve):
        CommitmentKey.load_or_setup_cache(str(tmp_path), "bn256", k, BN254_G1)


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
@pytest.mark.parametrize("dist", [0, 1])
def test_sliced_host_commit_equals_unsliced(gpu, curve, dist):
    """Host-buffer commits fold the scalar vector into the buckets slice by slice (H2D overlap); the result must
    not depend on the slicing.  A 40-scalar minimum forces 3-4 ragged, geometrically growing slices at these sizes."""
    from mira_b200 import CommitmentKey, combine_partials
    for n in (999, 4001, 70_001):
        bases = O.gen_bases(curve, 21, n)
        sc = O.gen_scalars(curve, 22 + n, n, dist)
        ck = CommitmentKey(curve, bases)
        ck.set_slice_min(0)
        whole = ck.commit(sc)
        for c in (0, 6, 13):
            ck.set_window(c)
            ck.set_slice_min(40)
            assert ck.commit(sc) == whole
            assert combine_partials(curve, ck.partial(sc)) == whole
        ck.set_window(0)
        assert whole == O.commit(curve, bases, sc)
    # one heavy bucket: every scalar equal => a single run spanning every chunk of every slice
    n = 50_000
    bases = O.gen_bases(curve, 23, n)
    one = O.gen_scalars(curve, 5, 1, 0) * n
    ck = CommitmentKey(curve, bases)
    ck.set_slice_min(500)
    assert ck.commit(one) == O.commit(curve, bases, one)


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_batched_commit_equals_individual_commits(gpu, curve):
    """mira_msm_commit_batch == `vs.iter().map(|v| ck.commit(v))` (src/nifs/vanilla/mod.rs:124-127), bit for bit:
    the vectors share one sort / accumulation / reduction but each keeps its own bucket set."""
    from mira_b200 import CommitmentKey, TooLongInput
    n = 5000
    bases = O.gen_bases(curve, 31, n)
    ck = CommitmentKey(curve, bases)
    vecs = [O.gen_scalars(curve, 40 + k, n, k % 2) for k in range(6)]
    vecs[3] = bytes(32 * n)                                        # an all-zero cross term (the `None` arm, :118)
    vecs[4] = O.gen_scalars(curve, 5, 1, 0) * n                    # one heavy bucket per window
    want = [O.commit(curve, bases, v) for v in vecs]
    assert want[3] == bytes(64)
    devs = [torch.frombuffer(bytearray(v), dtype=torch.uint8).cuda() for v in vecs]
    for c in (0, 7, 12):
        ck.set_window(c)
        for count in (1, 2, 6):
            got = ck.commit_batch_device([d.data_ptr() for d in devs[:count]], n)
            assert got == want[:count]
    ck.set_window(0)
    assert ck.commit_batch_device([], n) == []
    m = 1237                                                       # prefix of the key, ragged length
    got = ck.commit_batch_device([d.data_ptr() for d in devs[:3]], m)
    assert got == [O.commit(curve, bases[:64 * m], v[:32 * m]) for v in vecs[:3]]
    with pytest.raises(TooLongInput):
        ck.commit_batch_device([devs[0].data_ptr()], n + 1)
    assert ck.commit(vecs[0]) == want[0]                           # the single-vector path still works afterwards


def test_device_copy_of_host_scalars_is_reusable(gpu):
    """mira_msm_scalars_device: the device copy a host-buffer commit made (sliced or staged from pageable memory) holds
    exactly the caller's bytes, so evaluation and fold can reuse it without a second H2D."""
    from mira_b200 import BN254_G1, CommitmentKey
    n = 30_011
    bases = O.gen_bases(R.BN254, 51, n)
    sc = O.gen_scalars(R.BN254, 52, n, 1)
    ck = CommitmentKey(BN254_G1, bases)
    assert ck.scalars_device() is None
    ck.set_slice_min(500)
    want = ck.commit(sc)                                   # pageable bytes object: staged path, 4 slices
    view = ck.scalars_device()
    assert gpu.to_bytes(view) == sc
    assert ck.commit_device(view.data_ptr(), n) == want == O.commit(R.BN254, bases, sc)
    pinned = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()
    assert ck.commit(pinned) == want and gpu.to_bytes(ck.scalars_device()) == sc


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_adaptive_window_only_changes_speed(gpu, curve):
    """Commits of >= 2^18 scalars pick their window from a sample of the scalars (witness-like vectors get a much
    narrower one than uniform vectors); the commitment must not depend on it, from device or host memory."""
    from mira_b200 import CommitmentKey
    n = (1 << 18) + 5
    bases = gpu.gen_bases_dev(curve, 61, n)
    ck = CommitmentKey(curve, bases, on_device=True)
    host_bases = gpu.to_bytes(bases)
    seen = {}
    for dist in (0, 1):
        sc = gpu.gen_scalars_dev(curve, 62 + dist, n, dist)
        host_sc = gpu.to_bytes(sc)
        want = O.commit(curve, host_bases, host_sc)
        ck.set_adaptive_window(True)
        assert ck.commit_device(sc.data_ptr(), n) == want
        seen[dist] = ck.stats()["window_bits"]
        assert ck.commit(host_sc) == want                      # host path samples through 64 small copies
        ck.set_adaptive_window(False)
        assert ck.commit_device(sc.data_ptr(), n) == want
        default_c = ck.stats()["window_bits"]
    assert seen[1] < seen[0] and seen[0] == default_c          # sparse -> narrower window; uniform -> the size heuristic's


def test_randomized_commit_configurations(gpu):
    """Seeded random sweep over (curve, length, distribution, window, slicing, adaptive, source) — every combination must
    give the oracle's bytes.  60 configurations, sizes up to 40,000."""
    from mira_b200 import CommitmentKey
    rng = random.Random(int(os.environ.get("MIRA_SWEEP_SEED", "20261018")))
    keys = {}
    for it in range(int(os.environ.get("MIRA_SWEEP_CONFIGS", "60"))):      # raise for a soak run
        curve = rng.choice([R.BN254, R.GRUMPKIN])
        n_key = rng.choice([1, 2, 33, 1000, 9973, 40_000])
        if (curve, n_key) not in keys:
            b = O.gen_bases(curve, 1000 + n_key, n_key)
            keys[(curve, n_key)] = (b, CommitmentKey(curve, b))
        bases, ck = keys[(curve, n_key)]
        n = rng.choice([0, 1, n_key // 2, max(n_key - 1, 0), n_key])
        dist = rng.choice([0, 1])
        sc = O.gen_scalars(curve, 5000 + it, n, dist)
        if n and rng.random() < 0.2:                                  # a block of identical scalars: heavy buckets
            one = O.gen_scalars(curve, 7, 1)
            k = rng.randrange(1, n + 1)
            sc = one * k + sc[32 * k:]
        ck.set_window(rng.choice([0, 0, 0, 2, 3, 7, 11, 16, 21]))
        ck.set_slice_min(rng.choice([0, 1, 50, 1 << 19]))
        ck.set_adaptive_window(rng.random() < 0.5)
        ck.set_affine_levels(rng.choice([0, 0, 1, 2, 4, 6]))
        ck.set_pipeline(rng.choice([0, 1, 2, 3, 4, 7, 16]), rng.choice([0, 1, 100, 5000]))
        want = O.commit(curve, bases, sc)
        src = rng.choice(["bytes", "pinned", "device"])
        if src == "bytes":
            got = ck.commit(sc)
        elif src == "pinned":
            got = ck.commit(torch.frombuffer(bytearray(sc) if sc else bytearray(32), dtype=torch.uint8).pin_memory()[: len(sc)])
        else:
            d = torch.frombuffer(bytearray(sc) if sc else bytearray(32), dtype=torch.uint8).cuda()
            got = ck.commit_device(d.data_ptr(), n)
        assert got == want, (it, curve, n_key, n, dist, src, ck.stats())


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_batched_affine_levels_only_change_speed(gpu, curve):
    """The experimental affine pre-reduction (mira_msm_set_affine_levels): every level count gives the oracle's bytes on
    multi-tile lists (131 k points, ~1.7 M pairs), uniform and witness-like scalars (one bucket holding a third of all
    pairs), a block of identical scalars on identical points (tangent additions at every level), through the sliced
    host path and the batched commit."""
    from mira_b200 import CommitmentKey
    n = (1 << 17) + 11
    bases_dev = gpu.gen_bases_dev(curve, 8181, n)
    b = bytearray(gpu.to_bytes(bases_dev))
    for i in range(5000, 5400):
        b[64 * i:64 * i + 64] = b[64 * 4999:64 * 5000]            # 401 copies of one point
    b[64 * 9:64 * 10] = bytes(64)                                  # an identity generator
    b[64 * 11:64 * 12] = O.point_neg(curve, bytes(b[64 * 10:64 * 11]))
    bases = bytes(b)
    ck = CommitmentKey(curve, bases)
    one = O.gen_scalars(curve, 8182, 1)
    vectors = []
    for dist in (0, 1):
        sc = bytearray(O.gen_scalars(curve, 8183 + dist, n, dist))
        sc[32 * 4999:32 * 5400] = one * 401                        # same scalar on the copies: P + P in every window
        sc[32 * 11:32 * 12] = sc[32 * 10:32 * 11]                  # s*P + s*(-P): cancellation inside a bucket
        vectors.append(bytes(sc))
    want = [O.commit(curve, bases, v) for v in vectors]
    for levels in (0, 1, 3, 6, CommitmentKey.AFFINE_THREAD_LOCAL_PAIRS):
        ck.set_affine_levels(levels)
        for c in (0, 9):
            ck.set_window(c)
            ck.set_slice_min(0)
            assert [ck.commit(v) for v in vectors] == want, (levels, c)
            ck.set_slice_min(20_000)                               # four slices folded into live buckets
            assert ck.commit(vectors[0]) == want[0], (levels, c, "sliced")
        d = [torch.frombuffer(bytearray(v), dtype=torch.uint8).cuda() for v in vectors]
        assert ck.commit_batch_device([t.data_ptr() for t in d], n) == want, (levels, "batch")


def test_contexts_release_their_device_memory(gpu):
    """Create / use / destroy in a loop: the free device memory must come back (every workspace, the fixed-base tables
    and the optional affine-level buffers are owned by the context)."""
    from mira_b200 import CommitmentKey
    curve, n = R.BN254, 20_000
    bases = O.gen_bases(curve, 91, n)
    sc = O.gen_scalars(curve, 92, n)
    want = O.commit(curve, bases, sc)

    def once(levels):
        ck = CommitmentKey(curve, bases)
        ck.set_affine_levels(levels)
        assert ck.commit(sc) == want
        d = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        assert ck.commit_batch_device([d.data_ptr(), d.data_ptr()], n) == [want, want]
        ck.close()

    once(2)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free0, _ = torch.cuda.mem_get_info()
    for it in range(12):
        once(it % 3)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 32 << 20, f"{(free0 - free1) >> 20} MiB not returned after 12 create/destroy cycles"


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_reference_digest_kat_on_the_gpu(gpu, curve):
    """The reference's own curve-level known answer, `digest::tests::consistency` (src/digest.rs:99-114):
    (r - 1) * G == -G for bn256, written with literal bytes — G = (1, 2), so -G = (1, p - 2) — and no oracle in the loop.
    Run through every commit route that can differ (windows, device / host scalars, batch, affine levels).  The Grumpkin
    line is the same identity on the cycle's other curve (generator (1, sqrt(-16)) taken from the oracle)."""
    from mira_b200 import CommitmentKey
    pm, sm = R.base_mod(curve), R.scalar_mod(curve)
    if curve == R.BN254:
        g = R.to_mont_bytes(1, pm) + R.to_mont_bytes(2, pm)
        neg_g = R.to_mont_bytes(1, pm) + R.to_mont_bytes(pm - 2, pm)
    else:
        g = O.generator(curve)
        neg_g = O.point_neg(curve, g)
    s = R.to_mont_bytes(sm - 1, sm)
    ck = CommitmentKey(curve, g)
    for c in (0, 2, 7, 16, 22):
        ck.set_window(c)
        for levels in (0, 2):
            ck.set_affine_levels(levels)
            assert ck.commit(s) == neg_g
            d = torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda()
            assert ck.commit_device(d.data_ptr(), 1) == neg_g
            assert ck.commit_batch_device([d.data_ptr(), d.data_ptr()], 1) == [neg_g, neg_g]
    # and on a longer key: s * G + 1 * G == identity
    two = CommitmentKey(curve, g + g)
    assert two.commit(s + R.to_mont_bytes(1, sm)) == bytes(64)


def test_reference_lagrange_kat_on_the_gpu(gpu):
    """`lagrange::tests` known answers (src/polynomial/lagrange.rs:113-126, committed as tests/golden/lagrange_kat_fr.json):
    L_i(2) on the 2^2 domain, computed with the DEVICE field kernels only (multiplication, subtraction, inversion,
    Montgomery conversions) from integer literals — no oracle in the loop."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagrange_kat_fr.json")) as f:
        want = [int(v) for v in json.load(f)["raw_strings"]]
    MUL, SUB, INV, TO_CANON, FROM_CANON = 0, 2, 4, 5, 6
    fr = lambda v: gpu.field_op(R.FR, FROM_CANON, (v % R.R_).to_bytes(32, "little"), bytes(32))
    op = lambda o, a, b=None: gpu.field_op(R.FR, o, a, b if b is not None else bytes(32))
    root = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c     # halo2curves Fr::ROOT_OF_UNITY
    w = fr(root)
    for _ in range(2, 28):                        # get_omega_or_inv (src/fft.rs:12-24): omega of the 2^2 domain
        w = op(MUL, w, w)
    n, X = 4, fr(2)
    xn = fr(1)
    for _ in range(n):
        xn = op(MUL, xn, X)
    num, ninv = op(SUB, xn, fr(1)), op(INV, fr(n))
    wi, got = fr(1), []
    for _ in range(n):
        den = op(INV, op(SUB, X, wi))
        v = op(MUL, op(MUL, wi, ninv), op(MUL, num, den))
        got.append(int.from_bytes(op(TO_CANON, v), "little"))
        wi = op(MUL, wi, w)
    assert got == want


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_pipelined_commit_equals_unpipelined(gpu, curve):
    """The slice pipeline (digits + sort of part k+1 on a second stream while part k is accumulated) only changes speed:
    every slice count gives the oracle's bytes, for device and for page-locked host scalars, uniform and witness-like,
    back to back on the same context (buffer sets and events are reused), and through the asynchronous partial path."""
    from mira_b200 import CommitmentKey, combine_partials_device
    n = 70_001
    bases = O.gen_bases(curve, 661, n)
    ck = CommitmentKey(curve, bases)
    for dist in (0, 1):
        sc = O.gen_scalars(curve, 662 + dist, n, dist)
        want = O.commit(curve, bases, sc)
        d = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        pinned = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()
        for slices, min_slice in ((1, 0), (2, 1), (4, 1000), (16, 1), (5, 3000), (0, 0)):
            ck.set_pipeline(slices, min_slice)
            ck.set_slice_min(4000)                       # host path: H2D slices, pipelined as well
            for _ in range(2):
                assert ck.commit_device(d.data_ptr(), n) == want, (slices, min_slice, dist)
                assert ck.commit(pinned) == want, (slices, min_slice, dist, "pinned")
            assert ck.commit(sc) == want                 # pageable source: staged, not pipelined
            out = torch.zeros(128, dtype=torch.uint8, device="cuda")
            ck.partial_batch_device([d.data_ptr()], n, out.data_ptr())
            ck.partial_batch_device([d.data_ptr()], n, out.data_ptr())     # queued twice without a host sync in between
            torch.cuda.synchronize()
            assert combine_partials_device(curve, out.data_ptr(), 1, 1, 128) == [want]
    # more slices than 256-scalar boundaries: parts that round to nothing are dropped (found by the soak run)
    ck.set_pipeline(16, 1)
    for m in (1, 16, 300, 4097):
        sc = O.gen_scalars(curve, 670 + m, m)
        d = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
        assert ck.commit_device(d.data_ptr(), m) == O.commit(curve, bases, sc), m


def test_msd_partition_on_small_and_edge_inputs(gpu):
    """The MSD partition (k_msd_* in msm_kernels.cuh) is used from 2^25 pairs on, so in this suite only the 2^22-point
    tests reach it by size.  MIRA_SORT_MSD_MIN_LOG=0 (read once per process, hence the child process) removes the size
    gate: the edge-case keys (identity / duplicate / negated generators, one repeated point), every window width from 17
    bits up, the batched commit with its per-set bucket ranges, sliced host commits and the randomized sweep then all go
    through it and must still give the oracle's bytes."""
    import subprocess
    import sys
    env = dict(os.environ, MIRA_SORT_MSD_MIN_LOG="0", MIRA_SWEEP_CONFIGS="24")
    here = os.path.abspath(__file__)
    sel = ["test_commit_edge_cases", "test_commit_all_window_widths", "test_batched_affine_levels_only_change_speed",
           "test_randomized_commit_configurations"]
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-x", "-q", "-m", "gpu", "-k", " or ".join(sel)],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(here)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
