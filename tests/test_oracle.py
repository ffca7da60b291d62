"""Pins the CPU oracle (oracle/) against every known-answer test the reference holds for the
field/curve layer, and against an independent Python big-int model.  CPU only."""
import random

import pytest

import oracle_lib as O
import pyref as R

FR_KAT_FFT = [  # /root/reference/src/fft.rs:239-258 : fft([0..7], k=3) over BN254 Fr
    28,
    68918385373930674424918168212551896122229959265833979749191472831399925654,
    17631683881184975370165255887551781615748388533673675138856,
    68918385373930639161550405842601155791718184162270748252414405484049647934,
    21888242871839275222246405745257275088548364400416034343698204186575808495613,
    21819324486465344583084855339414673932756646216253763595445789781091758847675,
    21888242871839275204614721864072299718383108512864252727949815652902133356753,
    21819324486465344547821487577044723192426134441150200363949012713744408569955,
]
LAGRANGE_KAT = [  # /root/reference/src/polynomial/lagrange.rs:113-126 : L_i(2), domain 2^2
    5472060717959818805561601436314318772137091100104008585924551046643952123908,
    5472060717959818798949719980869953008325120142272090480018905346516323946831,
    5472060717959818805561601436314318772137091100104008585924551046643952123903,
    5472060717959818812173482891758684535949062057935926691830196746771580300976,
]
ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C  # halo2curves Fr, S = 28


def fr(v):
    return O.fe_from_canonical(R.FR, (v % R.R_).to_bytes(32, "little"))


def fr_int(b):
    return int.from_bytes(O.fe_to_canonical(R.FR, b), "little")


def omega(k):
    w = fr(ROOT_OF_UNITY)
    for _ in range(k, 28):  # fft.rs:12-24 get_omega_or_inv
        w = O.fe_op("mul", R.FR, w, w)
    return w


def oracle_fft(vals, k):
    """radix-2 DIT as src/fft.rs:51-115 computes it, using only oracle field ops."""
    n = 1 << k
    a = list(vals)
    for i in range(n):
        j = int(format(i, f"0{k}b")[::-1], 2)
        if i < j:
            a[i], a[j] = a[j], a[i]
    w_n = omega(k)
    m = 1
    for _ in range(k):
        w_m = w_n
        for _ in range(k - 1 - (m.bit_length() - 1)):
            w_m = O.fe_op("mul", R.FR, w_m, w_m)
        for s in range(0, n, 2 * m):
            w = fr(1)
            for t in range(m):
                u = a[s + t]
                v = O.fe_op("mul", R.FR, a[s + t + m], w)
                a[s + t] = O.fe_op("add", R.FR, u, v)
                a[s + t + m] = O.fe_op("sub", R.FR, u, v)
                w = O.fe_op("mul", R.FR, w, w_m)
        m *= 2
    return a


def test_fft_kat_pins_fr_arithmetic():
    out = oracle_fft([fr(i) for i in range(8)], 3)
    assert [fr_int(x) for x in out] == FR_KAT_FFT


def test_lagrange_kat_pins_fr_inversion():
    # L_i(X) = (w^i / n) * (X^n - 1) / (X - w^i)  on the 2^2 domain, X = 2
    k, n, X = 2, 4, fr(2)
    w = omega(k)
    xn = fr(1)
    for _ in range(n):
        xn = O.fe_op("mul", R.FR, xn, X)
    num = O.fe_op("sub", R.FR, xn, fr(1))
    ninv = O.fe_inv(R.FR, fr(n))
    wi = fr(1)
    got = []
    for _ in range(n):
        den = O.fe_inv(R.FR, O.fe_op("sub", R.FR, X, wi))
        v = O.fe_op("mul", R.FR, O.fe_op("mul", R.FR, wi, ninv), O.fe_op("mul", R.FR, num, den))
        got.append(fr_int(v))
        wi = O.fe_op("mul", R.FR, wi, w)
    assert got == LAGRANGE_KAT


def test_digest_kat_generator_times_r_minus_1():
    # /root/reference/src/digest.rs:99-114 : (r-1) * G == -G on bn256
    g = O.generator(R.BN254)
    s = fr(R.R_ - 1)
    assert O.scalar_mul(R.BN254, g, s) == O.point_neg(R.BN254, g)


@pytest.mark.parametrize("field", [R.FQ, R.FR])
def test_field_ops_vs_python(field):
    m = R.FIELD_MOD[field]
    rng = random.Random(1234 + field)
    edge = [0, 1, 2, m - 1, m - 2, (1 << 253), (1 << 128) - 1]
    vals = edge + [rng.randrange(m) for _ in range(200)]
    for a in vals[:40]:
        for b in vals[:40]:
            A, B = R.to_mont_bytes(a, m), R.to_mont_bytes(b, m)
            assert O.fe_op("mul", field, A, B) == R.to_mont_bytes(a * b, m)
            assert O.fe_op("add", field, A, B) == R.to_mont_bytes(a + b, m)
            assert O.fe_op("sub", field, A, B) == R.to_mont_bytes(a - b, m)
    for a in vals:
        A = R.to_mont_bytes(a, m)
        assert O.fe_to_canonical(field, A) == a.to_bytes(32, "little")
        assert O.fe_from_canonical(field, a.to_bytes(32, "little")) == A
        inv = pow(a, -1, m) if a else 0
        assert O.fe_inv(field, A) == R.to_mont_bytes(inv, m)


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_generators_and_group_law(curve):
    g = O.generator(curve)
    assert g == R.point_to_bytes(R.CURVES[curve][3], curve)
    assert O.is_on_curve(curve, g) and R.is_on_curve(R.CURVES[curve][3], curve)
    assert O.is_on_curve(curve, bytes(64))
    gp = R.CURVES[curve][3]
    rng = random.Random(99 + curve)
    sm = R.scalar_mod(curve)
    for k in [0, 1, 2, 3, sm - 1, sm - 2] + [rng.randrange(sm) for _ in range(8)]:
        got = O.scalar_mul(curve, g, R.to_mont_bytes(k, sm))
        assert got == R.point_to_bytes(R.mul(k, gp, curve), curve)
        assert O.is_on_curve(curve, got)
    # order of the group is the scalar modulus (2-cycle): sm * G = identity
    assert R.mul(sm - 1, gp, curve) == R.neg(gp, curve)
    # add: P+Q, P+P, P+(-P), P+0
    p2 = R.mul(2, gp, curve)
    assert O.point_add(curve, g, g) == R.point_to_bytes(p2, curve)
    assert O.point_add(curve, g, O.point_neg(curve, g)) == bytes(64)
    assert O.point_add(curve, g, bytes(64)) == g
    assert O.point_add(curve, bytes(64), g) == g


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_synthetic_generators_match_python_definition(curve):
    seed = 0x4D495241
    sm = R.scalar_mod(curve)
    for dist in (0, 1):
        sc = O.gen_scalars(curve, seed, 64, dist)
        for i in range(64):
            assert sc[32 * i:32 * i + 32] == R.to_mont_bytes(R.gen_scalar(curve, seed, i, dist), sm)
    # windowed `first` offset is consistent
    assert O.gen_scalars(curve, seed, 8, 0, first=56) == O.gen_scalars(curve, seed, 64, 0)[56 * 32:]
    bases = O.gen_bases(curve, seed + 1, 24)
    gp = R.CURVES[curve][3]
    for i in range(24):
        k = R.gen_canon(sm, seed + 1, i)
        assert bases[64 * i:64 * i + 64] == R.point_to_bytes(R.mul(k, gp, curve), curve)
    assert O.gen_bases(curve, seed + 1, 5, first=19) == bases[19 * 64:]


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 31, 32, 33, 100, 257])
def test_commit_vs_python_bigint(curve, n):
    seed = 7 + n
    bases = O.gen_bases(curve, seed, max(n, 1) + 3)
    scalars = O.gen_scalars(curve, seed + 1000, n)
    want = R.commit_bytes(curve, bases, scalars)
    for threads in (1, 3, 8):
        assert O.commit(curve, bases, scalars, threads) == want
    if n <= 33:
        assert O.commit_naive(curve, bases, scalars) == want


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_commit_edge_cases(curve):
    sm = R.scalar_mod(curve)
    seed = 4242
    n = 48
    bases = bytearray(O.gen_bases(curve, seed, n))
    # identity bases, duplicated bases, a base and its negation
    bases[64 * 3:64 * 4] = bytes(64)
    bases[64 * 5:64 * 6] = bases[64 * 4:64 * 5]
    bases[64 * 7:64 * 8] = O.point_neg(curve, bytes(bases[64 * 6:64 * 7]))
    bases = bytes(bases)
    vals = [R.gen_scalar(curve, seed + 1, i) for i in range(n)]
    vals[0] = 0; vals[1] = 1; vals[2] = sm - 1; vals[4] = vals[5] = 5; vals[6] = vals[7] = 9
    scalars = b"".join(R.to_mont_bytes(v, sm) for v in vals)
    want = R.commit_bytes(curve, bases, scalars)
    assert O.commit(curve, bases, scalars, 4) == want
    # all-zero scalars -> identity (0,0); src/poseidon/poseidon_hash.rs:129-143 consumes it as (0,0)
    assert O.commit(curve, bases, bytes(32 * n), 4) == bytes(64)
    # cancelling sum: s*P + (m-s)*P = identity
    two = bases[:64] * 2
    sc = R.to_mont_bytes(12345, sm) + R.to_mont_bytes(sm - 12345, sm)
    assert O.commit(curve, two, sc, 1) == bytes(64)
    # TooLongInput (src/commitment.rs:82-85)
    with pytest.raises(O.TooLongInput):
        O.commit(curve, bases[:64 * 4], scalars[:32 * 5])


@pytest.mark.parametrize("curve", [R.BN254, R.GRUMPKIN])
def test_commit_is_homomorphism(curve):
    """What is_sat_relaxed pins (src/plonk/mod.rs:547-557): commit(a + r*b) = commit(a) + r*commit(b)."""
    n = 1000
    sm = R.scalar_mod(curve)
    bases = O.gen_bases(curve, 31337, n)
    a = [R.gen_scalar(curve, 1, i, 1) for i in range(n)]
    b = [R.gen_scalar(curve, 2, i, 0) for i in range(n)]
    r = R.gen_scalar(curve, 3, 0)
    enc = lambda v: b"".join(R.to_mont_bytes(x, sm) for x in v)
    ca = O.commit(curve, bases, enc(a))
    cb = O.commit(curve, bases, enc(b))
    cf = O.commit(curve, bases, enc([(x + r * y) % sm for x, y in zip(a, b)]))
    rcb = O.scalar_mul(curve, cb, R.to_mont_bytes(r, sm))
    assert cf == O.point_add(curve, ca, rcb)
    # prefix-of-key semantics: a shorter vector uses ck[..len]
    assert O.commit(curve, bases, enc(a[:100])) == O.commit(curve, bases[:64 * 100], enc(a[:100]))


def test_golden_commit_vectors_freeze_the_oracle():
    """tests/golden/commit_vectors.json (ORACLE-GENERATED, see tests/golden/make_golden.py): the oracle and its
    input generators still produce the committed bytes."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "commit_vectors.json")) as f:
        vecs = json.load(f)
    for v in vecs:
        bases = O.gen_bases(v["curve"], v["seed_bases"], v["n"])
        sc = O.gen_scalars(v["curve"], v["seed_scalars"], v["n"], v["dist"])
        assert hashlib.sha256(bases).hexdigest() == v["bases_sha256"]
        assert hashlib.sha256(sc).hexdigest() == v["scalars_sha256"]
        assert O.commit(v["curve"], bases, sc).hex() == v["commit_hex"]
        if v["n"] <= 64:
            assert O.commit_naive(v["curve"], bases, sc).hex() == v["commit_hex"]


# ---- external pin of the curve layer: Ethereum's alt_bn128 (= BN254 G1) precompile vectors, EIP-196 ----------------
def _eip196():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "eip196_alt_bn128.json")) as f:
        return json.load(f)


def _pt(xy):
    return R.point_to_bytes((int(xy[0], 16), int(xy[1], 16)), R.BN254)


def test_eip196_vectors_are_self_consistent():
    """The recorded vectors against an independent chord-and-tangent computation on Python integers (no oracle code):
    guards the fixture itself against transcription errors."""
    v = _eip196()
    for c in v["ec_add"]:
        a, b, s = (tuple(int(t, 16) for t in c[k]) for k in ("a", "b", "sum"))
        assert R.is_on_curve(a, R.BN254) and R.is_on_curve(b, R.BN254) and R.add(a, b, R.BN254) == s
    for c in v["ec_mul"]:
        p, s = (tuple(int(t, 16) for t in c[k]) for k in ("p", "product"))
        assert R.is_on_curve(p, R.BN254) and R.mul(int(c["k"], 16), p, R.BN254) == s


def test_oracle_matches_eip196_ecadd_ecmul():
    """oracle point_add / scalar_mul / commit (the restated best_multiexp + to_affine, i.e. what CommitmentKey::commit
    returns, src/commitment.rs:78-87) reproduce the public alt_bn128 vectors: a pin of the oracle's G1 arithmetic,
    Montgomery layout included, that does not come from this repository."""
    v = _eip196()
    one = R.to_mont_bytes(1, R.R_)
    for c in v["ec_add"]:
        a, b, s = _pt(c["a"]), _pt(c["b"]), _pt(c["sum"])
        assert O.is_on_curve(R.BN254, a) and O.is_on_curve(R.BN254, b)
        assert O.point_add(R.BN254, a, b) == s
        assert O.commit(R.BN254, a + b, one + one) == s              # MSM with unit scalars = ecAdd
        assert O.commit_naive(R.BN254, a + b, one + one) == s
    for c in v["ec_mul"]:
        p, s, k = _pt(c["p"]), _pt(c["product"]), R.to_mont_bytes(int(c["k"], 16), R.R_)
        assert O.scalar_mul(R.BN254, p, k) == s
        assert O.commit(R.BN254, p, k) == s                          # MSM of one term = ecMul
    # both at once: k*P + 1*A + 1*B through the Pippenger path with padding points
    c_add, c_mul = v["ec_add"][1], v["ec_mul"][1]
    bases = _pt(c_mul["p"]) + _pt(c_add["a"]) + _pt(c_add["b"])
    scal = R.to_mont_bytes(int(c_mul["k"], 16), R.R_) + one + one
    want = O.point_add(R.BN254, _pt(c_mul["product"]), _pt(c_add["sum"]))
    assert O.commit(R.BN254, bases, scal) == want
    assert R.point_from_bytes(want, R.BN254) == R.add(tuple(int(t, 16) for t in c_mul["product"]),
                                                       tuple(int(t, 16) for t in c_add["sum"]), R.BN254)
