"""Independent Python big-integer model of the BN254 / Grumpkin arithmetic.

Second opinion for the C oracle (oracle/mira_oracle.c) and the CUDA path: evaluates
sum(s_i * P_i) with affine formulas over Python ints — no Montgomery form, no windows,
no shared code with either implementation.  Only used by tests (small sizes).

Layouts follow the reference boundary (src/commitment.rs:78-87):
field element = 32 bytes LE of (value * 2^256 mod m); affine point = x || y; identity = 64 zero bytes.
"""
from __future__ import annotations

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # BN254 Fq
R_ = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # BN254 Fr
MONT_R = 1 << 256

BN254, GRUMPKIN = 0, 1
FQ, FR = 0, 1

FIELD_MOD = {FQ: P, FR: R_}
# curve id -> (base modulus, scalar modulus, b, generator)
CURVES = {
    BN254: (P, R_, 3, (1, 2)),
    GRUMPKIN: (R_, P, (-17) % R_, (1, 0x2CF135E7506A45D632D270D45F1181294833FC48D823F272C)),
}


def base_mod(curve): return CURVES[curve][0]
def scalar_mod(curve): return CURVES[curve][1]


def to_mont_bytes(v: int, m: int) -> bytes:
    return ((v % m) * MONT_R % m).to_bytes(32, "little")


def from_mont_bytes(b: bytes, m: int) -> int:
    return int.from_bytes(b, "little") * pow(MONT_R, -1, m) % m


def point_to_bytes(pt, curve) -> bytes:
    m = base_mod(curve)
    if pt is None:
        return bytes(64)
    return to_mont_bytes(pt[0], m) + to_mont_bytes(pt[1], m)


def point_from_bytes(b: bytes, curve):
    m = base_mod(curve)
    if b == bytes(64):
        return None
    return (from_mont_bytes(b[:32], m), from_mont_bytes(b[32:64], m))


def is_on_curve(pt, curve) -> bool:
    if pt is None:
        return True
    m, _, b, _ = CURVES[curve]
    x, y = pt
    return (y * y - x * x * x - b) % m == 0


def add(p1, p2, curve):
    m = base_mod(curve)
    if p1 is None: return p2
    if p2 is None: return p1
    x1, y1 = p1; x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % m == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, m) % m
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, m) % m
    x3 = (lam * lam - x1 - x2) % m
    return (x3, (lam * (x1 - x3) - y1) % m)


def neg(p, curve):
    if p is None: return None
    return (p[0], (-p[1]) % base_mod(curve))


def mul(k: int, p, curve):
    k %= scalar_mod(curve)
    acc = None
    while k:
        if k & 1:
            acc = add(acc, p, curve)
        p = add(p, p, curve)
        k >>= 1
    return acc


def msm(scalars, points, curve):
    acc = None
    for s, p in zip(scalars, points):
        acc = add(acc, mul(s, p, curve), curve)
    return acc


def commit_bytes(curve, bases: bytes, scalars: bytes) -> bytes:
    """CommitmentKey::commit on raw reference-layout bytes; uses the key prefix."""
    n = len(scalars) // 32
    assert len(bases) // 64 >= n
    sm = scalar_mod(curve)
    ss = [from_mont_bytes(scalars[32 * i:32 * i + 32], sm) for i in range(n)]
    ps = [point_from_bytes(bases[64 * i:64 * i + 64], curve) for i in range(n)]
    return point_to_bytes(msm(ss, ps, curve), curve)


# ---- synthetic input definition (same as oracle_gen_scalars / mira testgen) ----
MASK64 = (1 << 64) - 1


def sm64_word(seed: int, k: int) -> int:
    z = (seed + (k + 1) * 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def gen_canon(m: int, seed: int, i: int) -> int:
    l = [sm64_word(seed, 4 * i + k) for k in range(4)]
    l[3] &= 0x3FFFFFFFFFFFFFFF
    v = l[0] | (l[1] << 64) | (l[2] << 128) | (l[3] << 192)
    return v - m if v >= m else v


def gen_scalar(curve, seed, i, dist=0) -> int:
    v = gen_canon(scalar_mod(curve), seed, i)
    if dist == 1:
        sel = sm64_word(seed ^ 0x5EED, i) % 100
        if sel < 60: v = 0
        elif sel < 85: v &= 1
        elif sel < 95: v &= 0xFFFFFFFF
    return v
