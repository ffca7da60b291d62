"""CPU tests pinning the witness-side oracle (oracle/mira_oracle_witness.c): row evaluator, fold, concat, FFT.
The evaluator tests restate the reference's own (src/polynomial/graph_evaluator.rs:447-634): evaluator output
vs direct field arithmetic (here: Python big integers)."""
import json
import os
import random

import pytest

import graph_evaluator_model as G
import oracle_lib as O
import pyref as R
from witness_util import Domain, mont, pack_program, random_expr, unmont

FR, FQ = R.FR, R.FQ
M = R.R_
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def run(expr, dom: Domain, field=FR, rows=None):
    ge = G.GraphEvaluator.new(expr, dom.m)
    out = O.eval_rows(field, pack_program(ge), dom.as_bytes())
    vals = unmont(out, dom.m)
    return vals if rows is None else [vals[r] for r in rows]


def empty_domain(m=M, row_size=1):
    # Mock::default() has no columns; row_size() would panic upstream, the tests only evaluate row 0 of constants.
    return Domain(m, row_size, 0, 1, 0, 0, 1, 0, seed=1)


# ---- graph_evaluator.rs:447-506  constant / sum_const / product_const / neg_const
def test_constant_sum_product_neg():
    rng = random.Random(1)
    d = empty_domain()
    a, b = rng.randrange(M), rng.randrange(M)
    assert run(G.Constant(a), d, rows=[0]) == [a]
    assert run(G.Constant(a) + G.Constant(b), d, rows=[0]) == [(a + b) % M]
    assert run(G.Constant(a) * G.Constant(b), d, rows=[0]) == [a * b % M]
    assert run(-G.Constant(a), d, rows=[0]) == [(-a) % M]


# ---- graph_evaluator.rs:508-572  poly: selectors / fixed / advice with rotations wrapping modulo the row count
def test_poly_rotations_and_column_dispatch():
    d = Domain(M, 2, 2, 2, 2, 0, 1, 0, seed=2)
    ns, nf = 2, 2
    for col in range(2):
        for rot in (0, 1, -1, 2):
            for row in (0, 1):
                want_adv = d.w1[0][col * 2 + (row + rot) % 2]
                assert run(G.Polynomial(ns + nf + col, rot), d, rows=[row]) == [want_adv]
                assert run(G.Polynomial(ns + col, rot), d, rows=[row]) == [d.fixed[col][(row + rot) % 2]]
                assert run(G.Polynomial(col, rot), d, rows=[row]) == [d.selectors[col][(row + rot) % 2]]


# ---- graph_evaluator.rs:574-590  challenge
def test_challenge():
    d = Domain(M, 1, 0, 1, 0, 0, 1, 3, seed=3)
    for i in range(3):
        assert run(G.Challenge(i), d, rows=[0]) == [d.challenges[i]]
    with pytest.raises(O.EvalError) as e:
        run(G.Challenge(3), d)
    assert e.value.rc == -11        # ChallengeIndexOutOfBoundary


# ---- graph_evaluator.rs:592-634  eval: (a0 + a1 + a1) * (f0 + a0)
def test_eval_example():
    d = Domain(M, 2, 2, 2, 2, 0, 1, 0, seed=4)
    adv = lambda c: G.Polynomial(4 + c)
    fx = lambda c: G.Polynomial(2 + c)
    zero = G.Constant(0)
    s1 = adv(0) + (adv(1) + (adv(1) + zero))
    s2 = fx(0) + (adv(0) + zero)
    got = run(s1 * s2, d, rows=[0])
    a00, a01, f00 = d.w1[0][0], d.w1[0][2], d.fixed[0][0]
    assert got == [(a00 + a01 + a01) * (f00 + a00) % M]


def test_invalid_witness_index_is_an_error():
    d = Domain(M, 4, 0, 1, 2, 0, 1, 0, seed=5)
    with pytest.raises(O.EvalError) as e:
        run(G.Polynomial(1 + 2 * 2), d)      # beyond both instances' advice columns
    assert e.value.rc == -14


@pytest.mark.parametrize("field,m", [(FR, R.R_), (FQ, R.P)])
@pytest.mark.parametrize("seed", range(6))
def test_random_expressions_vs_direct(field, m, seed):
    rng = random.Random(100 + seed)
    n_sel, n_fix, n_adv, n_ch = 2, 3, 4, 3
    d = Domain(m, 8, n_sel, n_fix, n_adv, 0, 1, n_ch, seed=seed, sparse=(seed % 2 == 1))
    n_cols = n_sel + n_fix + 2 * n_adv
    for _ in range(6):
        e = random_expr(rng, m, n_cols, n_ch, depth=6, rotations=(0, 1, -1, 3))
        assert run(e, d, field) == d.direct(e, range(8))


def test_lookup_column_mapping_two_and_three_rounds():
    # PlonkEvalDomain::eval_advice_var's index_map (src/plonk/eval.rs:166-203)
    for n_w in (2, 3):
        d = Domain(M, 4, 1, 1, 2, 2, n_w, 0, seed=20 + n_w)
        width = 2 + 5 * 2
        for idx in range(2 * width):
            e = G.Polynomial(2 + idx)
            assert run(e, d) == d.direct(e, range(4))


@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_cross_term_programs_fold_identity(T, n_gates):
    """sum_k X^k T_k(row) == G_hom(W1 + X*W2, ch1 + X*ch2)(row): the identity folding relies on
    (src/nifs/vanilla/mod.rs:80-140 + src/polynomial/grouped_poly.rs:88-151)."""
    progs, meta = G.cross_term_programs(T, n_gates, M)
    rows = 4
    per = meta["num_challenges"] // 2
    d = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=30 + n_gates, sparse=True)
    outs = [unmont(O.eval_rows(FR, pack_program(p), d.as_bytes()), M) for p in progs]
    for k, e in enumerate(meta["exprs"]):
        assert outs[k] == d.direct(e, range(rows))
    # T_0 = G(W1), then evaluate G on the folded inputs
    X = 0x1234567
    nf = meta["num_fixed"]
    folded = Domain(M, rows, 0, nf, meta["num_advice"], 0, 1, per, seed=0)
    folded.fixed = d.fixed
    folded.w1 = [[(a + X * b) % M for a, b in zip(d.w1[0], d.w2[0])]]
    folded.challenges = [(d.challenges[i] + X * d.challenges[per + i]) % M for i in range(per)]
    gates = [G.main_gate_expr(T, g * (T + 2), 0, nf, g * (3 * T + 3)) for g in range(n_gates)]
    expr = gates[0]
    if n_gates > 1:
        acc = G.Constant(0)
        for g in gates:
            acc = g + (acc * G.Challenge(0))
        expr = acc
    hom, deg = G.homogeneous(expr, lambda i: i >= nf, per - 1)
    assert deg == meta["degree"] == len(progs)
    lhs = folded.direct(hom, range(rows))
    t0 = d.direct(hom, range(rows))
    for r in range(rows):
        assert lhs[r] == (t0[r] + sum(pow(X, k + 1, M) * outs[k][r] for k in range(len(outs)))) % M


# ---- fold (src/plonk/mod.rs:1097-1134)
@pytest.mark.parametrize("field,m", [(FR, R.R_), (FQ, R.P)])
def test_fold_w_and_e(field, m):
    rng = random.Random(7)
    n = 300
    w1 = [rng.choice([0, 1, m - 1, rng.randrange(m)]) for _ in range(n)]
    w2 = [rng.choice([0, 1, m - 1, rng.randrange(m)]) for _ in range(n)]
    r = rng.randrange(m)
    assert unmont(O.fold_w(field, mont(w1, m), mont(w2, m), R.to_mont_bytes(r, m)), m) == [(a + r * b) % m for a, b in zip(w1, w2)]
    for n_terms in (0, 1, 5, 6):
        e = [rng.randrange(m) for _ in range(n)]
        ts = [[rng.randrange(m) for _ in range(n)] for _ in range(n_terms)]
        got = unmont(O.fold_e(field, mont(e, m), [mont(t, m) for t in ts], R.to_mont_bytes(r, m)), m)
        assert got == [(e[i] + sum(pow(r, k + 1, m) * ts[k][i] for k in range(n_terms))) % m for i in range(n)]
    assert O.fold_w(field, b"", b"", R.to_mont_bytes(r, m)) == b""


# ---- concatenate_with_padding (src/util.rs:189-193)
def test_concat_pad():
    cols = [[1, 2, 3], [], [4, 5, 6, 7, 8], [9]]
    got = unmont(O.concat_pad([mont(c, M) for c in cols], 4), M)
    assert got == [1, 2, 3, 0, 0, 0, 0, 0, 4, 5, 6, 7, 8, 9, 0, 0, 0]
    assert O.concat_pad([], 4) == b""


REFERENCE_CONCAT_CASES = [  # the reference's own tests, src/util.rs:208-263
    ([], 4, []),
    ([[1, 2]], 4, [1, 2, 0, 0]),
    ([[1, 2, 3, 4]], 4, [1, 2, 3, 4]),
    ([[1, 2], [3], [4, 5, 6]], 4, [1, 2, 0, 0, 3, 0, 0, 0, 4, 5, 6, 0]),
    ([[1], [2, 3]], 1, [1, 2, 3]),
]


@pytest.mark.parametrize("cols,pad,want", REFERENCE_CONCAT_CASES)
def test_concat_pad_reference_cases(cols, pad, want):
    assert unmont(O.concat_pad([mont(c, M) for c in cols], pad), M) == want


# ---- FFT (src/fft.rs)
def test_fft_kat_from_reference_golden():
    """tests/golden/fft_kat_fr.json holds the vector of src/fft.rs:239-258 verbatim."""
    with open(os.path.join(GOLDEN, "fft_kat_fr.json")) as f:
        g = json.load(f)
    a = mont(g["input"], M)
    out = O.fft(FR, a, g["log_n"])
    assert unmont(out, M) == [int(x) for x in g["output"]]
    assert O.ifft(FR, out, g["log_n"]) == a


@pytest.mark.parametrize("k", [0, 1, 4, 5, 6, 7, 8, 11])
def test_fft_roundtrip_and_dft_definition(k):
    # fft_random_input_test (src/fft.rs:266-279) + the defining sum at small sizes
    rng = random.Random(k)
    n = 1 << k
    vals = [rng.randrange(M) for _ in range(n)]
    a = mont(vals, M)
    out = O.fft(FR, a, k)
    assert O.ifft(FR, out, k) == a
    if k <= 6:
        w = R.from_mont_bytes(O.fft_omega(FR, k), M)
        assert pow(w, n, M) == 1 and (n == 1 or pow(w, n // 2, M) != 1)
        assert unmont(out, M) == [sum(vals[j] * pow(w, i * j, M) for j in range(n)) % M for i in range(n)]


def test_fft_omega_limits():
    with pytest.raises(ValueError):
        O.fft_omega(FR, 29)
    with pytest.raises(ValueError):
        O.fft_omega(FQ, 2)


def test_golden_eval_vectors_freeze_the_oracle():
    import hashlib
    with open(os.path.join(GOLDEN, "eval_vectors.json")) as f:
        vecs = json.load(f)
    cache = {}
    for v in vecs:
        key = (v["T"], v["n_gates"])
        if key not in cache:
            progs, meta = G.cross_term_programs(v["T"], v["n_gates"], M)
            d = Domain(M, v["rows"], 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"],
                       seed=v["domain_seed"], sparse=True)
            cache[key] = (progs, d.as_bytes())
        progs, dom = cache[key]
        p = progs[v["term"] - 1]
        assert len(p.calculations) == v["nodes"]
        assert hashlib.sha256(O.eval_rows(FR, pack_program(p), dom)).hexdigest() == v["sha256"]


# ---- lookup argument (src/plonk/lookup.rs:278-319), SPS rounds 2 / 3
def _lookup_case(m, seed, n_l=400, n_t=128):
    rng = random.Random(seed)
    table = [rng.randrange(m) for _ in range(n_t - 8)] + [0, 1, m - 1]
    t = table + [table[0], table[5], table[5], 0, rng.randrange(m)][: n_t - len(table)]     # duplicates in t
    l = [rng.choice(table + [rng.randrange(m)]) for _ in range(n_l)]
    return l, t


@pytest.mark.parametrize("field,m", [(FR, R.R_), (FQ, R.P)])
def test_lookup_m_h_g_vs_python(field, m):
    l, t = _lookup_case(m, 11)
    got = unmont(O.lookup_m(field, mont(l, m), mont(t, m)), m)
    seen, want = set(), []
    for x in t:                                   # the reference's loop, literally
        want.append(0 if x in seen else l.count(x))
        seen.add(x)
    assert got == want and sum(got) <= len(l) and any(v > 1 for v in got)
    r = (-l[3]) % m                               # forces a zero denominator
    n = len(t)
    h, g = O.lookup_h_g(field, mont(l[:n], m), mont(t, m), mont(got, m), R.to_mont_bytes(r, m))
    inv = lambda x: pow(x, m - 2, m) if x % m else 0
    assert unmont(h, m) == [inv((x + r) % m) for x in l[:n]] and 0 in unmont(h, m)
    assert unmont(g, m) == [c * inv((x + r) % m) % m for x, c in zip(t, got)]
    assert O.lookup_m(field, b"", mont(t, m)) == bytes(32 * n) and O.lookup_m(field, mont(l, m), b"") == b""


def test_log_derivative_identity():
    """sum_i h_i == sum_i g_i when every l value is in the table (what the lookup argument proves), for the h, g, m
    the three functions produce together."""
    m = M
    rng = random.Random(12)
    table = [rng.randrange(m) for _ in range(64)]
    l = [rng.choice(table) for _ in range(64)]
    ms = O.lookup_m(FR, mont(l, m), mont(table, m))
    r = R.to_mont_bytes(rng.randrange(m), m)
    h, g = O.lookup_h_g(FR, mont(l, m), mont(table, m), ms, r)
    assert sum(unmont(h, m)) % m == sum(unmont(g, m)) % m


def test_lookup_eval_domain_uses_separate_advice_columns():
    """LookupEvalDomain::eval_advice_var (src/plonk/eval.rs:125-135): advice[index][row]."""
    d = Domain(M, 6, 1, 2, 3, 0, 1, 1, seed=13)
    cols = [d.w1[0][c * 6:(c + 1) * 6] for c in range(3)]
    dom = {"row_size": 6, "num_advice": 3, "num_lookup": 1, "selectors": [bytes(s) for s in d.selectors],
           "fixed": [mont(c, M) for c in d.fixed], "w1": [mont(c, M) for c in cols], "w2": [], "challenges": mont(d.challenges, M),
           "flags": 1}
    e = (G.Polynomial(3 + 1) * G.Challenge(0) + G.Polynomial(3 + 2, 1)) * G.Polynomial(1) + G.Polynomial(0)
    ge = G.GraphEvaluator.new(e, M)
    assert unmont(O.eval_rows(FR, pack_program(ge), dom), M) == d.direct(e, range(6))
    with pytest.raises(O.EvalError) as ex:
        O.eval_rows(FR, pack_program(G.GraphEvaluator.new(G.Polynomial(3 + 3), M)), dom)
    assert ex.value.rc == -12        # ColumnVariableIndexOutOfBoundary
    dom["w1"][2] = dom["w1"][2][:32 * 4]
    with pytest.raises(O.EvalError) as ex:
        O.eval_rows(FR, pack_program(G.GraphEvaluator.new(G.Polynomial(3 + 2), M)), dom)
    assert ex.value.rc == -13        # RowIndexOutOfBoundary


@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_is_sat_relaxed_holds_on_the_folded_instance(T, n_gates):
    """`is_sat_relaxed` (src/plonk/mod.rs:495-560) on the CPU: fold a satisfied incoming instance into a relaxed
    accumulator through the oracle's evaluator, fold and commit; the homogeneous gate program evaluated on the folded
    witness equals the folded error vector row by row, and re-committing W and E gives the homomorphically folded
    commitments.  The same construction runs on the GPU in tests/test_gpu_witness.py."""
    from witness_util import RelaxedFold
    f = RelaxedFold(M, 6, T, n_gates, seed=1000 + n_gates)
    assert all(v == 0 for v in f.incoming_gate_values())                 # the incoming instance is satisfied
    b = f.bytes()
    rows, meta = f.rows, f.meta
    cross = [O.eval_rows(FR, pack_program(p), f.domain_bytes(b["w1"], b["w2"], b["challenges"])) for p in f.progs]
    assert unmont(cross[-1], M) == [0] * rows                            # T_d = the gate on the incoming instance alone
    w_f = O.fold_w(FR, b["w1"], b["w2"], b["r"])
    e_f = O.fold_e(FR, b["e1"], cross, b["r"])
    w_int, u, c = f.folded()
    assert unmont(w_f, M) == w_int
    # evaluation check of is_sat_relaxed: hom(W'; c', u') == E' for every row (oracle evaluator, then Python integers)
    hom = pack_program(meta["hom_program"])
    got = O.eval_rows(FR, hom, f.domain_bytes(w_f, None, mont(c + [u], M)))
    assert got == e_f
    assert unmont(e_f, M) == f.hom_on(w_int, c, u)
    # commitment checks of is_sat_relaxed: commit(W') == C_W1 + r C_W2, commit(E') == C_E1 + sum r^k C_Tk
    n_w = meta["num_advice"] * rows
    bases = O.gen_bases(R.BN254, 31337, n_w)
    cw = O.point_add(R.BN254, O.commit(R.BN254, bases, b["w1"]), O.scalar_mul(R.BN254, O.commit(R.BN254, bases, b["w2"]), b["r"]))
    assert O.commit(R.BN254, bases, w_f) == cw
    ce = O.commit(R.BN254, bases[: 64 * rows], b["e1"])
    rk = 1
    for t in cross:
        rk = rk * f.r % M
        ce = O.point_add(R.BN254, ce, O.scalar_mul(R.BN254, O.commit(R.BN254, bases[: 64 * rows], t), R.to_mont_bytes(rk, M)))
    assert O.commit(R.BN254, bases[: 64 * rows], e_f) == ce
