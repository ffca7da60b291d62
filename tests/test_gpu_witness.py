"""GPU parity tests for the witness side (cross-term row evaluator, fold, concat, FFT): the CUDA path through the
C ABI against the CPU oracle on identical bytes — bit-exact.  Run on the B200 box: pytest -m gpu."""
import hashlib
import json
import os
import random

import pytest

import graph_evaluator_model as G
import oracle_lib as O
import pyref as R
from witness_util import Domain, mont, pack_program, random_expr, unmont

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FR, FQ = R.FR, R.FQ
M = R.R_
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def W():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mira_b200 import witness
    return witness


def dev(b: bytes):
    return torch.frombuffer(bytearray(b) if b else bytearray(1), dtype=torch.uint8).cuda()[: len(b)]


def host(t) -> bytes:
    return t.cpu().numpy().tobytes()


def gpu_domain(W, d: Domain):
    b = d.as_bytes()
    return W.PlonkEvalDomain(d.num_advice, d.num_lookup, b["challenges"], [dev(s) for s in b["selectors"]],
                             [dev(f) for f in b["fixed"]], [dev(w) for w in b["w1"]], [dev(w) for w in b["w2"]],
                             row_size=d.row_size)


def gpu_eval(W, field, ge: G.GraphEvaluator, gd):
    p = pack_program(ge)
    prog = W.GraphEvaluator(field, p["code"], p["constants"], p["rotations"], p["num_intermediates"])
    out = host(prog.evaluate_rows(gd))
    st = prog.stats()
    prog.close()
    return out, st


def both(W, field, expr, d: Domain, gd=None):
    ge = G.GraphEvaluator.new(expr, d.m)
    want = O.eval_rows(field, pack_program(ge), d.as_bytes())
    got, _ = gpu_eval(W, field, ge, gd or gpu_domain(W, d))
    return got, want


# ---- restatements of the reference's evaluator tests (graph_evaluator.rs:447-634), GPU vs oracle vs direct
def test_constants_and_challenges(W):
    d = Domain(M, 3, 0, 1, 0, 0, 1, 3, seed=3)
    gd = gpu_domain(W, d)
    rng = random.Random(1)
    a, b = rng.randrange(M), rng.randrange(M)
    for e, want in ((G.Constant(a), a), (G.Constant(a) + G.Constant(b), (a + b) % M), (G.Constant(a) * G.Constant(b), a * b % M),
                    (-G.Constant(a), (-a) % M), (G.Challenge(2), d.challenges[2]), (G.Constant(0), 0)):
        got, ora = both(W, FR, e, d, gd)
        assert got == ora == mont([want] * 3, M)


def test_poly_rotations_and_column_dispatch(W):
    d = Domain(M, 2, 2, 2, 2, 0, 1, 0, seed=2)
    gd = gpu_domain(W, d)
    for base in (0, 2, 4, 6):         # selectors, fixed, advice of instance 1, advice of instance 2
        for col in range(2):
            for rot in (0, 1, -1, 2, -3):
                e = G.Polynomial(base + col, rot)
                got, ora = both(W, FR, e, d, gd)
                assert got == ora == mont(d.direct(e, range(2)), M)


def test_eval_example(W):
    d = Domain(M, 2, 2, 2, 2, 0, 1, 0, seed=4)
    adv = lambda c: G.Polynomial(4 + c)
    zero = G.Constant(0)
    e = (adv(0) + (adv(1) + (adv(1) + zero))) * (G.Polynomial(2) + (adv(0) + zero))
    got, ora = both(W, FR, e, d)
    assert got == ora == mont(d.direct(e, range(2)), M)


def test_index_errors_match_the_reference(W):
    d = Domain(M, 4, 0, 1, 2, 0, 1, 1, seed=5)
    gd = gpu_domain(W, d)
    for e, kind in ((G.Challenge(1), "ChallengeIndexOutOfBoundary"), (G.Polynomial(1 + 2 * 2), "InvalidWitnessIndex")):
        ge = G.GraphEvaluator.new(e, M)
        with pytest.raises(O.EvalError):
            O.eval_rows(FR, pack_program(ge), d.as_bytes())
        with pytest.raises(W.EvalError) as ex:
            gpu_eval(W, FR, ge, gd)
        assert ex.value.kind == kind
    # a witness vector shorter than (j+1)*row_size is InvalidWitnessIndex too (src/plonk/eval.rs:206-214)
    short = gpu_domain(W, d)
    short.W1s = [short.W1s[0][: 32 * 7]]
    with pytest.raises(W.EvalError) as ex:
        gpu_eval(W, FR, G.GraphEvaluator.new(G.Polynomial(1 + 1), M), short)
    assert ex.value.kind == "InvalidWitnessIndex"


@pytest.mark.parametrize("field,m", [(FR, R.R_), (FQ, R.P)])
@pytest.mark.parametrize("seed", range(int(os.environ.get("MIRA_EXPR_SEEDS", "4"))))      # raise for a soak run
def test_random_expressions(W, field, m, seed):
    rng = random.Random(200 + seed)
    n_sel, n_fix, n_adv, n_ch = 2, 3, 4, 3
    rows = 37                                   # ragged: not a multiple of the block size
    d = Domain(m, rows, n_sel, n_fix, n_adv, 0, 1, n_ch, seed=seed, sparse=(seed % 2 == 1))
    gd = gpu_domain(W, d)
    n_cols = n_sel + n_fix + 2 * n_adv
    for _ in range(8):
        e = random_expr(rng, m, n_cols, n_ch, depth=7, rotations=(0, 1, -1, 5))
        got, ora = both(W, field, e, d, gd)
        assert got == ora
    e = random_expr(rng, m, n_cols, n_ch, depth=5)
    assert both(W, field, e, d, gd)[0] == mont(d.direct(e, range(rows)), m)


def test_lookup_column_mapping(W):
    for n_w in (2, 3):
        d = Domain(M, 4, 1, 1, 2, 2, n_w, 0, seed=20 + n_w)
        gd = gpu_domain(W, d)
        for idx in range(2 * (2 + 5 * 2)):
            got, ora = both(W, FR, G.Polynomial(2 + idx), d, gd)
            assert got == ora


def test_horner_and_non_ssa_programs(W):
    """Calculation::Horner is never emitted by add_expression but is part of the evaluator (graph_evaluator.rs:139-146)."""
    d = Domain(M, 5, 0, 2, 2, 0, 1, 1, seed=9)
    gd = gpu_domain(W, d)
    K, I, P, CH = G.VS_CONSTANT, G.VS_INTERMEDIATE, G.VS_POLY, G.VS_CHALLENGE
    ge = G.GraphEvaluator(M)
    ge.constants += [12345, M - 7]
    ge.rotations = [0, 1]
    ge.calculations = [
        ((G.OP_STORE, (P, 2, 0)), 0),
        ((G.OP_STORE, (P, 3, 1)), 1),
        ((G.OP_HORNER, (I, 0, 0), (CH, 0, 0), (I, 1, 0), (K, 3, 0), (P, 0, 0)), 2),   # start, factor, parts...
        ((G.OP_HORNER, (I, 2, 0), (K, 4, 0)), 3),                                       # no parts: plain copy
        ((G.OP_MUL, (I, 3, 0), (I, 5, 0)), 4),                                          # intermediate 5 never written: ZERO
        ((G.OP_ADD, (I, 4, 0), (I, 2, 0)), 2),                                          # re-defines 2 (non-SSA)
        ((G.OP_SQUARE, (I, 2, 0)), 6),
    ]
    ge.num_intermediates = 7
    want = O.eval_rows(FR, pack_program(ge), d.as_bytes())
    got, st = gpu_eval(W, FR, ge, gd)
    assert got == want
    # direct: h = ((a*c + b)*c + k)*c + f0 ; result = h^2
    c = d.challenges[0]
    exp = []
    for r in range(5):
        a, b, f0 = d.column(2, r), d.column(3, (r + 1) % 5), d.column(0, r)
        h = (((a * c + b) * c + 12345) * c + f0) % M
        exp.append(h * h % M)
    assert got == mont(exp, M)


@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_cross_term_programs_vs_oracle(W, T, n_gates):
    progs, meta = G.cross_term_programs(T, n_gates, M)
    rows = 1 << 11
    d = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=77 + n_gates, sparse=True)
    gd = gpu_domain(W, d)
    db = d.as_bytes()
    for p in progs:
        got, st = gpu_eval(W, FR, p, gd)
        assert got == O.eval_rows(FR, pack_program(p), db)
        assert st["muls"] == p.counts()["mul"] and st["slots"] <= 64 and st["fused"] > 0


def test_golden_eval_vectors(W):
    with open(os.path.join(GOLDEN, "eval_vectors.json")) as f:
        vecs = json.load(f)
    cache = {}
    for v in vecs:
        key = (v["T"], v["n_gates"])
        if key not in cache:
            progs, meta = G.cross_term_programs(v["T"], v["n_gates"], M)
            d = Domain(M, v["rows"], 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"],
                       seed=v["domain_seed"], sparse=True)
            cache[key] = (progs, gpu_domain(W, d))
        progs, gd = cache[key]
        got, _ = gpu_eval(W, FR, progs[v["term"] - 1], gd)
        assert hashlib.sha256(got).hexdigest() == v["sha256"]


def test_cross_terms_feed_commit_without_leaving_the_device(W):
    """commit_cross_terms (src/nifs/vanilla/mod.rs:80-140): evaluate on the GPU, commit the device vector, and
    compare with the oracle doing both steps on the CPU."""
    from mira_b200 import BN254_G1, CommitmentKey
    progs, meta = G.cross_term_programs(5, 1, M)
    rows = 1 << 10
    d = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=123, sparse=True)
    gd = gpu_domain(W, d)
    bases = O.gen_bases(R.BN254, 0x4D495241, rows)
    ck = CommitmentKey(BN254_G1, bases)
    for p in progs[:3]:
        pk = pack_program(p)
        prog = W.GraphEvaluator(FR, pk["code"], pk["constants"], pk["rotations"], pk["num_intermediates"])
        t_dev = prog.evaluate_rows(gd)
        torch.cuda.synchronize()
        got = ck.commit_device(t_dev.data_ptr(), rows)
        want = O.commit(R.BN254, bases, O.eval_rows(FR, pk, d.as_bytes()))
        assert got == want


# ---- fold (src/plonk/mod.rs:1097-1134)
@pytest.mark.parametrize("field,curve,m", [(FR, R.BN254, R.R_), (FQ, R.GRUMPKIN, R.P)])
def test_fold_w_and_e(W, field, curve, m):
    n = 100_003
    w1 = O.gen_scalars(curve, 1, n, 1)
    w2 = O.gen_scalars(curve, 2, n, 0)
    r = O.gen_scalars(curve, 3, 1, 0)
    d1, d2 = dev(w1), dev(w2)
    assert host(W.fold_w(field, d1, d2, r)) == O.fold_w(field, w1, w2, r)
    for n_terms in (0, 1, 5, 6):
        ts = [O.gen_scalars(curve, 10 + k, n, k % 2) for k in range(n_terms)]
        got = host(W.fold_e(field, d1, [dev(t) for t in ts], r))
        assert got == O.fold_e(field, w1, ts, r)
    W.fold_w(field, d1, d2, r, out=d1)          # in place, as a Rust `W = ...` replacing the accumulator would
    assert host(d1) == O.fold_w(field, w1, w2, r)
    edge = mont([0, 1, m - 1, m - 2, 2], m)
    rm1 = R.to_mont_bytes(m - 1, m)
    assert host(W.fold_w(field, dev(edge), dev(edge), rm1)) == O.fold_w(field, edge, edge, rm1) == bytes(160)
    assert W.fold_w(field, dev(b""), dev(b""), r).numel() == 0


def test_fold_is_what_is_sat_relaxed_checks(W):
    """commit(W1 + r*W2) == commit(W1) + r*commit(W2): fold on the GPU, commit on the GPU, combine on the CPU
    (src/plonk/mod.rs:547-557 via src/nifs/vanilla/tests.rs:137-244)."""
    from mira_b200 import BN254_G1, CommitmentKey
    n = 3000
    bases = O.gen_bases(R.BN254, 5, n)
    ck = CommitmentKey(BN254_G1, bases)
    w1, w2 = O.gen_scalars(R.BN254, 6, n, 1), O.gen_scalars(R.BN254, 7, n, 1)
    r = O.gen_scalars(R.BN254, 8, 1)
    folded = W.fold_w(FR, dev(w1), dev(w2), r)
    torch.cuda.synchronize()
    lhs = ck.commit_device(folded.data_ptr(), n)
    rhs = O.point_add(R.BN254, ck.commit(w1), O.scalar_mul(R.BN254, ck.commit(w2), r))
    assert lhs == rhs


# ---- concatenate_with_padding (src/util.rs:189-193)
def test_concat_pad(W):
    cols = [mont(c, M) for c in ([1, 2, 3], [], [4, 5, 6, 7, 8], [9])]
    got = host(W.concatenate_with_padding([dev(c) for c in cols], 4))
    assert got == O.concat_pad(cols, 4)
    big = [O.gen_scalars(R.BN254, 40 + i, 1000 + i, 1) for i in range(7)]
    assert host(W.concatenate_with_padding([dev(c) for c in big], 1024)) == O.concat_pad(big, 1024)


@pytest.mark.parametrize("cols,pad,want", [  # the reference's own tests, src/util.rs:208-263
    ([], 4, []),
    ([[1, 2]], 4, [1, 2, 0, 0]),
    ([[1, 2, 3, 4]], 4, [1, 2, 3, 4]),
    ([[1, 2], [3], [4, 5, 6]], 4, [1, 2, 0, 0, 3, 0, 0, 0, 4, 5, 6, 0]),
    ([[1], [2, 3]], 1, [1, 2, 3]),
])
def test_concat_pad_reference_cases(W, cols, pad, want):
    got = host(W.concatenate_with_padding([dev(mont(c, M)) for c in cols], pad))
    assert unmont(got, M) == want


# ---- FFT (src/fft.rs)
def test_fft_reference_kat(W):
    with open(os.path.join(GOLDEN, "fft_kat_fr.json")) as f:
        g = json.load(f)
    a = dev(mont(g["input"], M))
    W.fft(FR, a, g["log_n"])
    assert unmont(host(a), M) == [int(x) for x in g["output"]]
    W.ifft(FR, a, g["log_n"])
    assert unmont(host(a), M) == g["input"]


@pytest.mark.parametrize("k", [0, 1, 2, 4, 5, 6, 7, 8, 10, 11, 12, 13, 16, 18, 19])
def test_fft_vs_oracle_and_roundtrip(W, k):
    n = 1 << k
    vals = O.gen_scalars(R.BN254, 900 + k, n, 0)
    a = dev(vals)
    W.fft(FR, a, k)
    assert host(a) == O.fft(FR, vals, k)
    W.ifft(FR, a, k)
    assert host(a) == vals                      # fft_random_input_test (src/fft.rs:266-279)
    w = O.fft_omega(FR, k, inverse=True)        # best_fft with a caller-supplied omega
    b = dev(vals)
    W.best_fft(FR, b, w, k)
    assert host(b) == O.best_fft(FR, vals, k, w)


def test_fft_large_roundtrip_and_linearity(W):
    """2^22 elements: size-independent properties (round trip; FFT(a + c*b) = FFT(a) + c*FFT(b))."""
    from gpu_util import gen_scalars_dev
    k = 22
    n = 1 << k
    a = gen_scalars_dev(R.BN254, 1, n)
    b = gen_scalars_dev(R.BN254, 2, n)
    c = O.gen_scalars(R.BN254, 3, 1)
    ab = W.fold_w(FR, a, b, c)
    a0 = a.clone()
    W.fft(FR, a, k); W.fft(FR, b, k); W.fft(FR, ab, k)
    assert torch.equal(W.fold_w(FR, a, b, c), ab)
    W.ifft(FR, a, k)
    assert torch.equal(a, a0)


def test_fft_size_limits(W):
    a = dev(mont([1, 2], M))
    with pytest.raises(ValueError):
        W.fft(FQ, a, 1)                        # no ROOT_OF_UNITY path for Fq (S = 1)
    with pytest.raises(ValueError):
        W.fft(FR, a, 2)                        # assert_eq!(n, 1 << log_n)


def test_row_range_shards_concatenate_to_the_whole(W):
    """Row-range sharding: ranges evaluated separately (as ranks would) concatenate to the full
    vector, also with rotations that reach outside the range; a range past row_size is RowIndexOutOfBoundary."""
    rng = random.Random(4)
    rows = 1000
    d = Domain(M, rows, 1, 2, 3, 0, 1, 2, seed=8)
    gd = gpu_domain(W, d)
    e = random_expr(rng, M, 1 + 2 + 6, 2, depth=6, rotations=(0, 1, -1, 17, -400))
    p = pack_program(G.GraphEvaluator.new(e, M))
    prog = W.GraphEvaluator(FR, p["code"], p["constants"], p["rotations"], p["num_intermediates"])
    whole = host(prog.evaluate_rows(gd))
    assert whole == O.eval_rows(FR, p, d.as_bytes())
    cuts = [0, 1, 333, 334, 900, 1000]
    parts = b"".join(host(prog.evaluate_rows(gd, rows=(a, b))) for a, b in zip(cuts, cuts[1:]))
    assert parts == whole
    assert prog.evaluate_rows(gd, rows=(5, 5)).numel() == 0
    with pytest.raises(W.EvalError) as ex:
        prog.evaluate_rows(gd, rows=(0, rows + 1))
    assert ex.value.kind == "RowIndexOutOfBoundary"


# ---- lookup argument: SPS rounds 2 / 3 (src/plonk/mod.rs:748-907, src/plonk/lookup.rs:212-319)
@pytest.mark.parametrize("field,curve,m", [(FR, R.BN254, R.R_), (FQ, R.GRUMPKIN, R.P)])
@pytest.mark.parametrize("seed", range(int(os.environ.get("MIRA_LOOKUP_SEEDS", "1"))))      # raise for a soak run
def test_lookup_m_and_h_g(W, field, curve, m, seed):
    rng = random.Random(21 + seed)
    for n_l, n_t in ((1, 1), (300, 64), (5000, 4096), (70_000, 1 << 15)):
        table = [rng.randrange(m) for _ in range(max(n_t - 6, 1))]
        t = (table + [0, 1, m - 1, table[0], table[-1], table[-1]])[:n_t]
        l = [rng.choice(t + [rng.randrange(m)]) for _ in range(n_l)]
        lb, tb = mont(l, m), mont(t, m)
        want_m = O.lookup_m(field, lb, tb)
        ld, td = dev(lb), dev(tb)
        md = W.evaluate_m(field, ld, td)
        assert host(md) == want_m
        n = min(n_l, n_t)
        r = R.to_mont_bytes((-l[0]) % m, m)                     # a zero denominator in h
        l2, t2, m2 = dev(lb[:32 * n]), dev(tb[:32 * n]), dev(want_m[:32 * n])
        h, g = W.evaluate_h_g(field, l2, t2, r, m2)
        wh, wg = O.lookup_h_g(field, lb[:32 * n], tb[:32 * n], want_m[:32 * n], r)
        assert host(h) == wh and host(g) == wg and wh[:32] == bytes(32)
    assert W.evaluate_m(field, dev(b""), dev(tb)).cpu().numpy().tobytes() == bytes(len(tb))


def test_sps_round_with_lookup_end_to_end(W):
    """run_sps_protocol_3's lookup rounds (src/plonk/mod.rs:833-881): l, t evaluated over a LookupEvalDomain, m, then
    W2 = concat(ls, ts, ms) committed; h, g, then W3 = concat(hs, gs) committed — GPU vs oracle, bit for bit."""
    from mira_b200 import BN254_G1, CommitmentKey
    rows = 256
    d = Domain(M, rows, 1, 2, 3, 0, 1, 1, seed=31)
    rng = random.Random(32)
    # make the lookup hold: advice column 0 takes values from fixed column 0 (the table)
    table = d.fixed[0]
    cols = [[rng.choice(table) for _ in range(rows)], d.w1[0][rows:2 * rows], d.w1[0][2 * rows:3 * rows]]
    l_expr = G.Polynomial(3 + 0) + G.Challenge(0) * G.Polynomial(0)             # selector-gated lookup value
    t_expr = G.Polynomial(1) + G.Challenge(0) * G.Polynomial(0)
    dom_o = {"row_size": rows, "num_advice": 3, "num_lookup": 1, "selectors": [bytes(s) for s in d.selectors],
             "fixed": [mont(c, M) for c in d.fixed], "w1": [mont(c, M) for c in cols], "w2": [], "challenges": mont(d.challenges, M),
             "flags": 1}
    dom_g = W.LookupEvalDomain(1, dom_o["challenges"], [dev(s) for s in dom_o["selectors"]], [dev(f) for f in dom_o["fixed"]],
                               [dev(c) for c in dom_o["w1"]], row_size=rows)
    bases = O.gen_bases(R.BN254, 77, 3 * rows)
    ck = CommitmentKey(BN254_G1, bases)
    lt_o, lt_g = [], []
    for e in (l_expr, t_expr):
        p = pack_program(G.GraphEvaluator.new(e, M))
        lt_o.append(O.eval_rows(FR, p, dom_o))
        prog = W.GraphEvaluator(FR, p["code"], p["constants"], p["rotations"], p["num_intermediates"])
        lt_g.append(prog.evaluate_rows(dom_g))
        assert host(lt_g[-1]) == lt_o[-1]
    m_o = O.lookup_m(FR, lt_o[0], lt_o[1])
    m_g = W.evaluate_m(FR, lt_g[0], lt_g[1])
    w2_o = O.concat_pad([lt_o[0], lt_o[1], m_o], rows)
    w2_g = W.concatenate_with_padding([lt_g[0], lt_g[1], m_g], rows)
    torch.cuda.synchronize()
    assert host(w2_g) == w2_o
    assert ck.commit_device(w2_g.data_ptr(), 3 * rows) == O.commit(R.BN254, bases, w2_o)
    r2 = O.gen_scalars(R.BN254, 33, 1)
    h_o, g_o = O.lookup_h_g(FR, lt_o[0], lt_o[1], m_o, r2)
    h_g, g_g = W.evaluate_h_g(FR, lt_g[0], lt_g[1], r2, m_g)
    w3_g = W.concatenate_with_padding([h_g, g_g], rows)
    torch.cuda.synchronize()
    assert host(w3_g) == O.concat_pad([h_o, g_o], rows)
    assert ck.commit_device(w3_g.data_ptr(), 2 * rows) == O.commit(R.BN254, bases[:64 * 2 * rows], O.concat_pad([h_o, g_o], rows))


@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_merged_cross_terms_equal_separate_evaluation(W, T, n_gates):
    """mira_eval_rows_multi: all cross terms in one launch, shared sub-products computed once — every output is
    bit-identical to the oracle's evaluation of its own program, also on a row range."""
    progs, meta = G.cross_term_programs(T, n_gates, M)
    rows = 1 << 10
    d = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=91 + n_gates, sparse=True)
    gd = gpu_domain(W, d)
    db = d.as_bytes()
    packed = [pack_program(p) for p in progs]
    gp = [W.GraphEvaluator(FR, p["code"], p["constants"], p["rotations"], p["num_intermediates"]) for p in packed]
    want = [O.eval_rows(FR, p, db) for p in packed]
    got = W.evaluate_rows_multi(gp, gd)
    assert [host(g) for g in got] == want
    st = gp[0].stats()
    assert st["muls"] < 0.6 * sum(p.counts()["mul"] for p in progs)
    part = W.evaluate_rows_multi(gp, gd, rows=(100, 357))
    assert [host(g) for g in part] == [w[32 * 100:32 * 357] for w in want]
    assert host(gp[2].evaluate_rows(gd)) == want[2]                 # a program still evaluates alone afterwards


def test_binding_cache_refreshes_challenges_and_notices_new_columns(W):
    """The linked program is cached per (programs, column pointers); a second call with other challenge VALUES must
    use them, and a call with different column buffers must re-link."""
    progs, meta = G.cross_term_programs(5, 1, M)
    rows = 300
    packed = [pack_program(p) for p in progs]
    gp = [W.GraphEvaluator(FR, p["code"], p["constants"], p["rotations"], p["num_intermediates"]) for p in packed]
    d1 = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=71, sparse=True)
    gd = gpu_domain(W, d1)
    assert [host(g) for g in W.evaluate_rows_multi(gp, gd)] == [O.eval_rows(FR, p, d1.as_bytes()) for p in packed]
    # same device columns, new challenges (what every fold step does)
    d1.challenges = [(c * 7 + 3) % M for c in d1.challenges]
    gd.challenges = mont(d1.challenges, M)
    assert [host(g) for g in W.evaluate_rows_multi(gp, gd)] == [O.eval_rows(FR, p, d1.as_bytes()) for p in packed]
    # new contents in the SAME buffers are picked up (columns are read at run time)
    d1.fixed[0] = [(v + 1) % M for v in d1.fixed[0]]
    gd.fixed[0].copy_(dev(mont(d1.fixed[0], M)))
    assert [host(g) for g in W.evaluate_rows_multi(gp, gd)] == [O.eval_rows(FR, p, d1.as_bytes()) for p in packed]
    # different buffers: re-link
    d2 = Domain(M, rows, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=72)
    gd2 = gpu_domain(W, d2)
    assert [host(g) for g in W.evaluate_rows_multi(gp, gd2)] == [O.eval_rows(FR, p, d2.as_bytes()) for p in packed]
    assert host(gp[0].evaluate_rows(gd)) == O.eval_rows(FR, packed[0], d1.as_bytes())


# ---- is_sat_relaxed (src/plonk/mod.rs:495-560) on a folded instance, everything on the GPU
@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_is_sat_relaxed_on_the_folded_instance(W, T, n_gates):
    """The check `IVC::verify` runs on the accumulator (src/ivc/incrementally_verifiable_computation.rs:617-680) and the
    reference's folding tests rely on (src/nifs/vanilla/tests.rs:137-244), at k = 11 through the C ABI:
      commit_cross_terms   T_k = mira_eval_rows_multi over (W1, W2), C_Tk = mira_msm_commit_batch
      fold                 W' = W1 + r W2, E' = E1 + sum r^k T_k        (mira_fold_w / mira_fold_e)
      is_sat_relaxed       the un-grouped HOMOGENEOUS gate program over W' with challenges c' | u' (mira_eval_rows)
                           == E' row by row;  commit(W') == C_W1 + r C_W2;  commit(E') == C_E1 + sum r^k C_Tk
    with every device result also compared with the oracle's bytes."""
    from mira_b200 import BN254_G1, CommitmentKey
    from witness_util import RelaxedFold
    f = RelaxedFold(M, 11, T, n_gates, seed=2000 + n_gates)
    b, rows, meta = f.bytes(), f.rows, f.meta
    n_w = meta["num_advice"] * rows
    fixed = [dev(c) for c in b["fixed"]]
    w1, w2, e1 = dev(b["w1"]), dev(b["w2"]), dev(b["e1"])
    bases = O.gen_bases(R.BN254, 4711, n_w)
    ck = CommitmentKey(BN254_G1, bases)
    mk = lambda p: W.GraphEvaluator(FR, p["code"], p["constants"], p["rotations"], p["num_intermediates"])
    # commit_cross_terms
    dom12 = W.PlonkEvalDomain(meta["num_advice"], 0, b["challenges"], [], fixed, [w1], [w2])
    progs = [mk(pack_program(p)) for p in f.progs]
    cross = W.evaluate_rows_multi(progs, dom12)
    want_cross = [O.eval_rows(FR, pack_program(p), f.domain_bytes(b["w1"], b["w2"], b["challenges"])) for p in f.progs]
    assert [host(t) for t in cross] == want_cross
    assert host(cross[-1]) == bytes(32 * rows)                       # T_d vanishes: the incoming instance is satisfied
    c_t = ck.commit_batch_device([t.data_ptr() for t in cross], rows)
    assert c_t[-1] == bytes(64)                                      # commitment to the zero vector is the identity
    c_w1, c_w2, c_e1 = ck.commit(b["w1"]), ck.commit(b["w2"]), ck.commit(b["e1"])
    # fold
    w_f = W.fold_w(FR, w1, w2, b["r"])
    e_f = W.fold_e(FR, e1, cross, b["r"])
    assert host(w_f) == O.fold_w(FR, b["w1"], b["w2"], b["r"])
    assert host(e_f) == O.fold_e(FR, b["e1"], want_cross, b["r"])
    w_int, u, c = f.folded()
    # is_sat_relaxed, evaluation half: hom(W'; c', u') == E' on every row
    hom = mk(pack_program(meta["hom_program"]))
    dom_f = W.PlonkEvalDomain(meta["num_advice"], 0, mont(c + [u], M), [], fixed, [w_f], [])
    got = hom.evaluate_rows(dom_f)
    torch.cuda.synchronize()
    assert bool((got == e_f).all().item())                          # mismatch_count == 0
    assert host(got) == O.eval_rows(FR, pack_program(meta["hom_program"]), f.domain_bytes(host(w_f), None, mont(c + [u], M)))
    # ... and it does detect a violated row: perturb one cell of W'
    bad = w_f.clone()
    bad[32 * (3 * rows + 17)] ^= 1
    got_bad = hom.evaluate_rows(W.PlonkEvalDomain(meta["num_advice"], 0, mont(c + [u], M), [], fixed, [bad], []))
    diff = (got_bad.view(rows, 32) != e_f.view(rows, 32)).any(dim=1)
    assert int(diff.sum().item()) == 1 and bool(diff[17].item())
    # is_sat_relaxed, commitment half
    want_cw = O.point_add(R.BN254, c_w1, O.scalar_mul(R.BN254, c_w2, b["r"]))
    assert ck.commit_device(w_f.data_ptr(), n_w) == want_cw == O.commit(R.BN254, bases, host(w_f))
    want_ce, rk = c_e1, 1
    for ct in c_t:
        rk = rk * f.r % M
        want_ce = O.point_add(R.BN254, want_ce, O.scalar_mul(R.BN254, ct, R.to_mont_bytes(rk, M)))
    assert ck.commit_device(e_f.data_ptr(), rows) == want_ce == O.commit(R.BN254, bases[: 64 * rows], host(e_f))
    for p in progs + [hom]:
        p.close()
