"""CPU-only checks of the host-side linker in mira_b200/csrc/witness.cu (Store forwarding, Horner expansion,
dead-code removal, liveness-based slot allocation): the linked device program is fetched through the
`mira_test_eval_link` hook and SIMULATED here in Python with big integers, then compared with the oracle's
evaluation of the original program.  No compute entry point of the library is called (no GPU here)."""
import ctypes as C
import os
import random

import pytest

import graph_evaluator_model as G
import oracle_lib as O
import pyref as R
from witness_util import Domain, mont, pack_program, random_expr, unmont

M = R.R_


def _host_domain(d: Domain):
    """mira_eval_domain whose column pointers are HOST buffers (the linker only records them)."""
    from mira_b200 import _native as N
    b = d.as_bytes()
    keep = []

    def arr(bufs):
        ks = [C.create_string_buffer(bytes(x), max(len(x), 1)) for x in bufs]
        keep.extend(ks)
        a = (C.c_void_p * max(len(ks), 1))(*[C.cast(k, C.c_void_p).value for k in ks])
        keep.append(a)
        return a, ks
    sel, sel_b = arr(b["selectors"])
    fx, fx_b = arr(b["fixed"])
    w1, w1_b = arr(b["w1"])
    w2, w2_b = arr(b["w2"])
    l1 = (C.c_uint64 * max(len(b["w1"]), 1))(*[len(x) // 32 for x in b["w1"]])
    l2 = (C.c_uint64 * max(len(b["w2"]), 1))(*[len(x) // 32 for x in b["w2"]])
    ch = C.create_string_buffer(b["challenges"], max(len(b["challenges"]), 1))
    keep += [l1, l2, ch]
    dom = N.EvalDomain(d.row_size, len(b["selectors"]), len(b["fixed"]), d.num_advice, d.num_lookup, len(b["challenges"]) // 32,
                       len(b["w1"]), len(b["w2"]), 0, C.cast(sel, C.c_void_p), C.cast(fx, C.c_void_p), C.cast(w1, C.c_void_p),
                       C.cast(l1, C.c_void_p), C.cast(w2, C.c_void_p), C.cast(l2, C.c_void_p), C.cast(ch, C.c_void_p))
    return dom, keep


def link_and_simulate_multi(ges, d: Domain):
    """Links the programs (one merged device program), simulates it row by row; returns (rc, [outputs per program], stats)."""
    from mira_b200 import _native as N
    L = N.lib()
    handles = []
    for ge in ges:
        p = pack_program(ge)
        h = C.c_void_p()
        code = (C.c_uint32 * max(len(p["code"]), 1))(*p["code"])
        rots = (C.c_int32 * max(len(p["rotations"]), 1))(*p["rotations"])
        assert L.mira_eval_program_create(N.MIRA_FR, code, len(p["code"]), p["constants"], len(p["constants"]) // 32, rots,
                                          len(p["rotations"]), p["num_intermediates"], C.byref(h)) == 0
        handles.append(h)
    dom, keep = _host_domain(d)
    cap = 8192
    iw = (C.c_uint32 * (4 * cap))()
    aw = (C.c_uint64 * (3 * cap))()
    ub = C.create_string_buffer(32 * cap)
    ni, na, nu = C.c_size_t(), C.c_size_t(), C.c_size_t()
    ns = C.c_uint32()
    ph = (C.c_void_p * len(handles))(*[h.value for h in handles])
    rc = L.mira_test_eval_link_multi(ph, len(handles), C.byref(dom), iw, cap, C.byref(ni), aw, cap, C.byref(na), ub, cap, C.byref(nu),
                                     C.byref(ns))
    st = N.EvalStats()
    if not rc:
        L.mira_eval_program_stats(handles[0], C.byref(st))
    for h in handles:
        L.mira_eval_program_destroy(h)
    if rc:
        return rc, None, None
    uniforms = unmont(ub.raw[: 32 * nu.value], M)
    rows = d.row_size

    def fetch(kind, idx, slots, row):
        if kind == 0:
            assert slots[idx] is not None, "slot read before it was written"
            return slots[idx]
        if kind == 1:
            return uniforms[idx]
        ptr, rot, is_sel = aw[3 * idx], C.c_int64(aw[3 * idx + 1]).value, aw[3 * idx + 2]
        r = (row + rot) % rows
        if is_sel:
            return 1 if C.string_at(ptr + r, 1)[0] else 0
        return R.from_mont_bytes(C.string_at(ptr + 32 * r, 32), M)

    outs = [[None] * rows for _ in ges]
    for row in range(rows):
        slots = [None] * max(ns.value, 1)
        for k in range(ni.value):
            w0, a, b, w3 = iw[4 * k], iw[4 * k + 1], iw[4 * k + 2], iw[4 * k + 3]
            op, ak, bk, dst = w0 & 0xf, (w0 >> 4) & 0xf, (w0 >> 8) & 0xf, w0 >> 16
            x = fetch(ak, a, slots, row)
            if op == 9:                                  # out[dst] = a
                assert outs[dst][row] is None
                outs[dst][row] = x % M
                continue
            assert dst < ns.value
            if op in (7, 8):                             # fused a*b +- c*d
                y = fetch(bk, b, slots, row)
                c = fetch((w3 >> 14) & 3, w3 & 0x3fff, slots, row)
                dd = fetch(w3 >> 30, (w3 >> 16) & 0x3fff, slots, row)
                v = x * y + c * dd if op == 7 else x * y - c * dd
            elif op <= 2:
                y = fetch(bk, b, slots, row)
                v = (x + y) if op == 0 else (x - y) if op == 1 else x * y
            elif op == 3: v = x * x
            elif op == 4: v = 2 * x
            elif op == 5: v = -x
            else: v = x
            slots[dst] = v % M
    assert all(v is not None for o in outs for v in o)
    return 0, outs, {"instructions": st.instructions, "slots": ns.value, "accesses": na.value, "muls": st.muls, "adds": st.adds,
                     "loads": st.loads, "fused": st.fused, "device_words": ni.value}


def link_and_simulate(ge: G.GraphEvaluator, d: Domain):
    rc, outs, st = link_and_simulate_multi([ge], d)
    return rc, (outs[0] if outs else None), st


@pytest.mark.parametrize("seed", range(int(os.environ.get("MIRA_EXPR_SEEDS", "8"))))      # raise for a soak run
def test_linked_random_programs_match_oracle(seed):
    rng = random.Random(300 + seed)
    d = Domain(M, 6, 2, 3, 4, 0, 1, 3, seed=seed, sparse=(seed % 2 == 0))
    for _ in range(6):
        e = random_expr(rng, M, 2 + 3 + 8, 3, depth=7, rotations=(0, 1, -1, 4))
        ge = G.GraphEvaluator.new(e, M)
        rc, out, st = link_and_simulate(ge, d)
        assert rc == 0
        assert out == unmont(O.eval_rows(R.FR, pack_program(ge), d.as_bytes()), M) == d.direct(e, range(6))
        assert st["instructions"] <= len(ge.calculations)


@pytest.mark.parametrize("T,n_gates,max_slots", [(5, 1, 16), (5, 2, 64)])
def test_linked_cross_term_programs(T, n_gates, max_slots):
    progs, meta = G.cross_term_programs(T, n_gates, M)
    d = Domain(M, 3, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=55, sparse=True)
    for p in progs:
        rc, out, st = link_and_simulate(p, d)
        assert rc == 0
        assert out == unmont(O.eval_rows(R.FR, pack_program(p), d.as_bytes()), M)
        c = p.counts()
        assert st["muls"] == c["mul"] and st["adds"] == c["add"]        # nothing but Stores is removed
        assert st["instructions"] == c["mul"] + c["add"] - 2 * st["fused"]
        assert st["fused"] > 0.2 * c["add"]                                # the sum-of-two-products fusion fires
        assert st["slots"] <= max_slots, st


def test_fusion_can_be_disabled_and_gives_the_same_values(monkeypatch):
    progs, meta = G.cross_term_programs(5, 1, M)
    d = Domain(M, 3, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=56, sparse=True)
    rc, fused_out, st1 = link_and_simulate(progs[1], d)
    monkeypatch.setenv("MIRA_EVAL_FUSE", "0")
    rc2, plain_out, st0 = link_and_simulate(progs[1], d)
    assert rc == rc2 == 0 and fused_out == plain_out
    assert st0["fused"] == 0 and st1["fused"] > 0 and st1["instructions"] == st0["instructions"] - 2 * st1["fused"]
    # a - b with both products: difference of two squares, and an ADD of the same product twice is left alone
    e = (G.Polynomial(meta["num_fixed"]) * G.Polynomial(meta["num_fixed"])) - (G.Polynomial(1) * G.Polynomial(1))
    monkeypatch.delenv("MIRA_EVAL_FUSE")
    rc, out, st = link_and_simulate(G.GraphEvaluator.new(e, M), d)
    assert rc == 0 and st["fused"] == 1 and st["instructions"] == 1 and out == d.direct(e, range(3))
    x = G.Polynomial(2) * G.Polynomial(3)
    rc, out, st = link_and_simulate(G.GraphEvaluator.new(x + x, M), d)
    assert rc == 0 and st["fused"] == 0 and out == d.direct(x + x, range(3))


def test_linker_horner_non_ssa_undefined_and_empty():
    d = Domain(M, 4, 0, 2, 2, 0, 1, 1, seed=9)
    K, I, P, CH = G.VS_CONSTANT, G.VS_INTERMEDIATE, G.VS_POLY, G.VS_CHALLENGE
    ge = G.GraphEvaluator(M)
    ge.constants += [12345, M - 7]
    ge.rotations = [0, 1]
    ge.calculations = [
        ((G.OP_STORE, (P, 2, 0)), 0),
        ((G.OP_STORE, (P, 3, 1)), 1),
        ((G.OP_HORNER, (I, 0, 0), (CH, 0, 0), (I, 1, 0), (K, 3, 0), (P, 0, 0)), 2),
        ((G.OP_HORNER, (I, 2, 0), (K, 4, 0)), 3),
        ((G.OP_MUL, (I, 3, 0), (I, 5, 0)), 4),          # intermediate 5 is never written: reads ZERO
        ((G.OP_ADD, (I, 4, 0), (I, 2, 0)), 2),          # re-defines 2
        ((G.OP_SQUARE, (I, 2, 0)), 6),
    ]
    ge.num_intermediates = 7
    rc, out, st = link_and_simulate(ge, d)
    assert rc == 0
    assert out == unmont(O.eval_rows(R.FR, pack_program(ge), d.as_bytes()), M)
    empty = G.GraphEvaluator(M)
    rc, out, st = link_and_simulate(empty, d)
    assert rc == 0 and out == [0] * 4 and st["instructions"] == 0
    dead = G.GraphEvaluator(M)                          # a dead calculation before the result is dropped
    dead.rotations = [0]
    dead.calculations = [((G.OP_MUL, (P, 2, 0), (P, 3, 0)), 0), ((G.OP_STORE, (P, 1, 0)), 1)]
    dead.num_intermediates = 2
    rc, out, st = link_and_simulate(dead, d)
    assert rc == 0 and st["instructions"] == 0 and out == [d.column(1, r) for r in range(4)]


def test_linker_reports_the_reference_errors():
    from mira_b200 import _native as N
    d = Domain(M, 4, 0, 1, 2, 0, 1, 1, seed=5)
    for e, want in ((G.Challenge(1), N.MIRA_ERR_EVAL_CHALLENGE), (G.Polynomial(1 + 4), N.MIRA_ERR_EVAL_WITNESS_INDEX)):
        rc, _, _ = link_and_simulate(G.GraphEvaluator.new(e, M), d)
        assert rc == want
    bad = G.GraphEvaluator(M)
    bad.calculations = [((G.OP_ADD, (G.VS_CONSTANT, 99, 0), (G.VS_CONSTANT, 0, 0)), 0)]
    bad.num_intermediates = 1
    assert link_and_simulate(bad, d)[0] == N.MIRA_ERR_EVAL_PROGRAM


@pytest.mark.parametrize("T,n_gates", [(5, 1), (5, 2)])
def test_merged_cross_terms_share_subproducts_and_match(T, n_gates):
    """mira_eval_rows_multi's linker: all cross terms of a circuit as ONE program.  Every output equals the oracle's
    evaluation of its own program, and value numbering removes the products the terms have in common."""
    progs, meta = G.cross_term_programs(T, n_gates, M)
    d = Domain(M, 3, 0, meta["num_fixed"], meta["num_advice"], 0, 1, meta["num_challenges"], seed=57, sparse=True)
    rc, outs, st = link_and_simulate_multi(progs, d)
    assert rc == 0
    for p, o in zip(progs, outs):
        assert o == unmont(O.eval_rows(R.FR, pack_program(p), d.as_bytes()), M)
    separate = sum(p.counts()["mul"] for p in progs)
    assert st["muls"] < 0.6 * separate, (st, separate)           # 431 -> ~225, 1306 -> ~487
    assert st["slots"] <= 256


def test_merged_programs_with_distinct_constants_and_a_constant_output():
    d = Domain(M, 4, 1, 2, 2, 0, 1, 2, seed=58)
    rng = random.Random(59)
    exprs = [random_expr(rng, M, 1 + 2 + 4, 2, depth=6, rotations=(0, 1, -2)) for _ in range(5)]
    exprs.append(G.Constant(77))                                    # an output that is a uniform
    exprs.append(exprs[0])                                          # the same program twice
    ges = [G.GraphEvaluator.new(e, M) for e in exprs]
    rc, outs, st = link_and_simulate_multi(ges, d)
    assert rc == 0
    for e, o in zip(exprs, outs):
        assert o == d.direct(e, range(4))
    # programs that re-assign a target cannot be merged
    from mira_b200 import _native as N
    bad = G.GraphEvaluator(M)
    bad.rotations = [0]
    bad.calculations = [((G.OP_MUL, (G.VS_POLY, 1, 0), (G.VS_POLY, 2, 0)), 0), ((G.OP_ADD, (G.VS_INTERMEDIATE, 0, 0), (G.VS_POLY, 1, 0)), 0)]
    bad.num_intermediates = 1
    assert link_and_simulate_multi([ges[0], bad], d)[0] == N.MIRA_ERR_EVAL_PROGRAM
    rc, outs, _ = link_and_simulate_multi([bad], d)                # ... but still work alone
    assert rc == 0 and outs[0] == [(d.column(1, r) * d.column(2, r) + d.column(1, r)) % M for r in range(4)]
