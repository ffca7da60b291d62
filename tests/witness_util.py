"""Shared helpers for the witness-side tests: random domains, program packing, direct evaluation."""
from __future__ import annotations

import random

import graph_evaluator_model as G
import pyref as R


def mont(vals, m) -> bytes:
    return b"".join(R.to_mont_bytes(v, m) for v in vals)


def unmont(b: bytes, m):
    return [R.from_mont_bytes(b[i:i + 32], m) for i in range(0, len(b), 32)]


def pack_program(ge: G.GraphEvaluator) -> dict:
    """GraphEvaluator -> the dict oracle_lib.eval_rows / mira_b200.witness.EvalProgram take."""
    enc = ge.encode()
    enc["constants"] = mont(enc["constants"], ge.m)
    return enc


class Domain:
    """A PlonkEvalDomain (src/plonk/eval.rs:93-106) with integer contents, for both back ends."""

    def __init__(self, m, row_size, num_selectors, num_fixed, num_advice, num_lookup, n_w, n_challenges, seed,
                 sparse=False):
        rng = random.Random(seed)
        self.m, self.row_size = m, row_size
        self.num_advice, self.num_lookup = num_advice, num_lookup

        def val():
            if sparse:
                x = rng.random()
                if x < 0.5: return 0
                if x < 0.7: return 1
                if x < 0.85: return rng.randrange(1 << 32)
            return rng.randrange(m)

        self.selectors = [[rng.randrange(2) for _ in range(row_size)] for _ in range(num_selectors)]
        self.fixed = [[val() for _ in range(row_size)] for _ in range(num_fixed)]
        # W layout (src/plonk/mod.rs:674-907): W[0] = advice columns (+ 3 per lookup when 2 rounds), W[1], W[2] lookups
        if n_w == 1:
            lens = [num_advice * row_size]
        elif n_w == 2:
            lens = [(num_advice + 3 * num_lookup) * row_size, 2 * num_lookup * row_size]
        else:
            lens = [num_advice * row_size, 3 * num_lookup * row_size, 2 * num_lookup * row_size]
        self.w1 = [[val() for _ in range(l)] for l in lens]
        self.w2 = [[val() for _ in range(l)] for l in lens]
        self.challenges = [rng.randrange(m) for _ in range(n_challenges)]

    def as_bytes(self) -> dict:
        return {"row_size": self.row_size, "num_advice": self.num_advice, "num_lookup": self.num_lookup,
                "selectors": [bytes(s) for s in self.selectors], "fixed": [mont(c, self.m) for c in self.fixed],
                "w1": [mont(w, self.m) for w in self.w1], "w2": [mont(w, self.m) for w in self.w2],
                "challenges": mont(self.challenges, self.m)}

    # eval_column_var + PlonkEvalDomain::eval_advice_var with plain integers (independent restatement)
    def column(self, index, row):
        ns, nf = len(self.selectors), len(self.fixed)
        if index < ns: return self.selectors[index][row]
        if index < ns + nf: return self.fixed[index - ns][row]
        index -= ns + nf
        width = self.num_advice + 5 * self.num_lookup
        W = self.w1 if index < width else self.w2
        if index >= width: index -= width
        if index < self.num_advice:
            i, j = 0, index
        else:
            li, ls = divmod(index - self.num_advice, 5)
            first_round = ls < 3
            if not first_round: ls -= 3
            if len(W) == 2:
                i, j = (0, self.num_advice + 3 * li + ls) if first_round else (1, 2 * li + ls)
            else:
                i, j = (1, 3 * li + ls) if first_round else (2, 2 * li + ls)
        return W[i][j * self.row_size + row]

    def direct(self, expr, rows):
        return [G.eval_expr(expr, self.m, self.column, self.challenges, r, self.row_size) for r in rows]


def random_expr(rng, m, n_columns, n_challenges, depth, rotations=(0,)):
    """Random Expression tree over column queries, challenges and constants."""
    if depth == 0 or rng.random() < 0.15:
        x = rng.random()
        if x < 0.55 and n_columns:
            return G.Polynomial(rng.randrange(n_columns), rng.choice(rotations))
        if x < 0.7 and n_challenges:
            return G.Challenge(rng.randrange(n_challenges))
        return G.Constant(rng.choice([0, 1, 2, m - 1, rng.randrange(m)]))
    k = rng.random()
    a = random_expr(rng, m, n_columns, n_challenges, depth - 1, rotations)
    if k < 0.1: return -a
    if k < 0.2: return G.Scaled(a, rng.choice([0, 1, 2, rng.randrange(m)]))
    b = random_expr(rng, m, n_columns, n_challenges, depth - 1, rotations)
    if k < 0.5: return a + b
    if k < 0.6: return a - b
    if k < 0.65: return a * a
    return a * b


class RelaxedFold:
    """One folding step of a relaxed PLONK instance, built with plain Python integers, for `is_sat_relaxed`-shaped
    tests (src/plonk/mod.rs:495-560; src/nifs/vanilla/mod.rs:220-251):

      * incoming instance (W2, u2 = 1, E2 = 0): every MainGate row is SATISFIED -- the `out` cell of each gate is solved
        from the gate equation (it is linear in `out`), the other cells and the fixed columns are random;
      * accumulator (W1, u1, E1): random W1, random u1 and challenges, and E1 := hom(W1; c1, u1) row by row, i.e. a
        relaxed instance that satisfies its relation by construction;
      * challenge vector handed to the cross-term programs: c1 | u1 | c2 | 1 (src/nifs/vanilla/mod.rs:91).

    All vectors are exposed as Montgomery byte strings in the reference's layouts (W column-major)."""

    def __init__(self, m: int, log_rows: int, T: int, n_gates: int, seed: int):
        rng = random.Random(seed)
        self.m = m
        self.rows = rows = 1 << log_rows
        self.progs, self.meta = G.relaxed_circuit(T, n_gates, m)
        meta = self.meta
        nf, na, nch = meta["num_fixed"], meta["num_advice"], meta["per_instance_challenges"] - 1
        fpg, apg = meta["fixed_per_gate"], meta["advice_per_gate"]

        def sparse():
            x = rng.random()
            if x < 0.4: return 0
            if x < 0.6: return 1
            if x < 0.8: return rng.randrange(1 << 32)
            return rng.randrange(m)
        self.fixed = [[sparse() for _ in range(rows)] for _ in range(nf)]
        for g in range(n_gates):                      # q_o must be invertible to solve for `out`
            col = self.fixed[g * fpg + 3 * T + 1]
            for r in range(rows):
                if col[r] == 0:
                    col[r] = 1 + rng.randrange(m - 1)
        self.c1 = [rng.randrange(m) for _ in range(nch)]
        self.c2 = [rng.randrange(m) for _ in range(nch)]
        self.u1 = rng.randrange(m)
        self.w1 = [rng.randrange(m) for _ in range(na * rows)]
        self.w2 = [sparse() for _ in range(na * rows)]
        # satisfy every gate of the incoming instance: gate_g(out = 0) + q_o * out = 0
        gates = [G.main_gate_expr(T, g * apg, 0, nf, g * fpg) for g in range(n_gates)]
        for g, ge in enumerate(gates):
            out_col = g * apg + T + 1
            for r in range(rows):
                self.w2[out_col * rows + r] = 0
                rest = G.eval_expr(ge, m, self._col(self.w2), [], r, rows)
                q_o = self.fixed[g * fpg + 3 * T + 1][r]
                self.w2[out_col * rows + r] = (-rest * pow(q_o, -1, m)) % m
        # E1 := hom(W1; c1, u1)
        self.e1 = [G.eval_expr(meta["hom"], m, self._col(self.w1), self.c1 + [self.u1], r, rows) for r in range(rows)]
        self.r = rng.randrange(m)

    def _col(self, w):
        nf, rows = self.meta["num_fixed"], self.rows
        return lambda index, row: self.fixed[index][row] if index < nf else w[(index - nf) * rows + row]

    # what the reference computes on the CPU, with Python integers (independent of oracle/ and of the GPU)
    def incoming_gate_values(self):
        return [G.eval_expr(self.meta["gate"], self.m, self._col(self.w2), self.c2, r, self.rows) for r in range(self.rows)]

    def folded(self):
        m, r = self.m, self.r
        w = [(a + r * b) % m for a, b in zip(self.w1, self.w2)]
        u = (self.u1 + r) % m
        c = [(a + r * b) % m for a, b in zip(self.c1, self.c2)]
        return w, u, c

    def hom_on(self, w, c, u):
        return [G.eval_expr(self.meta["hom"], self.m, self._col(w), c + [u], r, self.rows) for r in range(self.rows)]

    def cross_term_challenges(self):
        return self.c1 + [self.u1] + self.c2 + [1]

    def bytes(self):
        m = self.m
        return {"fixed": [mont(c, m) for c in self.fixed], "w1": mont(self.w1, m), "w2": mont(self.w2, m), "e1": mont(self.e1, m),
                "r": mont([self.r], m), "challenges": mont(self.cross_term_challenges(), m)}

    def domain_bytes(self, w1: bytes, w2, challenges: bytes) -> dict:
        """oracle_lib.eval_rows domain over (W1s = [w1], W2s = [w2] or [])."""
        return {"row_size": self.rows, "num_advice": self.meta["num_advice"], "num_lookup": 0, "selectors": [],
                "fixed": [mont(c, self.m) for c in self.fixed], "w1": [w1], "w2": [w2] if w2 is not None else [],
                "challenges": challenges}
