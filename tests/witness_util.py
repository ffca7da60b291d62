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
