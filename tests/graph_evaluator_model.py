"""Python restatement of the reference's expression -> straight-line-program pipeline.  TEST INFRASTRUCTURE.

* `Expr` mirrors `polynomial::Expression` (/root/reference/src/polynomial/expression.rs:112-120).
* `GraphEvaluator` mirrors `GraphEvaluator::{default, new, add_rotation, add_constant, add_calculation,
  add_expression}` (src/polynomial/graph_evaluator.rs:180-351) line by line, including the CSE lookup, the
  constant-folding special cases and the `ValueSource` ordering used to canonicalise Add/Mul operands
  (`#[derive(PartialOrd)]`: variant index first, then fields).
* `GraphEvaluator.encode()` emits the u32 program encoding documented in include/mira_b200.h
  (mira_eval_program_create) — what the Rust-side `to_bytecode()` of INTEGRATION.md §4 would emit.
* `eval_expr` evaluates an `Expr` directly with Python big integers: the independent second opinion the
  reference's own tests use (graph_evaluator.rs:447-634 compare against direct field arithmetic).
* `grouped` restates `GroupedPoly::new` (src/polynomial/grouped_poly.rs:88-138): the coefficients of X^k in
  P(W1 + X*W2, ...), i.e. the cross-term expressions T_k that `commit_cross_terms` evaluates
  (src/nifs/vanilla/mod.rs:100-121); `main_gate_expr` builds the MainGate<T> custom gate
  (src/main_gate.rs:561-594) so the benchmark programs have the reference circuits' shape.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

# ------------------------------------------------------------------ Expression
CONST, POLY, CHAL, NEG, SUM, PROD, SCALED = range(7)


@dataclass(frozen=True)
class Expr:
    kind: int
    a: object = None      # Constant: int | Polynomial: (index, rotation) | Challenge: index | else: Expr
    b: object = None      # Sum/Product: Expr | Scaled: int

    def __add__(self, o): return Expr(SUM, self, o)
    def __mul__(self, o): return Expr(PROD, self, o)
    def __neg__(self): return Expr(NEG, self)
    def __sub__(self, o): return Expr(SUM, self, Expr(NEG, o))


def Constant(v: int) -> Expr: return Expr(CONST, v)
def Polynomial(index: int, rotation: int = 0) -> Expr: return Expr(POLY, (index, rotation))
def Challenge(index: int) -> Expr: return Expr(CHAL, index)
def Scaled(e: Expr, f: int) -> Expr: return Expr(SCALED, e, f)


def eval_expr(e: Expr, m: int, column, challenges, row: int, n_rows: int) -> int:
    """Direct evaluation.  column(index, row) -> int implements eval_column_var."""
    k = e.kind
    if k == CONST: return e.a % m
    if k == POLY:
        idx, rot = e.a
        return column(idx, (row + rot) % n_rows) % m
    if k == CHAL: return challenges[e.a] % m
    if k == NEG: return (-eval_expr(e.a, m, column, challenges, row, n_rows)) % m
    if k == SUM: return (eval_expr(e.a, m, column, challenges, row, n_rows) + eval_expr(e.b, m, column, challenges, row, n_rows)) % m
    if k == PROD: return (eval_expr(e.a, m, column, challenges, row, n_rows) * eval_expr(e.b, m, column, challenges, row, n_rows)) % m
    if k == SCALED: return (eval_expr(e.a, m, column, challenges, row, n_rows) * e.b) % m
    raise ValueError(k)


# ------------------------------------------------------------------ GraphEvaluator
VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_POLY, VS_CHALLENGE = range(5)
OP_ADD, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_HORNER, OP_STORE = range(8)
ValueSource = Tuple[int, int, int]          # (variant, index, rotation-index): compares like the derived PartialOrd


class GraphEvaluator:
    def __init__(self, modulus: int):
        self.m = modulus
        self.constants: List[int] = [0, 1, 2]           # Default: ZERO, ONE, from(2)
        self.rotations: List[int] = []
        self.num_intermediates = 0
        self.calculations: List[Tuple[tuple, int]] = []  # (calculation, target)

    @classmethod
    def new(cls, expr: Expr, modulus: int) -> "GraphEvaluator":
        s = cls(modulus)
        vs = s.add_expression(expr)
        s.add_calculation((OP_STORE, vs))
        return s

    def add_rotation(self, rot: int) -> int:
        if rot in self.rotations:
            return self.rotations.index(rot)
        self.rotations.append(rot)
        return len(self.rotations) - 1

    def add_constant(self, c: int) -> ValueSource:
        c %= self.m
        if c in self.constants:
            return (VS_CONSTANT, self.constants.index(c), 0)
        self.constants.append(c)
        return (VS_CONSTANT, len(self.constants) - 1, 0)

    def add_calculation(self, calc: tuple) -> ValueSource:
        for c, target in self.calculations:
            if c == calc:
                return (VS_INTERMEDIATE, target, 0)
        target = self.num_intermediates
        self.calculations.append((calc, target))
        self.num_intermediates += 1
        return (VS_INTERMEDIATE, target, 0)

    def add_expression(self, e: Expr) -> ValueSource:
        C0, C1, C2 = (VS_CONSTANT, 0, 0), (VS_CONSTANT, 1, 0), (VS_CONSTANT, 2, 0)
        k = e.kind
        if k == CONST:
            return self.add_constant(e.a)
        if k == POLY:
            idx, rot = e.a
            r = self.add_rotation(rot)
            return self.add_calculation((OP_STORE, (VS_POLY, idx, r)))
        if k == CHAL:
            return self.add_calculation((OP_STORE, (VS_CHALLENGE, e.a, 0)))
        if k == NEG:
            if e.a.kind == CONST:
                return self.add_constant(-e.a.a)
            ra = self.add_expression(e.a)
            return ra if ra == C0 else self.add_calculation((OP_NEGATE, ra))
        if k == SUM:
            if e.b.kind == NEG:                       # undo subtraction stored as a + (-b)
                ra = self.add_expression(e.a)
                rb = self.add_expression(e.b.a)
                if ra == C0:
                    return self.add_calculation((OP_NEGATE, rb))
                if rb == C0:
                    return ra
                return self.add_calculation((OP_SUB, ra, rb))
            ra = self.add_expression(e.a)
            rb = self.add_expression(e.b)
            return self.add_calculation((OP_ADD, ra, rb) if ra <= rb else (OP_ADD, rb, ra))
        if k == PROD:
            ra = self.add_expression(e.a)
            rb = self.add_expression(e.b)
            if ra == C0 or rb == C0: return C0
            if ra == C1: return rb
            if rb == C1: return ra
            if ra == C2: return self.add_calculation((OP_DOUBLE, rb))
            if rb == C2: return self.add_calculation((OP_DOUBLE, ra))
            if ra == rb: return self.add_calculation((OP_SQUARE, ra))
            return self.add_calculation((OP_MUL, ra, rb) if ra <= rb else (OP_MUL, rb, ra))
        if k == SCALED:
            f = e.b % self.m
            if f == 0: return C0
            if f == 1: return self.add_expression(e.a)
            cst = self.add_constant(f)
            ra = self.add_expression(e.a)
            return self.add_calculation((OP_MUL, ra, cst))
        raise ValueError(k)

    # ---- serialisation (include/mira_b200.h) ---------------------------------------------------
    def encode(self) -> dict:
        code: List[int] = []
        for calc, target in self.calculations:
            op, operands = calc[0], calc[1:]
            code.append(op | (len(operands) << 8))
            code.append(target)
            for (kind, index, rot) in operands:
                code.append(kind | (rot << 8))
                code.append(index)
        return {"code": code, "constants": list(self.constants), "rotations": list(self.rotations),
                "num_intermediates": self.num_intermediates}

    def counts(self) -> Dict[str, int]:
        c = {"nodes": len(self.calculations), "mul": 0, "add": 0, "load": 0}
        for calc, _ in self.calculations:
            if calc[0] in (OP_MUL, OP_SQUARE): c["mul"] += 1
            elif calc[0] in (OP_ADD, OP_SUB, OP_DOUBLE, OP_NEGATE): c["add"] += 1
            elif calc[0] == OP_STORE and calc[1][0] in (VS_POLY, VS_FIXED): c["load"] += 1
        return c


# ------------------------------------------------------------------ GroupedPoly (cross-term expressions)
def _poly_add(p: List[Optional[Expr]], q: List[Optional[Expr]]) -> List[Optional[Expr]]:
    out: List[Optional[Expr]] = []
    for i in range(max(len(p), len(q))):
        a = p[i] if i < len(p) else None
        b = q[i] if i < len(q) else None
        out.append(a if b is None else (b if a is None else a + b))
    return out


def _poly_mul(p, q):
    out: List[Optional[Expr]] = [None] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        if a is None: continue
        for j, b in enumerate(q):
            if b is None: continue
            t = a * b
            out[i + j] = t if out[i + j] is None else out[i + j] + t
    return out


def grouped(e: Expr, is_folded_var, shift_poly, shift_challenge) -> List[Optional[Expr]]:
    """Coefficients (index = power of X) of e with every folded variable v replaced by v1 + X*v2.
    is_folded_var(index) says whether a Polynomial query is a witness column (fixed/selectors are not
    folded); shift_poly / shift_challenge give the second instance's index for a folded variable."""
    k = e.kind
    if k == CONST: return [e]
    if k == POLY:
        idx, rot = e.a
        return [e, Polynomial(shift_poly(idx), rot)] if is_folded_var(idx) else [e]
    if k == CHAL: return [e, Challenge(shift_challenge(e.a))]
    if k == NEG: return [None if c is None else -c for c in grouped(e.a, is_folded_var, shift_poly, shift_challenge)]
    if k == SUM: return _poly_add(grouped(e.a, is_folded_var, shift_poly, shift_challenge), grouped(e.b, is_folded_var, shift_poly, shift_challenge))
    if k == PROD: return _poly_mul(grouped(e.a, is_folded_var, shift_poly, shift_challenge), grouped(e.b, is_folded_var, shift_poly, shift_challenge))
    if k == SCALED: return [None if c is None else Scaled(c, e.b) for c in grouped(e.a, is_folded_var, shift_poly, shift_challenge)]
    raise ValueError(k)


def degree(e: Expr, is_folded_var) -> int:
    k = e.kind
    if k == CONST: return 0
    if k == POLY: return 1 if is_folded_var(e.a[0]) else 0
    if k == CHAL: return 1
    if k in (NEG, SCALED): return degree(e.a, is_folded_var)
    if k == SUM: return max(degree(e.a, is_folded_var), degree(e.b, is_folded_var))
    return degree(e.a, is_folded_var) + degree(e.b, is_folded_var)


def homogeneous(e: Expr, is_folded_var, u_index: int) -> Tuple[Expr, int]:
    """Expression::homogeneous (src/polynomial/expression.rs:356-429), bottom-up: at every Sum the side of
    lower degree is multiplied by u^(difference), u = Challenge(u_index).  Returns (expr, degree)."""
    def u_pow(d: int) -> Expr:                     # challenge_in_degree
        x = Challenge(u_index)
        for _ in range(d - 1):
            x = x * Challenge(u_index)
        return x

    k = e.kind
    if k == CONST: return e, 0
    if k == POLY: return e, (1 if is_folded_var(e.a[0]) else 0)
    if k == CHAL: return e, 1
    if k == NEG:
        x, d = homogeneous(e.a, is_folded_var, u_index)
        return Expr(NEG, x), d
    if k == SCALED:
        x, d = homogeneous(e.a, is_folded_var, u_index)
        return Expr(SCALED, x, e.b), d
    (l, dl), (r, dr) = homogeneous(e.a, is_folded_var, u_index), homogeneous(e.b, is_folded_var, u_index)
    if k == PROD: return l * r, dl + dr
    if dl > dr: return l + (r * u_pow(dl - dr)), dl
    if dl < dr: return (l * u_pow(dr - dl)) + r, dr
    return l + r, dl


def main_gate_expr(T: int, col0: int, num_selectors: int, num_fixed_total: int, fixed0: int) -> Expr:
    """The MainGate<T> gate (src/main_gate.rs:561-594).  Column indices follow eval_column_var's order
    (selectors, then fixed, then advice): fixed column j of this gate is `num_selectors + fixed0 + j`, advice
    column j is `num_selectors + num_fixed_total + col0 + j`."""
    fx = lambda j: Polynomial(num_selectors + fixed0 + j)
    ad = lambda j: Polynomial(num_selectors + num_fixed_total + col0 + j)
    state = [ad(i) for i in range(T)]
    inp, out = ad(T), ad(T + 1)
    q_1 = [fx(i) for i in range(T)]
    q_5 = [fx(T + i) for i in range(T)]
    q_m = [fx(2 * T + i) for i in range(T)]
    q_i, q_o, rc = fx(3 * T), fx(3 * T + 1), fx(3 * T + 2)

    def pow_5(v):
        v2 = v * v
        return v2 * v2 * v

    init = q_m[0] * state[0] * state[1] + q_i * inp + rc + q_o * out
    if T >= 4:
        init = q_m[1] * state[2] * state[3] + init
    acc = init
    for s, q1, q5 in zip(state, q_1, q_5):
        acc = acc + (q1 * s + q5 * pow_5(s))
    return acc


def relaxed_circuit(T: int, n_gates: int, modulus: int):
    """Everything `is_sat_relaxed` (src/plonk/mod.rs:495-560) and `commit_cross_terms` need for a circuit shaped like
    the reference's IVC circuits: the compressed gate expression, its homogeneous form over ONE instance (challenges
    followed by u), the GraphEvaluator of that form, the cross-term programs and the column bookkeeping."""
    progs, meta = cross_term_programs(T, n_gates, modulus)
    hom = meta["hom"]
    meta = dict(meta)
    meta["hom_program"] = GraphEvaluator.new(hom, modulus)     # GraphEvaluator::new(custom_gates_lookup_compressed.homogeneous())
    return progs, meta


def cross_term_programs(T: int, n_gates: int, modulus: int):
    """Programs shaped like the reference's IVC circuits: `n_gates` MainGate<T> instances
    (1 for the secondary circuit, 2 for the primary), compressed with a challenge when n_gates > 1
    (src/plonk/util.rs:97-117), homogenised with u, grouped by powers of X.  Returns
    (programs, meta) where programs[k-1] is the GraphEvaluator of cross term T_k."""
    num_selectors = 0
    fixed_per = 3 * T + 3
    adv_per = T + 2
    num_fixed = fixed_per * n_gates
    num_advice = adv_per * n_gates
    gates = [main_gate_expr(T, g * adv_per, num_selectors, num_fixed, g * fixed_per) for g in range(n_gates)]
    n_ch = 1 if n_gates > 1 else 0
    expr = gates[0]
    if n_gates > 1:                              # compress_expression: fold(0, |acc, e| e + acc * y)
        acc = Constant(0)
        for g in gates:
            acc = g + (acc * Challenge(0))
        expr = acc
    first_adv = num_selectors + num_fixed
    is_folded = lambda idx: idx >= first_adv
    hom, _deg = homogeneous(expr, is_folded, n_ch)   # u = Challenge(num_challenges): per instance [c..., u]
    per_instance = n_ch + 1
    coeffs = grouped(hom, is_folded, lambda idx: idx + num_advice, lambda ci: ci + per_instance)
    progs = [GraphEvaluator.new(c if c is not None else Constant(0), modulus) for c in coeffs[1:]]
    meta = {"num_selectors": num_selectors, "num_fixed": num_fixed, "num_advice": num_advice, "num_lookup": 0,
            "num_challenges": 2 * per_instance, "degree": len(coeffs) - 1, "exprs": coeffs[1:],
            "gate": expr, "hom": hom, "per_instance_challenges": per_instance, "fixed_per_gate": fixed_per,
            "advice_per_gate": adv_per, "T": T, "n_gates": n_gates}
    return progs, meta
