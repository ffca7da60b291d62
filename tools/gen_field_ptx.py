#!/usr/bin/env python3
"""Generator for the 8x32-bit-limb Montgomery multiplier / squarer used by the CUDA kernels.

The multiplier is an interleaved (word-serial) Montgomery product that keeps TWO accumulator
rows, one aligned at even and one at odd 32-bit word positions, so that every partial product
a_j*b_i is a `mad.lo.cc.u32 / madc.hi.cc.u32` pair on an even-aligned register pair.  ptxas
fuses each pair into a single IMAD.WIDE.U32[.X] with the carry chained through a predicate
(SURVEY.md §9.2), i.e. one integer-pipe instruction per 32x32->64 multiply-accumulate:
128 wide MACs + 8 IMAD per product.

This script has two back ends over ONE op list:
  * `simulate(ops, inputs)` executes the op list in Python with an explicit carry flag, and
    `selftest()` checks it against big-integer arithmetic on random and edge inputs — this is
    how the carry logic is validated on a box without a GPU;
  * `emit_cuda()` renders the same op list as inline PTX with the modulus limbs as immediates
    into mira_b200/csrc/field_gen.cuh.

Run:  python tools/gen_field_ptx.py            (self-test + regenerate header)
"""
from __future__ import annotations

import os
import random
import sys

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FIELDS = {"Fq": P, "Fr": R}
W = 1 << 32
M32 = W - 1
N = 8


def limbs(v, n=N):
    return [(v >> (32 * i)) & M32 for i in range(n)]


class Prog:
    """Tiny PTX-subset builder. Registers are strings; immediates are ints."""

    def __init__(self):
        self.ops = []
        self.tmp = 0

    def reg(self, prefix="t"):
        self.tmp += 1
        return f"{prefix}{self.tmp}"

    def op(self, name, dst, *src):
        self.ops.append((name, dst, src))

    # helpers -----------------------------------------------------------
    def mul_wide(self, lo, hi, a, b):
        # one mul.wide.u32 (-> IMAD.WIDE.U32) instead of a mul.lo/mul.hi pair, which ptxas does NOT fuse: it emits
        # IMAD + IMAD.HI.U32, and IMAD.HI issues slower than IMAD.WIDE on B200 (profiles/r01_intpipe_microbench.jsonl)
        self.op("mul.wide.u32", (lo, hi), a, b)

    def mad_chain(self, acc, a_list, b, carry_in=False, top=None, acc_in=None):
        """acc[2k],acc[2k+1] (+)= a_list[k]*b for k in range(len), carries chained.
        acc_in: optional per-word source list (None entries mean literal 0) when the addend
        comes from different registers than the destination (used for the shifted row)."""
        first = True
        for k, a in enumerate(a_list):
            for half, w in (("lo", 2 * k), ("hi", 2 * k + 1)):
                src = acc[w] if acc_in is None else acc_in[w]
                src = 0 if src is None else src
                if first and not carry_in:
                    name = f"mad.{half}.cc.u32"
                else:
                    name = f"madc.{half}.cc.u32"
                self.op(name, acc[w], a, b, src)
                first = False
        if top is not None:
            dst, src = top
            self.op("addc.u32", dst, 0 if src is None else src, 0)


def build_mont_mul(mod: int, square: bool = False) -> Prog:
    """r = a*b/2^256 mod `mod`, inputs < mod, output < mod.  Registers: a0..a7, b0..b7 -> r0..r7."""
    p = Prog()
    n0inv = (-pow(mod, -1, W)) % W
    ml = limbs(mod)
    a = [f"a{i}" for i in range(N)]
    b = a if square else [f"b{i}" for i in range(N)]
    a_even, a_odd = a[0::2], a[1::2]
    m_even, m_odd = ml[0::2], ml[1::2]

    ev = [p.reg("e") for _ in range(N + 1)]   # even-aligned row, word k at position k
    od = [p.reg("o") for _ in range(N + 1)]   # odd-aligned row, word k at position k+1
    m = p.reg("m")

    # ---- row 0: products written as carry chains over a zero addend.  A mul.lo/mul.hi pair (or mul.wide.u32) is
    # split by ptxas into IMAD + IMAD.HI.U32, which issues slower than the IMAD.WIDE.U32 the chain form fuses into.
    zeros = [None] * (N + 1)
    p.mad_chain(ev, a_even, b[0], acc_in=zeros, top=(ev[8], None))
    p.mad_chain(od, a_odd, b[0], acc_in=zeros, top=(od[8], None))
    p.op("mul.lo.u32", m, ev[0], n0inv)
    p.mad_chain(od, m_odd_regs(m_odd), m, top=(od[8], od[8]))
    p.mad_chain(ev, m_odd_regs(m_even), m, top=(ev[8], ev[8]))
    # ---- rows 1..7
    for i in range(1, N):
        # divide by 2^32: ev[0] is zero, ev[1] is a lone word at the new position 0,
        # ev[2..8] become the new odd-aligned row, od becomes the new even-aligned row.
        pend = ev[1]
        old_ev = ev
        new_ev = od
        new_od = [p.reg("o") for _ in range(N + 1)]
        shifted = [old_ev[2 + k] if 2 + k <= N else None for k in range(N)]   # addends for new_od[0..7]
        # new_ev[0] += pend, carry goes to position 1 == first word of the odd row chain
        p.op("add.cc.u32", new_ev[0], new_ev[0], pend)
        p.mad_chain(new_od, a_odd, b[i], carry_in=True, acc_in=shifted, top=(new_od[8], None))
        p.mad_chain(new_ev, a_even, b[i], top=(new_ev[8], new_ev[8]))
        p.op("mul.lo.u32", m, new_ev[0], n0inv)
        p.mad_chain(new_od, m_odd_regs(m_odd), m, top=(new_od[8], new_od[8]))
        p.mad_chain(new_ev, m_odd_regs(m_even), m, top=(new_ev[8], new_ev[8]))
        ev, od = new_ev, new_od
    # ---- final shift + merge: V = ev[1] + od + ev[2..8]*W
    t = [p.reg("v") for _ in range(N)]
    p.op("add.cc.u32", t[0], od[0], ev[1])
    for k in range(1, N):
        p.op("addc.cc.u32" if k < N - 1 else "addc.u32", t[k], od[k], ev[k + 1])
    # (od[8] + carry must be 0: V < 2*mod < 2^255)
    # ---- conditional subtraction of the modulus
    s = [p.reg("s") for _ in range(N)]
    brw = p.reg("w")
    p.op("sub.cc.u32", s[0], t[0], ml[0])
    for k in range(1, N):
        p.op("subc.cc.u32", s[k], t[k], ml[k])
    p.op("subc.u32", brw, 0, 0)          # 0 if no borrow (t >= mod), 0xffffffff if borrow
    for k in range(N):
        p.op("selp_nz", f"r{k}", t[k], s[k], brw)   # r = borrow ? t : s
    return p


def m_odd_regs(lst):
    return list(lst)


def build_mont_mul2(mod: int) -> Prog:
    """r = (a*b + c*d) / 2^256 mod `mod`: TWO products under ONE Montgomery reduction (192 wide MACs instead of 256).
    Inputs < mod, output < mod.  Registers: a0..a7, b0..b7, c0..c7 (named b8..b15), d0..d7 (b16..b23) -> r0..r7.
    Bound: the accumulator stays below 2*mod^2/2^256 + mod < 2*mod (mod < 2^254), so one conditional subtraction."""
    p = Prog()
    n0inv = (-pow(mod, -1, W)) % W
    ml = limbs(mod)
    a = [f"a{i}" for i in range(N)]
    b = [f"b{i}" for i in range(N)]
    c = [f"b{8 + i}" for i in range(N)]
    d = [f"b{16 + i}" for i in range(N)]
    a_even, a_odd, c_even, c_odd = a[0::2], a[1::2], c[0::2], c[1::2]
    m_even, m_odd = ml[0::2], ml[1::2]
    ev = [p.reg("e") for _ in range(N + 1)]
    od = [p.reg("o") for _ in range(N + 1)]
    m = p.reg("m")
    zeros = [None] * (N + 1)
    p.mad_chain(ev, a_even, b[0], acc_in=zeros, top=(ev[8], None))
    p.mad_chain(od, a_odd, b[0], acc_in=zeros, top=(od[8], None))
    p.mad_chain(od, c_odd, d[0], top=(od[8], od[8]))
    p.mad_chain(ev, c_even, d[0], top=(ev[8], ev[8]))
    p.op("mul.lo.u32", m, ev[0], n0inv)
    p.mad_chain(od, m_odd, m, top=(od[8], od[8]))
    p.mad_chain(ev, m_even, m, top=(ev[8], ev[8]))
    for i in range(1, N):
        pend = ev[1]
        old_ev = ev
        new_ev = od
        new_od = [p.reg("o") for _ in range(N + 1)]
        shifted = [old_ev[2 + k] if 2 + k <= N else None for k in range(N)]
        p.op("add.cc.u32", new_ev[0], new_ev[0], pend)
        p.mad_chain(new_od, a_odd, b[i], carry_in=True, acc_in=shifted, top=(new_od[8], None))
        p.mad_chain(new_ev, a_even, b[i], top=(new_ev[8], new_ev[8]))
        p.mad_chain(new_od, c_odd, d[i], top=(new_od[8], new_od[8]))
        p.mad_chain(new_ev, c_even, d[i], top=(new_ev[8], new_ev[8]))
        p.op("mul.lo.u32", m, new_ev[0], n0inv)
        p.mad_chain(new_od, m_odd, m, top=(new_od[8], new_od[8]))
        p.mad_chain(new_ev, m_even, m, top=(new_ev[8], new_ev[8]))
        ev, od = new_ev, new_od
    t = [p.reg("v") for _ in range(N)]
    p.op("add.cc.u32", t[0], od[0], ev[1])
    for k in range(1, N):
        p.op("addc.cc.u32" if k < N - 1 else "addc.u32", t[k], od[k], ev[k + 1])
    s = [p.reg("s") for _ in range(N)]
    brw = p.reg("w")
    p.op("sub.cc.u32", s[0], t[0], ml[0])
    for k in range(1, N):
        p.op("subc.cc.u32", s[k], t[k], ml[k])
    p.op("subc.u32", brw, 0, 0)
    for k in range(N):
        p.op("selp_nz", f"r{k}", t[k], s[k], brw)
    return p


def build_mont_redc(mod: int) -> Prog:
    """r = a / 2^256 mod `mod` for ANY 256-bit a (Montgomery -> canonical is this with a < mod): the reduction half of
    the product alone, 64 wide MACs + 8 IMAD instead of 136.  Same even / odd accumulator rows as build_mont_mul; the
    value starts in the even-aligned row, and the word shifted out after every row re-enters through the carry-in of
    the next row's odd chain.  Bound: (a + M*mod) / 2^256 < 1 + mod, so one conditional subtraction."""
    p = Prog()
    n0inv = (-pow(mod, -1, W)) % W
    ml = limbs(mod)
    a = [f"a{i}" for i in range(N)]
    m_even, m_odd = ml[0::2], ml[1::2]
    ev = [p.reg("e") for _ in range(N + 1)]
    od = [p.reg("o") for _ in range(N + 1)]
    m = p.reg("m")
    zeros = [None] * (N + 1)
    p.op("mul.lo.u32", m, a[0], n0inv)
    p.mad_chain(ev, m_even, m, acc_in=a + [None], top=(ev[8], None))
    p.mad_chain(od, m_odd, m, acc_in=zeros, top=(od[8], None))
    for i in range(1, N):
        pend = ev[1]
        old_ev = ev
        new_ev = od
        new_od = [p.reg("o") for _ in range(N + 1)]
        shifted = [old_ev[2 + k] if 2 + k <= N else None for k in range(N)]
        lo = p.reg("l")
        p.op("add.cc.u32", lo, new_ev[0], pend)       # carry -> position 1 = first word of the odd chain
        p.op("mul.lo.u32", m, lo, n0inv)              # mul.lo leaves the carry flag alone
        p.mad_chain(new_od, m_odd, m, carry_in=True, acc_in=shifted, top=(new_od[8], None))
        p.op("mov_into", new_ev[0], lo)
        p.mad_chain(new_ev, m_even, m, top=(new_ev[8], new_ev[8]))
        ev, od = new_ev, new_od
    t = [p.reg("v") for _ in range(N)]
    p.op("add.cc.u32", t[0], od[0], ev[1])
    for k in range(1, N):
        p.op("addc.cc.u32" if k < N - 1 else "addc.u32", t[k], od[k], ev[k + 1])
    s = [p.reg("s") for _ in range(N)]
    brw = p.reg("w")
    p.op("sub.cc.u32", s[0], t[0], ml[0])
    for k in range(1, N):
        p.op("subc.cc.u32", s[k], t[k], ml[k])
    p.op("subc.u32", brw, 0, 0)
    for k in range(N):
        p.op("selp_nz", f"r{k}", t[k], s[k], brw)
    return p


def build_add(mod: int) -> Prog:
    """r = a + b mod `mod` (inputs < mod)."""
    p = Prog()
    ml = limbs(mod)
    t = [p.reg("v") for _ in range(N)]
    s = [p.reg("s") for _ in range(N)]
    brw = p.reg("w")
    p.op("add.cc.u32", t[0], "a0", "b0")
    for k in range(1, N):
        p.op("addc.cc.u32" if k < N - 1 else "addc.u32", t[k], f"a{k}", f"b{k}")
    p.op("sub.cc.u32", s[0], t[0], ml[0])
    for k in range(1, N):
        p.op("subc.cc.u32", s[k], t[k], ml[k])
    p.op("subc.u32", brw, 0, 0)
    for k in range(N):
        p.op("selp_nz", f"r{k}", t[k], s[k], brw)
    return p


def build_sub(mod: int) -> Prog:
    """r = a - b mod `mod` (inputs < mod)."""
    p = Prog()
    ml = limbs(mod)
    t = [p.reg("v") for _ in range(N)]
    brw = p.reg("w")
    p.op("sub.cc.u32", t[0], "a0", "b0")
    for k in range(1, N):
        p.op("subc.cc.u32", t[k], f"a{k}", f"b{k}")
    p.op("subc.u32", brw, 0, 0)           # all-ones if a < b
    msk = [p.reg("k") for _ in range(N)]
    for k in range(N):
        p.op("and.b32", msk[k], brw, ml[k])
    p.op("add.cc.u32", "r0", t[0], msk[0])
    for k in range(1, N):
        p.op("addc.cc.u32" if k < N - 1 else "addc.u32_wrap", f"r{k}", t[k], msk[k])
    return p


# ------------------------------------------------------------------ simulator
def simulate(prog: Prog, inputs: dict) -> dict:
    regs = dict(inputs)
    cc = 0

    def val(x):
        return x if isinstance(x, int) else regs[x]

    for name, dst, src in prog.ops:
        s = [val(x) for x in src]
        if name == "mul.wide.u32":
            regs[dst[0]] = (s[0] * s[1]) & M32
            regs[dst[1]] = (s[0] * s[1]) >> 32
        elif name == "mul.lo.u32":
            regs[dst] = (s[0] * s[1]) & M32
        elif name == "mul.hi.u32":
            regs[dst] = (s[0] * s[1]) >> 32
        elif name in ("mad.lo.cc.u32", "madc.lo.cc.u32", "mad.hi.cc.u32", "madc.hi.cc.u32", "madc.hi.u32", "madc.lo.u32"):
            prod = s[0] * s[1]
            part = (prod & M32) if ".lo" in name else (prod >> 32)
            cin = cc if name.startswith("madc") else 0
            tot = part + s[2] + cin
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = tot >> 32
        elif name in ("add.cc.u32", "addc.cc.u32", "addc.u32"):
            cin = cc if name.startswith("addc") else 0
            tot = s[0] + s[1] + cin
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = tot >> 32
            else:
                assert tot >> 32 == 0, "carry lost out of addc.u32"
        elif name in ("sub.cc.u32", "subc.cc.u32", "subc.u32"):
            bin_ = cc if name.startswith("subc") else 0     # PTX: CC.CF holds the borrow
            tot = s[0] - s[1] - bin_
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = 1 if tot < 0 else 0
        elif name == "selp_nz":
            regs[dst] = s[0] if s[2] != 0 else s[1]
        elif name == "and.b32":
            regs[dst] = s[0] & s[1]
        elif name == "mov_into":
            regs[dst] = s[0]
        elif name == "addc.u32_wrap":       # top word of a sum that wraps mod 2^256 on purpose
            regs[dst] = (s[0] + s[1] + cc) & M32
        else:
            raise ValueError(name)
    return regs


def selftest():
    rng = random.Random(20261018)
    for fname, mod in FIELDS.items():
        for square in (False, True):
            prog = build_mont_mul(mod, square)
            edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 254) % mod, W - 1, (W << 224) % mod,
                    int("ffffffff" * 8, 16) % mod]
            cases = [(x, y) for x in edge for y in edge]
            cases += [(rng.randrange(mod), rng.randrange(mod)) for _ in range(3000)]
            rinv = pow(1 << 256, -1, mod)
            for x, y in cases:
                if square:
                    y = x
                inp = {f"a{i}": v for i, v in enumerate(limbs(x))}
                inp.update({f"b{i}": v for i, v in enumerate(limbs(y))})
                out = simulate(prog, inp)
                got = sum(out[f"r{i}"] << (32 * i) for i in range(N))
                want = x * y * rinv % mod
                assert got == want, (fname, square, hex(x), hex(y), hex(got), hex(want))
            nwide = sum(1 for o in prog.ops if o[0] in ("mul.wide.u32", "mad.lo.cc.u32", "madc.lo.cc.u32"))
            print(f"selftest {fname} square={square}: {len(cases)} cases ok; ops={len(prog.ops)} wideMACs={nwide}")


def selftest_mul2():
    rng = random.Random(99)
    for fname, mod in FIELDS.items():
        prog = build_mont_mul2(mod)
        rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, mod - 1, mod - 2, (1 << 253), (W << 224) % mod]
        cases = [(x, y, z, w) for x in edge for y in edge for z in edge for w in edge]
        cases += [tuple(rng.randrange(mod) for _ in range(4)) for _ in range(4000)]
        for x, y, z, w in cases:
            inp = {f"a{i}": v for i, v in enumerate(limbs(x))}
            inp.update({f"b{i}": v for i, v in enumerate(limbs(y))})
            inp.update({f"b{8 + i}": v for i, v in enumerate(limbs(z))})
            inp.update({f"b{16 + i}": v for i, v in enumerate(limbs(w))})
            out = simulate(prog, inp)
            got = sum(out[f"r{i}"] << (32 * i) for i in range(N))
            assert got == (x * y + z * w) * rinv % mod, (fname, hex(x), hex(y), hex(z), hex(w))
        nwide = sum(1 for o in prog.ops if o[0] in ("mul.wide.u32", "mad.lo.cc.u32", "madc.lo.cc.u32"))
        print(f"selftest {fname} mul2 (a*b + c*d): {len(cases)} cases ok; ops={len(prog.ops)} wideMACs={nwide}")


def selftest_redc():
    rng = random.Random(31)
    for fname, mod in FIELDS.items():
        prog = build_mont_redc(mod)
        rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, 2, mod - 1, mod, mod + 1, (1 << 256) - 1, (1 << 255), W - 1, (W - 1) << 224, 2 * mod, (1 << 256) - mod]
        cases = edge + [rng.randrange(mod) for _ in range(3000)] + [rng.randrange(1 << 256) for _ in range(3000)]
        for x in cases:
            out = simulate(prog, {f"a{i}": v for i, v in enumerate(limbs(x))})
            got = sum(out[f"r{i}"] << (32 * i) for i in range(N))
            assert got == x * rinv % mod, (fname, hex(x), hex(got))
        nwide = sum(1 for o in prog.ops if o[0] in ("mad.lo.cc.u32", "madc.lo.cc.u32"))
        print(f"selftest {fname} redc: {len(cases)} cases ok; ops={len(prog.ops)} wideMACs={nwide}")


def selftest_addsub():
    rng = random.Random(7)
    for fname, mod in FIELDS.items():
        pa, ps = build_add(mod), build_sub(mod)
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253), W - 1, mod >> 1, (mod >> 1) + 1]
        cases = [(x, y) for x in edge for y in edge] + [(rng.randrange(mod), rng.randrange(mod)) for _ in range(3000)]
        for x, y in cases:
            inp = {f"a{i}": v for i, v in enumerate(limbs(x))}
            inp.update({f"b{i}": v for i, v in enumerate(limbs(y))})
            o = simulate(pa, inp)
            assert sum(o[f"r{i}"] << (32 * i) for i in range(N)) == (x + y) % mod
            o = simulate(ps, inp)
            assert sum(o[f"r{i}"] << (32 * i) for i in range(N)) == (x - y) % mod
        print(f"selftest {fname} add/sub: {len(cases)} cases ok")


# ------------------------------------------------------------------ PTX emission
def emit_function(name: str, prog: Prog, square: bool, n_b: int = N) -> str:
    # collect temporaries
    temps = []
    seen = set()
    n_wide = sum(1 for o in prog.ops if o[0] == "mul.wide.u32")
    for _, dst, src in prog.ops:
        dsts = dst if isinstance(dst, tuple) else (dst,)
        for x in dsts + tuple(src):
            if isinstance(x, str) and not (x[0] in "abr" and x[1:].isdigit()) and x not in seen:
                seen.add(x)
                temps.append(x)
    # operand numbering: outputs r0..r7 = %0..%7, a = %8..%15, b = %16..%23
    def ref(x):
        if isinstance(x, int):
            return f"0x{x:08x}"
        if x[0] == "r" and x[1:].isdigit():
            return f"%{int(x[1:])}"
        if x[0] == "a" and x[1:].isdigit():
            return f"%{8 + int(x[1:])}"
        if x[0] == "b" and x[1:].isdigit():
            return f"%{16 + int(x[1:])}"
        return x
    lines = ["{", ".reg .u32 " + ", ".join(temps) + ";", ".reg .pred pb;"]
    if n_wide:
        lines.append(".reg .u64 " + ", ".join(f"wd{i}" for i in range(n_wide)) + ";")
    wide_ix = 0
    for opn, dst, src in prog.ops:
        if opn == "mul.wide.u32":
            lines.append(f"mul.wide.u32 wd{wide_ix}, {ref(src[0])}, {ref(src[1])};")
            lines.append(f"mov.b64 {{{ref(dst[0])}, {ref(dst[1])}}}, wd{wide_ix};")
            wide_ix += 1
        elif opn == "selp_nz":
            lines.append(f"setp.ne.u32 pb, {ref(src[2])}, 0;")
            lines.append(f"selp.u32 {ref(dst)}, {ref(src[0])}, {ref(src[1])}, pb;")
        elif opn == "addc.u32_wrap":
            lines.append(f"addc.u32 {ref(dst)}, " + ", ".join(ref(x) for x in src) + ";")
        elif opn == "mov_into":
            lines.append(f"mov.u32 {ref(dst)}, {ref(src[0])};")
        else:
            lines.append(f"{opn} {ref(dst)}, " + ", ".join(ref(x) for x in src) + ";")
    lines.append("}")
    body = "\n".join(f'      "{l}\\n\\t"' for l in lines)
    outs = ", ".join(f'"=r"(r[{i}])' for i in range(N))
    ins = ", ".join(f'"r"(a[{i}])' for i in range(N))
    if not square:
        ins += ", " + ", ".join(f'"r"(b[{i}])' for i in range(n_b))
    sig = "uint32_t (&r)[8], const uint32_t (&a)[8]" + ("" if square else f", const uint32_t (&b)[{n_b}]")
    return (f"__device__ __forceinline__ void {name}({sig}) {{\n"
            f"  uint32_t t0, t1, t2, t3, t4, t5, t6, t7;\n"
            f"  asm(\n{body}\n"
            f"      : \"=r\"(t0), \"=r\"(t1), \"=r\"(t2), \"=r\"(t3), \"=r\"(t4), \"=r\"(t5), \"=r\"(t6), \"=r\"(t7)\n"
            f"      : {ins});\n"
            f"  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3; r[4] = t4; r[5] = t5; r[6] = t6; r[7] = t7;\n"
            f"}}\n")


def emit_cuda(path: str):
    out = ["// GENERATED by tools/gen_field_ptx.py — do not edit.  8x32-limb Montgomery product, R = 2^256.",
           "// Each mad.lo.cc/madc.hi.cc pair is fused by ptxas into one IMAD.WIDE.U32[.X] (carry in a predicate).",
           "#pragma once", "#include <cstdint>", "namespace mira { namespace gen {", ""]
    for fname, mod in FIELDS.items():
        out.append(emit_function(f"mont_mul_{fname}", build_mont_mul(mod, False), False))
        out.append(emit_function(f"mont_sqr_{fname}", build_mont_mul(mod, True), True))
        out.append(emit_function(f"mont_mul2_{fname}", build_mont_mul2(mod), False, n_b=24))
        out.append(emit_function(f"mont_redc_{fname}", build_mont_redc(mod), True))
        out.append(emit_function(f"mod_add_{fname}", build_add(mod), False))
        out.append(emit_function(f"mod_sub_{fname}", build_sub(mod), False))
    out.append("} }  // namespace mira::gen")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    selftest()
    selftest_mul2()
    selftest_redc()
    selftest_addsub()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit_cuda(os.path.join(root, "mira_b200", "csrc", "field_gen.cuh"))
