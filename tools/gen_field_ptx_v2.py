#!/usr/bin/env python3
"""Generator v2 for the 8x32-bit-limb field multiplier: SEPARATED wide product and Montgomery reduction.

v1 (tools/gen_field_ptx.py) interleaves a*b_i rows with reduction rows: 128 wide MACs + 8 IMAD per product, always.
v2 splits the work so that wide MACs — the binding resource on B200 (IMAD.WIDE.U32 issues at 32 lanes/clk/SM, the
ALU pipe that takes IADD3/LOP3 is mostly idle) — can be saved:

  mul_wide   T = a*b, 16 limbs.  One level of subtractive Karatsuba on 4-limb halves: 3 x 16 = 48 wide MACs
             (+ ~75 ALU ops) instead of 64.
  sqr_wide   T = a^2: 28 cross products, one funnel-shift doubling, 8 diagonal products = 36 wide MACs.
  redc_F     word-serial Montgomery reduction of a 16-limb T < p*2^256 to 8 limbs < p: 64 wide MACs + 8 IMAD.
  wsub_F     D = T1 - T2 (+ p*2^256 if negative), 16 limbs: lets a difference of two products share ONE reduction
             (XYZZ mixed add: y3 = r*(qq - x3) - y*ppp).

Every partial product is a mad.lo.cc/madc.hi.cc pair on an even-aligned register pair (two accumulators, one for
even and one for odd limb positions) that ptxas fuses into a single IMAD.WIDE.U32[.X].

As in v1 there are two back ends over ONE op list: `simulate` (explicit carry flag, Python ints) drives the
self-test against big-integer arithmetic; `emit` renders inline PTX into mira_b200/csrc/field_gen_v2.cuh.

Run:  python tools/gen_field_ptx_v2.py
"""
from __future__ import annotations

import os
import random

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FIELDS = {"Fq": P, "Fr": R}
W = 1 << 32
M32 = W - 1


def limbs(v, n):
    return [(v >> (32 * i)) & M32 for i in range(n)]


class Prog:
    def __init__(self):
        self.ops = []
        self.tmp = 0

    def reg(self, prefix="t"):
        self.tmp += 1
        return f"{prefix}{self.tmp}"

    def op(self, name, dst, *src):
        self.ops.append((name, dst, src))


# ------------------------------------------------------------------ building blocks
class EvenOdd:
    """Two accumulators over limb positions: E holds words at positions 0,1,2,... (pairs start at even positions),
    O holds words at positions 1,2,3,... (pairs start at odd positions).  None = known zero."""

    def __init__(self, p: Prog, width: int):
        self.p = p
        self.width = width
        self.E = [None] * (width + 2)
        self.O = [None] * (width + 2)      # O[k] sits at position k + 1

    def chain(self, pairs):
        """pairs: list of (position, x, y) with positions increasing by 2: adds x*y at each position, one carry chain."""
        p = self.p
        if not pairs:
            return
        pos0 = pairs[0][0]
        acc, base = (self.E, 0) if pos0 % 2 == 0 else (self.O, 1)
        carry = False
        last_w = None
        for (pos, x, y) in pairs:
            w = pos - base
            assert last_w is None or w == last_w + 2
            for half, ww in (("lo", w), ("hi", w + 1)):
                cur = acc[ww]
                if cur is None and not carry:
                    dst = p.reg("m")
                    p.op(f"mul.{half}.u32", dst, x, y)
                else:
                    dst = cur if cur is not None else p.reg("m")
                    p.op(("madc." if carry else "mad.") + half + ".cc.u32", dst, x, y, cur if cur is not None else 0)
                    carry = True
                acc[ww] = dst
            last_w = w
        top = last_w + 2
        if carry and top + base < self.width:      # a carry out of the last pair ripples into the next word
            cur = acc[top]
            dst = cur if cur is not None else p.reg("c")
            # words above the chain's last pair are untouched by earlier rows or hold at most a previous carry bit
            p.op("addc.u32", dst, cur if cur is not None else 0, 0)
            acc[top] = dst

    def merge(self, n_words):
        """V = E + (O << 32), n_words limbs, fully carried.  Returns the register (or 0) of every limb."""
        p = self.p
        out = []
        carry = False
        for k in range(n_words):
            e = self.E[k]
            o = self.O[k - 1] if k >= 1 else None
            if e is None and o is None and not carry:
                out.append(0)
                continue
            if o is None and not carry:
                out.append(e)
                continue
            if e is None and not carry:
                out.append(o)
                continue
            dst = p.reg("v")
            last = k == n_words - 1
            name = ("addc" if carry else "add") + ("" if last else ".cc") + ".u32"
            p.op(name, dst, e if e is not None else 0, o if o is not None else 0)
            carry = not last
            out.append(dst)
        return out


def wide_product(p: Prog, a, b, skip=None):
    """Schoolbook product of two limb lists with the even/odd technique.  skip(i_a, i_b) -> True drops a product."""
    n, m = len(a), len(b)
    eo = EvenOdd(p, n + m)
    for j in range(m):
        for parity in (0, 1):
            pairs = [(j + i, a[i], b[j]) for i in range(parity, n, 2) if not (skip and skip(i, j))]
            # split where positions are not contiguous by 2 (only happens with skip)
            run = []
            for t in pairs:
                if run and t[0] != run[-1][0] + 2:
                    eo.chain(run)
                    run = []
                run.append(t)
            eo.chain(run)
    return eo


def materialize(p: Prog, regs):
    """Replace literal zeros by registers holding 0 (outputs must be registers)."""
    out = []
    for r in regs:
        if r == 0:
            z = p.reg("z")
            p.op("mov.u32", z, 0)
            out.append(z)
        else:
            out.append(r)
    return out


def add_chain(p: Prog, xs, ys, carry_in=False, keep_carry=False, prefix="s", wrap=False):
    """limb-wise xs + ys (equal length; entries may be 0).  Returns result regs; CC carries out if keep_carry;
    wrap: the sum is taken modulo 2^(32 n) on purpose (two's-complement subtraction)."""
    out = []
    n = len(xs)
    for k in range(n):
        dst = p.reg(prefix)
        first = k == 0 and not carry_in
        last = k == n - 1 and not keep_carry
        name = ("add" if first else "addc") + ("" if last else ".cc") + ".u32"
        if last and wrap:
            name = "addc.u32_wrap"
        p.op(name, dst, xs[k], ys[k])
        out.append(dst)
    return out


def sub_chain(p: Prog, xs, ys, prefix="d"):
    """limb-wise xs - ys; returns (regs, borrow_mask_reg) where mask = 0xffffffff if xs < ys."""
    out = []
    for k in range(len(xs)):
        dst = p.reg(prefix)
        p.op("sub.cc.u32" if k == 0 else "subc.cc.u32", dst, xs[k], ys[k])
        out.append(dst)
    m = p.reg("w")
    p.op("subc.u32", m, 0, 0)
    return out, m


def abs_diff(p: Prog, hi, lo):
    """|hi - lo| on 4 limbs; returns (regs, sign_mask) with mask all-ones when hi < lo."""
    d, m = sub_chain(p, hi, lo)
    x = []
    for k in range(len(d)):
        t = p.reg("x")
        p.op("xor.b32", t, d[k], m)
        x.append(t)
    out = []
    for k in range(len(x)):          # (d ^ m) - m
        t = p.reg("g")
        p.op("sub.cc.u32" if k == 0 else ("subc.cc.u32" if k < len(x) - 1 else "subc.u32"), t, x[k], m)
        out.append(t)
    return out, m


def build_mul_wide(karatsuba=True) -> Prog:
    """t0..t15 = a0..a7 * b0..b7."""
    p = Prog()
    a = [f"a{i}" for i in range(8)]
    b = [f"b{i}" for i in range(8)]
    if not karatsuba:
        eo = wide_product(p, a, b)
        t = materialize(p, eo.merge(16))
        for k in range(16):
            p.op("mov.u32", f"r{k}", t[k])
        return p
    a0, a1, b0, b1 = a[:4], a[4:], b[:4], b[4:]
    z0 = wide_product(p, a0, b0).merge(8)
    z2 = wide_product(p, a1, b1).merge(8)
    da, sa = abs_diff(p, a1, a0)
    db, sb = abs_diff(p, b1, b0)
    D = wide_product(p, da, db).merge(8)
    # z1 = z0 + z2 - (a1 - a0)(b1 - b0) = z0 + z2 + (neg ? +D : -D),  neg = sa ^ sb
    S = add_chain(p, z0, z2, keep_carry=True)
    s8 = p.reg("s")
    p.op("addc.u32", s8, 0, 0)
    sub_mask = p.reg("k")            # all-ones when D must be SUBTRACTED (signs equal)
    p.op("xor.b32", sub_mask, sa, sb)
    p.op("not.b32", sub_mask, sub_mask)
    Dx = []
    for k in range(8):
        t = p.reg("y")
        p.op("xor.b32", t, D[k], sub_mask)
        Dx.append(t)
    junk = p.reg("j")
    p.op("add.cc.u32", junk, sub_mask, sub_mask)          # carry = 1 iff subtracting (two's complement +1)
    z1 = add_chain(p, S + [s8], Dx + [sub_mask], carry_in=True, wrap=True)
    # T = z0 + z1 << 128 + z2 << 256
    lo = z0[:4]
    mid = add_chain(p, z0[4:] + z2[:4] + [z2[4]], z1[:9], keep_carry=True)
    top = []
    for k in range(5, 8):
        t = p.reg("u")
        p.op("addc.cc.u32" if k < 7 else "addc.u32", t, z2[k], 0)
        top.append(t)
    T = lo + mid + top
    for k in range(16):
        p.op("mov.u32", f"r{k}", T[k])
    return p


def build_sqr_wide() -> Prog:
    """t0..t15 = (a0..a7)^2: cross products once, doubled by a funnel shift, plus the diagonal."""
    p = Prog()
    a = [f"a{i}" for i in range(8)]
    eo = wide_product(p, a, a, skip=lambda i, j: i >= j)     # a_i * a_j for i < j
    c = eo.merge(16)                                         # < 2^511
    # doubled[k] = (c[k] << 1) | (c[k-1] >> 31)
    dbl = []
    for k in range(16):
        t = p.reg("h")
        lo = c[k - 1] if k >= 1 else 0
        p.op("shf.l.clamp.b32", t, lo, c[k], 1)
        dbl.append(t)
    # + diagonal: a_i^2 at position 2i, one carry chain over all 8 pairs
    out = []
    for i in range(8):
        for half, w in (("lo", 2 * i), ("hi", 2 * i + 1)):
            t = p.reg("q")
            first = i == 0 and half == "lo"
            last = i == 7 and half == "hi"
            name = ("mad." if first else "madc.") + half + ("" if last else ".cc") + ".u32"
            p.op(name, t, a[i], a[i], dbl[w])
            out.append(t)
    for k in range(16):
        p.op("mov.u32", f"r{k}", out[k])
    return p


def build_redc(mod: int) -> Prog:
    """r0..r7 = (t0..t15) / 2^256 mod `mod`, for T < mod * 2^256; output < mod."""
    p = Prog()
    n0inv = (-pow(mod, -1, W)) % W
    ml = limbs(mod, 8)
    m_even, m_odd = ml[0::2], ml[1::2]
    t = [f"a{i}" for i in range(16)]          # inputs named a0..a15 for the emitter
    ev = [p.reg("e") for _ in range(9)]
    for k in range(8):
        p.op("mov.u32", ev[k], t[k])
    p.op("mov.u32", ev[8], 0)
    od = None
    for i in range(8):
        if i > 0:
            # divide by 2^32: ev[0] == 0, ev[1] is a lone word at new position 0, ev[2..8] is the new odd row,
            # the old odd row becomes the even row
            pend = ev[1]
            new_od = [p.reg("o") for _ in range(9)]
            for k in range(7):
                p.op("mov.u32", new_od[k], ev[2 + k])
            p.op("mov.u32", new_od[7], 0)
            p.op("mov.u32", new_od[8], 0)
            ev, od = od, new_od
            p.op("add.cc.u32", ev[0], ev[0], pend)       # carry lands on position 1 = first word of the odd chain
        else:
            od = [p.reg("o") for _ in range(9)]
            for k in range(9):
                p.op("mov.u32", od[k], 0)
        m = p.reg("m")
        p.op("mul.lo.u32", m, ev[0], n0inv)
        first = True
        for k in range(4):                                # od += m * p_odd  (positions 1,3,5,7)
            for half, w in (("lo", 2 * k), ("hi", 2 * k + 1)):
                name = ("madc." if (not first or i > 0) else "mad.") + half + ".cc.u32"
                p.op(name, od[w], m_odd[k], m, od[w])
                first = False
        p.op("addc.u32", od[8], od[8], 0)
        for k in range(4):                                # ev += m * p_even (positions 0,2,4,6)
            for half, w in (("lo", 2 * k), ("hi", 2 * k + 1)):
                name = ("mad." if (k == 0 and half == "lo") else "madc.") + half + ".cc.u32"
                p.op(name, ev[w], m_even[k], m, ev[w])
        p.op("addc.u32", ev[8], ev[8], 0)
    # (T_lo + M*p) / 2^256 = ev[1] + od + ev[2..8]*W   then + T_hi
    v = []
    for k in range(8):
        d = p.reg("v")
        if k == 0:
            p.op("add.cc.u32", d, od[0], ev[1])
        else:
            p.op("addc.cc.u32" if k < 7 else "addc.u32", d, od[k], ev[k + 1])
        v.append(d)
    u = add_chain(p, v, t[8:], prefix="u")                # < 2*mod < 2^255: no carry out
    s, brw = sub_chain(p, u, ml, prefix="s")
    for k in range(8):
        p.op("selp_nz", f"r{k}", u[k], s[k], brw)
    return p


def build_wsub(mod: int) -> Prog:
    """r0..r15 = a0..a15 - b0..b15, plus mod*2^256 when the difference is negative (inputs < mod*2^256)."""
    p = Prog()
    ml = limbs(mod, 8)
    d, m = sub_chain(p, [f"a{i}" for i in range(16)], [f"b{i}" for i in range(16)])
    for k in range(8):
        p.op("mov.u32", f"r{k}", d[k])
    msk = []
    for k in range(8):
        t = p.reg("k")
        p.op("and.b32", t, m, ml[k])
        msk.append(t)
    for k in range(8):
        name = ("add.cc.u32" if k == 0 else ("addc.cc.u32" if k < 7 else "addc.u32_wrap"))
        p.op(name, f"r{8 + k}", d[8 + k], msk[k])
    return p


# ------------------------------------------------------------------ simulator
def simulate(prog: Prog, inputs: dict) -> dict:
    regs = dict(inputs)
    cc = 0

    def val(x):
        return x if isinstance(x, int) else regs[x]

    for name, dst, src in prog.ops:
        s = [val(x) for x in src]
        if name == "mov.u32":
            regs[dst] = s[0]
        elif name == "mul.lo.u32":
            regs[dst] = (s[0] * s[1]) & M32
        elif name == "mul.hi.u32":
            regs[dst] = (s[0] * s[1]) >> 32
        elif name.startswith("mad") and (".lo" in name or ".hi" in name):
            prod = s[0] * s[1]
            part = (prod & M32) if ".lo" in name else (prod >> 32)
            cin = cc if name.startswith("madc") else 0
            tot = part + s[2] + cin
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = tot >> 32
            else:
                assert tot >> 32 == 0, f"carry lost out of {name}"
        elif name in ("add.cc.u32", "addc.cc.u32", "addc.u32", "add.u32"):
            cin = cc if name.startswith("addc") else 0
            tot = s[0] + s[1] + cin
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = tot >> 32
            else:
                assert tot >> 32 == 0, f"carry lost out of {name} -> {dst}"
        elif name == "addc.u32_wrap":
            regs[dst] = (s[0] + s[1] + cc) & M32
        elif name in ("sub.cc.u32", "subc.cc.u32", "subc.u32"):
            bin_ = cc if name.startswith("subc") else 0
            tot = s[0] - s[1] - bin_
            regs[dst] = tot & M32
            if ".cc" in name:
                cc = 1 if tot < 0 else 0
        elif name == "selp_nz":
            regs[dst] = s[0] if s[2] != 0 else s[1]
        elif name == "and.b32":
            regs[dst] = s[0] & s[1]
        elif name == "xor.b32":
            regs[dst] = s[0] ^ s[1]
        elif name == "not.b32":
            regs[dst] = (~s[0]) & M32
        elif name == "shf.l.clamp.b32":       # d = ((hi:lo) << n) >> 32, inputs (lo, hi, n)
            n = min(s[2], 32)
            regs[dst] = (((s[1] << 32) | s[0]) << n >> 32) & M32
        else:
            raise ValueError(name)
    return regs


def _inp(prefix, v, n):
    return {f"{prefix}{i}": x for i, x in enumerate(limbs(v, n))}


def _out(regs, n):
    return sum(regs[f"r{i}"] << (32 * i) for i in range(n))


def count(prog: Prog):
    wide = sum(1 for o in prog.ops if ".lo" in o[0] and (o[0].startswith("mad") or o[0].startswith("mul.lo")))
    hi = sum(1 for o in prog.ops if ".hi" in o[0])
    alu = sum(1 for o in prog.ops if o[0].split(".")[0] in ("add", "addc", "sub", "subc", "xor", "and", "not", "shf", "selp_nz"))
    return {"ops": len(prog.ops), "wide_mac": hi, "lo_only": wide - hi, "alu": alu}


def selftest():
    rng = random.Random(20261018)
    edge256 = [0, 1, 2, W - 1, W, (1 << 128) - 1, 1 << 128, (1 << 128) + 1, (1 << 255), (1 << 256) - 1, (1 << 256) - W,
               int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), ((1 << 128) - 1) << 128]
    for kara in (False, True):
        prog = build_mul_wide(kara)
        cases = [(x, y) for x in edge256 for y in edge256] + [(rng.randrange(1 << 256), rng.randrange(1 << 256)) for _ in range(3000)]
        # halves ordered both ways so that every sign combination of the Karatsuba differences occurs
        for _ in range(500):
            lo, hi = sorted((rng.randrange(1 << 128), rng.randrange(1 << 128)))
            lo2, hi2 = sorted((rng.randrange(1 << 128), rng.randrange(1 << 128)))
            cases += [((hi << 128) | lo, (hi2 << 128) | lo2), ((lo << 128) | hi, (hi2 << 128) | lo2), ((lo << 128) | hi, (lo2 << 128) | hi2),
                      ((hi << 128) | hi, (lo2 << 128) | hi2)]
        for x, y in cases:
            out = simulate(prog, {**_inp("a", x, 8), **_inp("b", y, 8)})
            assert _out(out, 16) == x * y, (kara, hex(x), hex(y))
        print(f"selftest mul_wide karatsuba={kara}: {len(cases)} cases ok; {count(prog)}")
    prog = build_sqr_wide()
    cases = edge256 + [rng.randrange(1 << 256) for _ in range(4000)]
    for x in cases:
        out = simulate(prog, _inp("a", x, 8))
        assert _out(out, 16) == x * x, hex(x)
    print(f"selftest sqr_wide: {len(cases)} cases ok; {count(prog)}")
    for fname, mod in FIELDS.items():
        prog = build_redc(mod)
        rinv = pow(1 << 256, -1, mod)
        tmax = mod << 256
        cases = [0, 1, mod, mod - 1, (mod - 1) ** 2, tmax - 1, tmax - mod, (1 << 256) - 1, 1 << 256, (mod - 1) << 256,
                 ((mod - 1) << 256) | ((1 << 256) - 1)]
        cases += [rng.randrange(tmax) for _ in range(4000)] + [rng.randrange(mod) * rng.randrange(mod) for _ in range(2000)]
        for t in cases:
            out = simulate(prog, _inp("a", t, 16))
            assert _out(out, 8) == t * rinv % mod, (fname, hex(t))
        print(f"selftest redc_{fname}: {len(cases)} cases ok; {count(prog)}")
        prog = build_wsub(mod)
        cases = [(0, 0), (0, tmax - 1), (tmax - 1, 0), (1, 2), (tmax - 1, tmax - 1)]
        cases += [(rng.randrange(tmax), rng.randrange(tmax)) for _ in range(3000)]
        for x, y in cases:
            out = simulate(prog, {**_inp("a", x, 16), **_inp("b", y, 16)})
            got = _out(out, 16)
            assert got == (x - y if x >= y else x - y + tmax), (fname, hex(x), hex(y))
            assert got < tmax
        print(f"selftest wsub_{fname}: {len(cases)} cases ok; {count(prog)}")


# ------------------------------------------------------------------ PTX emission
def emit_function(name: str, prog: Prog, n_out: int, n_a: int, n_b: int) -> str:
    temps, seen = [], set()

    def is_io(x):
        return x[0] in "abr" and x[1:].isdigit()

    for _, dst, src in prog.ops:
        for x in (dst,) + tuple(src):
            if isinstance(x, str) and is_io(x):
                k = int(x[1:])
                assert k < {"r": n_out, "a": n_a, "b": max(n_b, 0)}[x[0]], f"temporary {x} collides with an operand name"
            if isinstance(x, str) and not is_io(x) and x not in seen:
                seen.add(x)
                temps.append(x)

    def ref(x):
        if isinstance(x, int):
            return f"0x{x:08x}"
        if is_io(x):
            k = int(x[1:])
            return f"%{k}" if x[0] == "r" else (f"%{n_out + k}" if x[0] == "a" else f"%{n_out + n_a + k}")
        return x

    lines = ["{", ".reg .u32 " + ", ".join(temps) + ";", ".reg .pred pb;"]
    for opn, dst, src in prog.ops:
        if opn == "selp_nz":
            lines.append(f"setp.ne.u32 pb, {ref(src[2])}, 0;")
            lines.append(f"selp.u32 {ref(dst)}, {ref(src[0])}, {ref(src[1])}, pb;")
        elif opn == "addc.u32_wrap":
            lines.append(f"addc.u32 {ref(dst)}, " + ", ".join(ref(x) for x in src) + ";")
        else:
            lines.append(f"{opn} {ref(dst)}, " + ", ".join(ref(x) for x in src) + ";")
    lines.append("}")
    body = "\n".join(f'      "{l}\\n\\t"' for l in lines)
    outs = ", ".join(f'"=r"(t{i})' for i in range(n_out))
    ins = ", ".join(f'"r"(a[{i}])' for i in range(n_a))
    if n_b:
        ins += ", " + ", ".join(f'"r"(b[{i}])' for i in range(n_b))
    sig = f"uint32_t (&r)[{n_out}], const uint32_t (&a)[{n_a}]" + (f", const uint32_t (&b)[{n_b}]" if n_b else "")
    decl = ", ".join(f"t{i}" for i in range(n_out))
    copy = " ".join(f"r[{i}] = t{i};" for i in range(n_out))
    return (f"__device__ __forceinline__ void {name}({sig}) {{\n  uint32_t {decl};\n  asm(\n{body}\n      : {outs}\n      : {ins});\n"
            f"  {copy}\n}}\n")


def emit_cuda(path: str):
    out = ["// GENERATED by tools/gen_field_ptx_v2.py — do not edit.",
           "// Separated wide product (Karatsuba / dedicated square) and word-serial Montgomery reduction, 8 x 32-bit limbs.",
           "#pragma once", "#include <cstdint>", "namespace mira { namespace gen2 {", ""]
    out.append(emit_function("mul_wide_kara", build_mul_wide(True), 16, 8, 8))
    out.append(emit_function("mul_wide_school", build_mul_wide(False), 16, 8, 8))
    out.append(emit_function("sqr_wide", build_sqr_wide(), 16, 8, 0))
    for fname, mod in FIELDS.items():
        out.append(emit_function(f"redc_{fname}", build_redc(mod), 8, 16, 0))
        out.append(emit_function(f"wsub_{fname}", build_wsub(mod), 16, 16, 16))
    out.append("} }  // namespace mira::gen2")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    selftest()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit_cuda(os.path.join(root, "mira_b200", "csrc", "field_gen_v2.cuh"))
