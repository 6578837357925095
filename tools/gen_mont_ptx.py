"""Generates zk-odst_b200/csrc/mont_ptx.cuh: 256-bit Montgomery multiplication for the two Pasta
fields as inline PTX carry chains over 32-bit limbs (device code only).

Scheme: the running sum is held in two interleaved 8-limb arrays E (limb positions 0..7) and O
(positions 1..8).  Each row adds a * b_i — products of even limbs of `a` into E, of odd limbs into O,
so every 32x32->64 product lands on a (lo, hi) register pair and ptxas can fuse each
`mad.lo.cc / madc.hi.cc` pair into one IMAD.WIDE with carry — then adds m * MOD with
m = -E0 (MOD = 1 mod 2^32 for both Pasta primes, so -MOD^-1 mod 2^32 = 2^32 - 1) which clears E0;
the frame then shifts by one limb, which swaps the roles of E and O.  The primes are
2^254 + t with t < 2^126: limbs 4..6 are zero and limb 0 is 1, so the reduction needs three real
products per row instead of eight.

The instruction list is built once, EXECUTED here on Python integers against a*b*R^-1 mod p
(random and extreme operands), and only then rendered to PTX, so the emitted text is a mechanical
image of a checked sequence.

Usage: python tools/gen_mont_ptx.py [--check-only]
"""
import os
import random
import sys

P = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
Q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
MASK = 0xFFFFFFFF


def limbs32(v):
    return [(v >> (32 * i)) & MASK for i in range(8)]


def reduction_ops(E, O, mod):
    """ops adding mi * MOD to the frame (E at positions 0..7, O at 1..8); E0 becomes 0."""
    M = limbs32(mod)
    assert M[0] == 1 and M[4] == 0 and M[5] == 0 and M[6] == 0
    ops = []
    # mi = E0 * (-MOD^-1 mod 2^32) = E0 * 0xffffffff = -E0
    ops.append(("sub", "mi", 0, E[0]))
    # O += mi * (M1, M3, M5 = 0, M7) at O0, O2, O4, O6     (carry out of O7 impossible: see header)
    ops.append(("mad.lo.cc", O[0], "mi", M[1], O[0]))
    ops.append(("madc.hi.cc", O[1], "mi", M[1], O[1]))
    ops.append(("madc.lo.cc", O[2], "mi", M[3], O[2]))
    ops.append(("madc.hi.cc", O[3], "mi", M[3], O[3]))
    ops.append(("addc.cc", O[4], O[4], 0))
    ops.append(("addc.cc", O[5], O[5], 0))
    ops.append(("madc.lo.cc", O[6], "mi", M[7], O[6]))
    ops.append(("madc.hi", O[7], "mi", M[7], O[7]))
    # E += mi * (M0 = 1, M2, M4 = 0, M6 = 0) at E0, E2, E4, E6; E0 + mi = 0 mod 2^32 (never read again)
    ops.append(("add.cc", "mj", "mi", E[0]))      # only the carry matters (the sum is 0 mod 2^32)
    ops.append(("addc.cc", E[1], E[1], 0))
    ops.append(("madc.lo.cc", E[2], "mi", M[2], E[2]))
    ops.append(("madc.hi.cc", E[3], "mi", M[2], E[3]))
    ops.append(("addc.cc", E[4], E[4], 0))
    ops.append(("addc.cc", E[5], E[5], 0))
    ops.append(("addc.cc", E[6], E[6], 0))
    ops.append(("addc.cc", E[7], E[7], 0))
    ops.append(("addc", O[7], O[7], 0))
    return ops


def first_row_ops(E, O, a, bi, mod):
    ops = []
    for j in range(0, 8, 2):
        ops.append(("mul.lo", O[j], a[j + 1], bi))
        ops.append(("mul.hi", O[j + 1], a[j + 1], bi))
    for j in range(0, 8, 2):
        ops.append(("mul.lo", E[j], a[j], bi))
        ops.append(("mul.hi", E[j + 1], a[j], bi))
    return ops + reduction_ops(E, O, mod)


def row_ops(E, O, a, bi, mod):
    """E: array at positions 0..7 of the new frame (the previous O), O: the previous E (whose limb 1
    sits at position 0 and limbs 2..7 at positions 1..6)."""
    ops = []
    ops.append(("add.cc", E[0], E[0], O[1]))
    for j in range(0, 6, 2):
        ops.append(("madc.lo.cc", O[j], a[j + 1], bi, O[j + 2]))
        ops.append(("madc.hi.cc", O[j + 1], a[j + 1], bi, O[j + 3]))
    ops.append(("madc.lo.cc", O[6], a[7], bi, 0))
    ops.append(("madc.hi", O[7], a[7], bi, 0))
    ops.append(("mad.lo.cc", E[0], a[0], bi, E[0]))
    ops.append(("madc.hi.cc", E[1], a[0], bi, E[1]))
    for j in range(2, 8, 2):
        ops.append(("madc.lo.cc", E[j], a[j], bi, E[j]))
        ops.append(("madc.hi.cc", E[j + 1], a[j], bi, E[j + 1]))
    ops.append(("addc", O[7], O[7], 0))
    return ops + reduction_ops(E, O, mod)


def merge_ops(E, O):
    """after the last row (called with (O, E)): result_j = E_j + O_{j+1}"""
    ops = [("add.cc", E[0], E[0], O[1])]
    for i in range(1, 7):
        ops.append(("addc.cc", E[i], E[i], O[i + 1]))
    ops.append(("addc", E[7], E[7], 0))
    return ops



def final_sub_ops(r, mod):
    """r in [0, 2 MOD) -> [0, MOD), branch-free: d = r - MOD; keep d unless the subtraction borrowed."""
    M = limbs32(mod)
    d = ["d%d" % i for i in range(8)]
    ops = [("sub.cc", d[0], r[0], M[0])]
    for i in range(1, 8):
        ops.append(("subc.cc", d[i], r[i], M[i]))
    ops.append(("subc", "bw", 0, 0))
    ops.append(("setp.eq", "pr", "bw", 0))
    for i in range(8):
        ops.append(("selp", r[i], d[i], r[i], "pr"))
    return ops


def add_ops(r, a, b, mod):
    ops = [("add.cc", r[0], a[0], b[0])]
    for i in range(1, 7):
        ops.append(("addc.cc", r[i], a[i], b[i]))
    ops.append(("addc", r[7], a[7], b[7]))
    return ops + final_sub_ops(r, mod)


def sub_ops(r, a, b, mod):
    M = limbs32(mod)
    ops = [("sub.cc", r[0], a[0], b[0])]
    for i in range(1, 8):
        ops.append(("subc.cc", r[i], a[i], b[i]))
    ops.append(("subc", "bw", 0, 0))
    for i in range(8):
        if M[i]:
            ops.append(("and", "d%d" % i, "bw", M[i]))
    first = True
    for i in range(8):
        src = ("d%d" % i) if M[i] else 0
        name = "add.cc" if first else ("addc.cc" if i < 7 else "addc")
        ops.append((name, r[i], r[i], src))
        first = False
    return ops


# ---- executing an op list on Python integers ---------------------------------------------------------
def execute(ops, regs):
    cc = 0

    def val(x):
        if isinstance(x, int):
            return x
        return regs[x]

    for op in ops:
        name, d = op[0], op[1]
        s = [val(x) for x in op[2:]]
        if name == "sub":
            r = (s[0] - s[1]) & MASK
            regs[d] = r
            continue
        if name == "and":
            regs[d] = s[0] & s[1]
            continue
        if name == "setp.eq":
            regs[d] = 1 if s[0] == s[1] else 0
            continue
        if name == "selp":
            regs[d] = s[0] if s[2] else s[1]
            continue
        if name in ("sub.cc", "subc.cc", "subc"):
            t = s[0] - s[1] - (cc if name != "sub.cc" else 0)
            regs[d] = t & MASK
            if name.endswith(".cc"):
                cc = 1 if t < 0 else 0
            continue
        if name == "mul.lo":
            regs[d] = (s[0] * s[1]) & MASK
            continue
        if name == "mul.hi":
            regs[d] = (s[0] * s[1]) >> 32
            continue
        base = name.split(".")
        carry_in = cc if base[0] in ("addc", "madc") else 0
        if base[0] in ("add", "addc"):
            t = s[0] + s[1] + carry_in
        else:
            prod = s[0] * s[1]
            part = (prod & MASK) if base[1] == "lo" else (prod >> 32)
            t = part + s[2] + carry_in
        regs[d] = t & MASK
        if name.endswith(".cc"):
            cc = t >> 32
        elif not regs.get("_allow_wrap"):
            assert t >> 32 == 0, ("dropped carry", op)
    return regs


def mont_mul_ops(mod):
    E = ["E%d" % i for i in range(8)]
    O = ["O%d" % i for i in range(8)]
    a = ["a%d" % i for i in range(8)]
    seq = []
    for i in range(8):
        if i == 0:
            seq += first_row_ops(E, O, a, "b0", mod)
        elif i % 2 == 1:
            seq += row_ops(O, E, a, "b%d" % i, mod)
        else:
            seq += row_ops(E, O, a, "b%d" % i, mod)
    seq += merge_ops(E, O)
    return seq


def check(mod, trials=20000):
    ops = mont_mul_ops(mod)
    rinv = pow(1 << 256, -1, mod)
    rnd = random.Random(1234)
    specials = [0, 1, 2, mod - 1, mod - 2, (1 << 254), (1 << 254) - 1, MASK, (1 << 128) - 1, mod >> 1]
    cases = [(x, y) for x in specials for y in specials]
    cases += [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(trials)]
    # sparse / saturated limb patterns
    for _ in range(2000):
        x = sum(rnd.choice([0, MASK, rnd.getrandbits(32)]) << (32 * i) for i in range(8)) % mod
        y = sum(rnd.choice([0, MASK, rnd.getrandbits(32)]) << (32 * i) for i in range(8)) % mod
        cases.append((x, y))
    for x, y in cases:
        regs = {}
        for i, v in enumerate(limbs32(x)):
            regs["a%d" % i] = v
        for i, v in enumerate(limbs32(y)):
            regs["b%d" % i] = v
        execute(ops, regs)
        got = sum(regs["E%d" % i] << (32 * i) for i in range(8))
        assert got < 2 * mod, "unreduced result out of range"
        if got >= mod:
            got -= mod
        assert got == x * y * rinv % mod, (hex(x), hex(y))
    # add / sub
    r = ["r%d" % i for i in range(8)]
    a = ["a%d" % i for i in range(8)]
    b = ["b%d" % i for i in range(8)]
    aops, sops = add_ops(r, a, b, mod), sub_ops(r, a, b, mod)
    for x, y in cases:
        for ops, want, wrap in ((aops, (x + y) % mod, False), (sops, (x - y) % mod, True)):
            regs = {"_allow_wrap": wrap}
            for i, v in enumerate(limbs32(x)):
                regs["a%d" % i] = v
            for i, v in enumerate(limbs32(y)):
                regs["b%d" % i] = v
            execute(ops, regs)
            got = sum(regs["r%d" % i] << (32 * i) for i in range(8))
            assert got == want, ("add/sub", hex(x), hex(y))
    return len(cases)


# ---- rendering -----------------------------------------------------------------------------------------
PTX_NAME = {"mad.lo.cc": "mad.lo.cc.u32", "madc.lo.cc": "madc.lo.cc.u32", "madc.hi.cc": "madc.hi.cc.u32",
            "madc.hi": "madc.hi.u32", "add.cc": "add.cc.u32", "addc.cc": "addc.cc.u32", "addc": "addc.u32",
            "mul.lo": "mul.lo.u32", "mul.hi": "mul.hi.u32", "sub": "sub.u32",
            "sub.cc": "sub.cc.u32", "subc.cc": "subc.cc.u32", "subc": "subc.u32", "and": "and.b32",
            "setp.eq": "setp.eq.u32", "selp": "selp.u32"}
LOCAL = ("mi", "mj", "bw", "pr") + tuple("d%d" % i for i in range(8))


def render_asm(ops, operand_index, indent="      "):
    """one asm statement; operand_index maps names to %n; `mi`/`mj` are block-local registers."""
    lines = []
    for op in ops:
        args = []
        for x in op[1:]:
            if isinstance(x, int):
                args.append("0x%x" % x if x > 9 else str(x))
            elif x in LOCAL:
                args.append(x)
            else:
                args.append("%%%d" % operand_index[x])
        lines.append('%s"%s %s;\\n\\t"' % (indent, PTX_NAME[op[0]], ", ".join(args)))
    return "\n".join(lines)


def render_field(tag, mod):
    E = ["E%d" % i for i in range(8)]
    O = ["O%d" % i for i in range(8)]
    a = ["a%d" % i for i in range(8)]
    names = E + O + a + ["bi"]
    idx = {n: i for i, n in enumerate(names)}
    out = []

    def fn(name, ops, first):
        out.append("__device__ __forceinline__ void %s_%s(uint32_t (&E)[8], uint32_t (&O)[8], const uint32_t (&a)[8], uint32_t bi) {" % (name, tag))
        out.append("  asm(\"{\\n\\t.reg .u32 mi, mj;\\n\\t\"")
        out.append(render_asm(ops, idx))
        out.append("      \"}\"")
        c = "=&r" if first else "+r"
        outs = ", ".join('"%s"(E[%d])' % (c, i) for i in range(8)) + ",\n        " + \
            ", ".join('"%s"(O[%d])' % (c, i) for i in range(8))
        ins = ", ".join('"r"(a[%d])' % i for i in range(8)) + ', "r"(bi)'
        out.append("      : %s\n      : %s);" % (outs, ins))
        out.append("}")

    fn("mont_first_row", first_row_ops(E, O, a, "bi", mod), True)
    fn("mont_row", row_ops(E, O, a, "bi", mod), False)

    r = ["r%d" % i for i in range(8)]
    b = ["b%d" % i for i in range(8)]
    decl = "{\\n\\t.reg .u32 bw, d0, d1, d2, d3, d4, d5, d6, d7;\\n\\t.reg .pred pr;\\n\\t"
    # final_sub: r in/out
    idx1 = {n: i for i, n in enumerate(r)}
    out.append("__device__ __forceinline__ void final_sub_%s(uint32_t (&r)[8]) {" % tag)
    out.append("  asm(\"%s\"" % decl)
    out.append(render_asm(final_sub_ops(r, mod), idx1))
    out.append("      \"}\"")
    out.append("      : %s);" % ", ".join('"+r"(r[%d])' % i for i in range(8)))
    out.append("}")
    idx2 = {n: i for i, n in enumerate(r + a + b)}
    for name, ops in (("add", add_ops(r, a, b, mod)), ("sub", sub_ops(r, a, b, mod))):
        out.append("__device__ __forceinline__ void %s_%s(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {" % (name, tag))
        out.append("  asm(\"%s\"" % decl)
        out.append(render_asm(ops, idx2))
        out.append("      \"}\"")
        out.append("      : %s" % ", ".join('"=&r"(r[%d])' % i for i in range(8)))
        out.append("      : %s,\n        %s);" % (", ".join('"r"(a[%d])' % i for i in range(8)),
                                                  ", ".join('"r"(b[%d])' % i for i in range(8))))
        out.append("}")
    return "\n".join(out)


HEADER = '''// GENERATED by tools/gen_mont_ptx.py — do not edit.
// 256-bit Montgomery multiplication rows for the Pasta fields as PTX carry chains (device only).
// Replaces the limb arithmetic of pasta_curves 0.5.1 `Fp::mul` / `Fq::mul` (Cargo.lock:1334-1347 of
// the reference) on the GPU; the instruction sequence was executed on Python integers against
// a*b*R^-1 mod p by the generator before being rendered.
#pragma once
#include <cstdint>
namespace zkodst {
namespace montptx {
'''

FOOTER = '''
// result limbs (may be >= MOD, < 2 MOD): E_j + O_{j+1} after eight rows
__device__ __forceinline__ void mont_merge(uint32_t (&E)[8], const uint32_t (&O)[8]) {
  asm("add.cc.u32 %0, %0, %8;\\n\\t"
      "addc.cc.u32 %1, %1, %9;\\n\\t"
      "addc.cc.u32 %2, %2, %10;\\n\\t"
      "addc.cc.u32 %3, %3, %11;\\n\\t"
      "addc.cc.u32 %4, %4, %12;\\n\\t"
      "addc.cc.u32 %5, %5, %13;\\n\\t"
      "addc.cc.u32 %6, %6, %14;\\n\\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7])
      : "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}
}  // namespace montptx
}  // namespace zkodst
'''


def main():
    for name, mod in (("Fp", P), ("Fq", Q)):
        ncase = check(mod)
        print("%s: %d products agree with a*b*R^-1 mod p" % (name, ncase))
    if "--check-only" in sys.argv:
        return
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "..", "zk-odst_b200", "csrc", "mont_ptx.cuh")
    body = HEADER + render_field("fp", P) + "\n" + render_field("fq", Q) + FOOTER
    with open(path, "w") as f:
        f.write(body)
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
