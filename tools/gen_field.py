#!/usr/bin/env python3
"""Generator for the sm_100a Montgomery field kernels (hekaton_system_b200/csrc/field_gen.cuh).

Why generated: on sm_100a ptxas lowers `mad.lo.cc/madc.hi.cc` to IMAD(.HI) + IADD3.X
(4 SASS per 32x32 MAC) but fuses `mul.wide.u32` + `add.cc.u64/addc.cc.u64` into ONE
`IMAD.WIDE.U32.X Rd, Pc, Ra, Rb, Rd, Pc` (multiply, 64-bit accumulate, carry in, carry
out).  So field elements are kept as 64-bit lanes and the Montgomery product is a set
of carry chains over two families of 64-bit lanes (even / odd 32-bit column alignment).
Each operation is emitted as a single fully unrolled PTX block with the modulus as
immediates; the same instruction list is executed by the simulator below
(tests/test_field_gen.py) against Python big integers, so the arithmetic is verified
without a GPU.

Algorithm (mont_mul, N 32-bit limbs, N even, p < 2^(32N-1)):
  lanes L[c] = 64-bit accumulator covering absolute 32-bit columns (c, c+1).
  row i (0..N-1):
    X_i : L[i+j]  += a_j*b_i   j even (first product also adds x_i, the high half of
                                L[i-1] left over from the previous row); carry -> L[i+N]
    Y_i : L[i+j]  += a_j*b_i   j odd ; no carry out (total < 2^(32(i+N+1)))
    m_i  = lo32(L[i]) * (-p^-1 mod 2^32)
    X'_i: L[i+j]  += m_i*p_j   j even; carry -> L[i+N]
    Y'_i: L[i+j]  += m_i*p_j   j odd
    now lo32(L[i]) == 0 ;  x_{i+1} = hi32(L[i])
  result = (columns N..2N-1) = even family + odd family, then one conditional subtract.
Cost: 2N^2 + N wide multiplies (+ N low multiplies for m_i).

Variants generated from the same lane scheme (each with its own carry-bound argument in the docstring and the same
instruction-level verification):
  gen_sqr       dedicated squaring, N(N+1)/2 + N^2 + N                      (fields with two spare top bits)
  gen_mul2      fused dual product (a*b + c*d) R^-1 with one reduction, 3N^2 + N   (3p < 2^(32N-1): BLS12-381 Fq)
  gen_negp      p - a, unreduced: the negated operand of gen_mul2
  gen_mul_kara  Karatsuba product + separate reduction, 3(N/2)^2 + N^2 + N   (opt-in -DHK_MUL_KARA: measured slower,
                see DESIGN.md 3.1)
tools/kara_experiment.py + tools/sass_mix.sh / sass_split.py give the per-pipe SASS mix of a variant without a GPU.
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))

M32 = (1 << 32) - 1
M64 = (1 << 64) - 1


class Prog:
    """Tiny PTX-subset program builder + simulator."""

    def __init__(self):
        self.ins = []       # (op, dst, srcs)
        self.regs64 = []
        self.regs32 = []
        self.preds = []
        self._n = 0

    def r64(self, hint="q"):
        self._n += 1
        name = f"{hint}{self._n}"
        self.regs64.append(name)
        return name

    def r32(self, hint="w"):
        self._n += 1
        name = f"{hint}{self._n}"
        self.regs32.append(name)
        return name

    def pred(self):
        self._n += 1
        name = f"p{self._n}"
        self.preds.append(name)
        return name

    def emit(self, op, dst, *srcs):
        self.ins.append((op, dst, srcs))

    # ---------------- simulation ----------------
    def run(self, env):
        """env: dict name -> int for inputs.  Returns env after execution."""
        R = dict(env)
        cf = 0

        def val(s):
            if isinstance(s, int):
                return s
            return R[s]

        for op, dst, srcs in self.ins:
            if op == "mul.wide.u32":
                R[dst] = (val(srcs[0]) & M32) * (val(srcs[1]) & M32)
            elif op == "mad.wide.u32":
                v = (val(srcs[0]) & M32) * (val(srcs[1]) & M32) + val(srcs[2])
                assert v <= M64, "mad.wide overflow (would silently wrap)"
                R[dst] = v
            elif op == "mul.lo.u32":
                R[dst] = (val(srcs[0]) * val(srcs[1])) & M32
            elif op in ("add.cc.u64", "addc.cc.u64", "addc.u64", "add.u64"):
                v = val(srcs[0]) + val(srcs[1]) + (cf if op.startswith("addc") else 0)
                R[dst] = v & M64
                if ".cc" in op:
                    cf = v >> 64
                elif op == "addc.u64":
                    pass
            elif op in ("sub.cc.u64", "subc.cc.u64", "subc.u64"):
                v = val(srcs[0]) - val(srcs[1]) - (cf if op.startswith("subc") else 0)
                R[dst] = v & M64
                if ".cc" in op:
                    cf = 1 if v < 0 else 0
            elif op == "subc.u32":
                v = val(srcs[0]) - val(srcs[1]) - cf
                R[dst] = v & M32
            elif op in ("add.cc.u32", "addc.cc.u32", "addc.u32"):
                v = val(srcs[0]) + val(srcs[1]) + (cf if op.startswith("addc") else 0)
                R[dst] = v & M32
                if ".cc" in op:
                    cf = v >> 32
            elif op in ("sub.cc.u32", "subc.cc.u32"):
                v = val(srcs[0]) - val(srcs[1]) - (cf if op.startswith("subc") else 0)
                R[dst] = v & M32
                cf = 1 if v < 0 else 0
            elif op == "xor.b32":
                R[dst] = (val(srcs[0]) ^ val(srcs[1])) & M32
            elif op == "assert_le1":
                assert val(srcs[0]) <= 1, "top limb of a Karatsuba middle term out of range"
            elif op == "unpack":          # mov.b64 {lo,hi}, src
                lo, hi = dst
                v = val(srcs[0])
                R[lo] = v & M32
                R[hi] = v >> 32
            elif op == "pack":            # mov.b64 dst, {lo,hi}
                R[dst] = (val(srcs[0]) & M32) | ((val(srcs[1]) & M32) << 32)
            elif op == "mov.u64":
                R[dst] = val(srcs[0])
            elif op == "setp.ne.u32":
                R[dst] = 1 if val(srcs[0]) != val(srcs[1]) else 0
            elif op == "selp.b64":
                R[dst] = val(srcs[0]) if R[srcs[2]] else val(srcs[1])
            elif op == "and.b64":
                R[dst] = val(srcs[0]) & val(srcs[1])
            elif op == "add.u32":
                R[dst] = (val(srcs[0]) + val(srcs[1])) & M32
            elif op == "shl.b32":
                R[dst] = (val(srcs[0]) << val(srcs[1])) & M32
            elif op == "shf.l.wrap.b32":      # upper 32 bits of (hi:lo) << n
                lo_, hi_, n_ = val(srcs[0]), val(srcs[1]), val(srcs[2]) & 31
                R[dst] = (((hi_ << 32) | lo_) << n_ >> 32) & M32
            elif op == "assert_nc":
                assert cf == 0, "dropped a non-zero carry"
            elif op == "assert_lo0":
                assert val(srcs[0]) & M32 == 0
            elif op == "assert_hi0":
                assert val(srcs[0]) == 0, "non-zero bits above the result"
            else:
                raise ValueError(op)
        return R

    # ---------------- PTX text ----------------
    def ptx(self, operand_map):
        """operand_map: name -> '%k' for asm operands.  Returns list of PTX lines."""
        def o(s):
            if isinstance(s, int):
                return hex(s)
            return operand_map.get(s, s)

        lines = []
        r64 = [r for r in self.regs64 if r not in operand_map]
        r32 = [r for r in self.regs32 if r not in operand_map]
        if r64:
            lines.append(".reg .u64 " + ", ".join(r64) + ";")
        if r32:
            lines.append(".reg .u32 " + ", ".join(r32) + ";")
        if self.preds:
            lines.append(".reg .pred " + ", ".join(self.preds) + ";")
        for op, dst, srcs in self.ins:
            if op.startswith("assert_"):
                continue
            if op == "unpack":
                lines.append(f"mov.b64 {{{o(dst[0])}, {o(dst[1])}}}, {o(srcs[0])};")
            elif op == "pack":
                lines.append(f"mov.b64 {o(dst)}, {{{o(srcs[0])}, {o(srcs[1])}}};")
            else:
                lines.append(f"{op} {o(dst)}, " + ", ".join(o(s) for s in srcs) + ";")
        return lines


def limbs32(x, n):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def lanes64(x, n2):
    return [(x >> (64 * i)) & M64 for i in range(n2)]


class FieldGen:
    def __init__(self, name, p):
        self.name = name
        self.p = p
        self.N = ((p.bit_length() + 63) // 64) * 2        # 32-bit limbs, even
        self.N2 = self.N // 2
        assert p.bit_length() <= 32 * self.N - 1, "need a spare top bit"
        self.R = 1 << (32 * self.N)
        self.m32 = (-pow(p, -1, 1 << 32)) % (1 << 32)
        self.pl = limbs32(p, self.N)
        self.p64 = lanes64(p, self.N2)
        # dedicated squaring needs 2a + p < 2^(32N) for every a < p (see gen_sqr)
        self.has_sqr = 3 * p < (1 << (32 * self.N))

    # -- shared tail: r (N2 lanes, < 2p) -> r mod p into output names
    def _final_sub(self, P, lanes, outs):
        t = [P.r64("t") for _ in range(self.N2)]
        for k in range(self.N2):
            P.emit("sub.cc.u64" if k == 0 else "subc.cc.u64", t[k], lanes[k], self.p64[k])
        bw = P.r32("bw")
        P.emit("subc.u32", bw, 0, 0)           # 0xffffffff if r < p (borrow), else 0
        pr = P.pred()
        P.emit("setp.ne.u32", pr, bw, 0)
        for k in range(self.N2):
            P.emit("selp.b64", outs[k], lanes[k], t[k], pr)

    def gen_mul(self):
        N, N2 = self.N, self.N2
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        bin_ = [f"b{k}" for k in range(N2)]
        a = []
        b = []
        for k in range(N2):
            lo, hi = P.r32("al"), P.r32("ah")
            P.emit("unpack", (lo, hi), ain[k])
            a += [lo, hi]
        for k in range(N2):
            lo, hi = P.r32("bl"), P.r32("bh")
            P.emit("unpack", (lo, hi), bin_[k])
            b += [lo, hi]
        L = {}
        x = None            # 32-bit register: hi32 of the retired lane

        def chain(cols, prods, carry_lane, first_addend=None):
            """prods: list of (u32 reg/imm, u32 reg/imm) per lane col."""
            carrying = False
            for idx, (c, (u, v)) in enumerate(zip(cols, prods)):
                t = P.r64("t")
                if idx == 0 and first_addend is not None:
                    x64 = P.r64("x")
                    P.emit("pack", x64, first_addend, 0)
                    P.emit("mad.wide.u32", t, u, v, x64)
                else:
                    P.emit("mul.wide.u32", t, u, v)
                if c in L:
                    P.emit("addc.cc.u64" if carrying else "add.cc.u64", L[c], L[c], t)
                    carrying = True
                else:
                    if carrying:
                        L[c] = P.r64("L")
                        P.emit("addc.cc.u64", L[c], t, 0)
                    else:
                        L[c] = t
            if carry_lane is None:
                if carrying:
                    P.emit("assert_nc", None)
            elif carrying:
                if carry_lane in L:
                    P.emit("addc.u64", L[carry_lane], L[carry_lane], 0)
                else:
                    L[carry_lane] = P.r64("L")
                    P.emit("addc.u64", L[carry_lane], 0, 0)

        for i in range(N):
            ev = list(range(0, N, 2))
            od = list(range(1, N, 2))
            chain([i + j for j in ev], [(a[j], b[i]) for j in ev], i + N, first_addend=x)
            chain([i + j for j in od], [(a[j], b[i]) for j in od], None)
            lo, hi = P.r32("lo"), P.r32("hi")
            P.emit("unpack", (lo, hi), L[i])
            m = P.r32("m")
            P.emit("mul.lo.u32", m, lo, self.m32)
            chain([i + j for j in ev], [(m, self.pl[j]) for j in ev], i + N)
            chain([i + j for j in od], [(m, self.pl[j]) for j in od], None)
            P.emit("assert_lo0", None, L[i])
            lo2, hi2 = P.r32("z"), P.r32("x")
            P.emit("unpack", (lo2, hi2), L[i])
            x = hi2
            del L[i]
        # merge: even family lanes L[N], L[N+2], ... ; odd family: x at col N, L[N+1].. L[2N-1]
        odd32 = {N: x}
        for c in range(N + 1, 2 * N, 2):
            if c in L:
                lo, hi = P.r32("ol"), P.r32("oh")
                P.emit("unpack", (lo, hi), L[c])
                odd32[c] = lo
                if c + 1 < 2 * N:
                    odd32[c + 1] = hi
        res = []
        for k in range(N2):
            c = N + 2 * k
            o64 = P.r64("o")
            P.emit("pack", o64, odd32.get(c, 0), odd32.get(c + 1, 0))
            lane = L.get(c)
            if lane is None:
                lane = P.r64("L")
                P.emit("mov.u64", lane, 0)
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", lane, lane, o64)
            res.append(lane)
        P.emit("assert_nc", None)
        self._final_sub(P, res, outs)
        return P, outs, ain + bin_

    def gen_mul_kara(self, variant=""):
        """Montgomery product with one level of Karatsuba on the integer product and a separate reduction:
            a = a0 + a1*B, b = b0 + b1*B, B = 2^(16N)          (H = N/2 limbs per half)
            z0 = a0*b0, z2 = a1*b1, zm = |a0 - a1| * |b1 - b0|   three H x H products, H^2 wide multiplies each
            z1 = z0 + z2 +- zm  (= a0*b1 + a1*b0)
            T  = z0 + z1*B + z2*B^2
        then N reduction rows (m_i = column_i * (-p^-1), T += m_i * p << 32i) on the lane scheme of gen_mul with the
        low half of T preloaded into the even lanes; the high half of T is added to the surviving lanes at the end.
        3(N/2)^2 + N^2 + N wide multiplies instead of 2N^2 + N (BLS12-381 Fq: 264 instead of 300); the glue
        (differences, conditional negations, the three merges) is ~130 32-bit add/sub/xor on the ALU pipe, which the
        multiplier-bound MSM kernels leave idle (IMAD.WIDE issues once per 4 cycles per scheduler).
        No carry is dropped: all lanes are non-negative and their weighted sum V stays below 2^(32(i+N+1)+30)
        after row i (T_lo < 2^(32N), sum m_k p 2^(32k) < 2^(32(i+1)) p), so the last lane of the odd chain
        (i+N-1) and the carry lane (i+N) cannot overflow; the simulator asserts every dropped carry is zero."""
        N, N2 = self.N, self.N2
        assert N % 4 == 0
        H = N // 2
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        bin_ = [f"b{k}" for k in range(N2)]
        a, b = [], []
        for k in range(N2):
            lo, hi = P.r32("al"), P.r32("ah")
            P.emit("unpack", (lo, hi), ain[k])
            a += [lo, hi]
        for k in range(N2):
            lo, hi = P.r32("bl"), P.r32("bh")
            P.emit("unpack", (lo, hi), bin_[k])
            b += [lo, hi]

        def chain(L, cols, prods, carry_lane, first_addend=None):
            carrying = False
            for idx, (c, (u, v)) in enumerate(zip(cols, prods)):
                t = P.r64("t")
                if idx == 0 and first_addend is not None:
                    x64 = P.r64("x")
                    P.emit("pack", x64, first_addend, 0)
                    P.emit("mad.wide.u32", t, u, v, x64)
                else:
                    P.emit("mul.wide.u32", t, u, v)
                if c in L:
                    P.emit("addc.cc.u64" if carrying else "add.cc.u64", L[c], L[c], t)
                    carrying = True
                else:
                    if carrying:
                        L[c] = P.r64("L")
                        P.emit("addc.cc.u64", L[c], t, 0)
                    else:
                        L[c] = t
            if carry_lane is None:
                if carrying:
                    P.emit("assert_nc", None)
            elif carrying:
                assert carry_lane not in L
                L[carry_lane] = P.r64("L")
                P.emit("addc.u64", L[carry_lane], 0, 0)

        def half_product(u, v):
            """u, v: H 32-bit registers each -> 2H 32-bit limbs of u*v."""
            L = {}
            ev = list(range(0, H, 2))
            od = list(range(1, H, 2))
            for i in range(H):
                chain(L, [i + j for j in ev], [(u[j], v[i]) for j in ev], i + H)
                chain(L, [i + j for j in od], [(u[j], v[i]) for j in od], None)
            # merge the two lane families: limb c = even-family half + odd-family half + carry
            e32, o32 = {}, {}
            for c in sorted(L):
                lo, hi = P.r32("pl"), P.r32("ph")
                P.emit("unpack", (lo, hi), L[c])
                fam = e32 if c % 2 == 0 else o32
                fam[c] = lo
                if c + 1 < 2 * H:
                    fam[c + 1] = hi
                else:
                    P.emit("assert_hi0", None, hi)
            z = []
            started = False
            for c in range(2 * H):
                if c in e32 and c in o32:
                    w = P.r32("z")
                    P.emit("addc.cc.u32" if started else "add.cc.u32", w, e32[c], o32[c])
                    started = True
                elif started:
                    w = P.r32("z")
                    P.emit("addc.cc.u32", w, e32.get(c, o32.get(c, 0)), 0)
                else:
                    w = e32.get(c, o32.get(c))
                z.append(w)
            P.emit("assert_nc", None)
            return z

        def abs_diff(u, v):
            """|u - v| over H limbs and the sign mask (0xffffffff when u < v)."""
            d = [P.r32("d") for _ in range(H)]
            for k in range(H):
                P.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", d[k], u[k], v[k])
            mask = P.r32("sm")
            P.emit("subc.u32", mask, 0, 0)
            out = []
            for k in range(H):
                t = P.r32("dx")
                P.emit("xor.b32", t, d[k], mask)
                out.append(t)
            res = [P.r32("ad") for _ in range(H)]
            for k in range(H):
                P.emit("sub.cc.u32" if k == 0 else "subc.cc.u32", res[k], out[k], mask)
            return res, mask

        a0, a1, b0, b1 = a[:H], a[H:], b[:H], b[H:]
        if "mfirst" in variant:
            da, sa = abs_diff(a0, a1)
            db, sb = abs_diff(b1, b0)
            zm = half_product(da, db)
            z0 = half_product(a0, b0)
            z2 = half_product(a1, b1)
        else:
            z0 = half_product(a0, b0)
            z2 = half_product(a1, b1)
            da, sa = abs_diff(a0, a1)
            db, sb = abs_diff(b1, b0)
            zm = half_product(da, db)
        sm = P.r32("sg")
        P.emit("xor.b32", sm, sa, sb)             # all ones: subtract zm
        # S = z0 + z2 (2H + 1 limbs)
        S = [P.r32("s") for _ in range(2 * H + 1)]
        for k in range(2 * H):
            P.emit("add.cc.u32" if k == 0 else "addc.cc.u32", S[k], z0[k], z2[k])
        P.emit("addc.u32", S[2 * H], 0, 0)
        # z1 = S + (zm ^ sm) + (sm & 1), sign-extended: two's complement subtraction when sm is all ones
        zx = []
        for k in range(2 * H):
            t = P.r32("zx")
            P.emit("xor.b32", t, zm[k], sm)
            zx.append(t)
        z1 = [P.r32("m") for _ in range(2 * H + 1)]
        cg = P.r32("cg")
        P.emit("add.cc.u32", cg, sm, sm)          # carry flag = (sm != 0)
        for k in range(2 * H):
            P.emit("addc.cc.u32", z1[k], S[k], zx[k])
        P.emit("addc.u32", z1[2 * H], S[2 * H], sm)
        P.emit("assert_le1", None, z1[2 * H])
        # T = z0 + z1 << 32H + z2 << 64H   (4H = 2N limbs)
        T = list(z0[:H])
        first = True
        for k in range(H, 4 * H):
            base = z0[k] if k < 2 * H else z2[k - 2 * H]
            add = z1[k - H] if k - H <= 2 * H else 0
            w = P.r32("T")
            P.emit("add.cc.u32" if first else "addc.cc.u32", w, base, add)
            first = False
            T.append(w)
        cz = None
        if "dep" in variant:
            # the carry out of T is provably zero; feeding it to the first reduction digit keeps the scheduler from
            # sinking the upper halves of the glue chains into the reduction (their carry predicates would stay live
            # across it: seven predicate registers are not enough and ptxas starts spilling them through P2R)
            cz = P.r32("cz")
            P.emit("addc.u32", cz, 0, 0)
            P.emit("assert_hi0", None, cz)
        else:
            P.emit("assert_nc", None)

        # ---- reduction: low half of T preloaded in the even lanes ----
        L = {}
        for c in range(0, N, 2):
            L[c] = P.r64("L")
            P.emit("pack", L[c], T[c], T[c + 1])
        x = None
        ev = list(range(0, N, 2))
        od = list(range(1, N, 2))
        for i in range(N):
            lo, hi = P.r32("lo"), P.r32("hi")
            P.emit("unpack", (lo, hi), L[i])
            m = P.r32("m")
            if x is None:
                if cz is not None:
                    lz = P.r32("lz")
                    P.emit("add.u32", lz, lo, cz)
                    lo = lz
                P.emit("mul.lo.u32", m, lo, self.m32)
            else:
                sx = P.r32("sx")
                P.emit("add.u32", sx, lo, x)
                P.emit("mul.lo.u32", m, sx, self.m32)
            chain(L, [i + j for j in ev], [(m, self.pl[j]) for j in ev], i + N, first_addend=x)
            chain(L, [i + j for j in od], [(m, self.pl[j]) for j in od], None)
            P.emit("assert_lo0", None, L[i])
            lo2, hi2 = P.r32("z"), P.r32("x")
            P.emit("unpack", (lo2, hi2), L[i])
            x = hi2
            del L[i]
        odd32 = {N: x}
        for c in range(N + 1, 2 * N, 2):
            if c in L:
                lo, hi = P.r32("ol"), P.r32("oh")
                P.emit("unpack", (lo, hi), L[c])
                odd32[c] = lo
                if c + 1 < 2 * N:
                    odd32[c + 1] = hi
                else:
                    P.emit("assert_hi0", None, hi)
        res = []
        for k in range(N2):
            c = N + 2 * k
            o64 = P.r64("o")
            P.emit("pack", o64, odd32.get(c, 0), odd32.get(c + 1, 0))
            lane = L.get(c)
            if lane is None:
                lane = P.r64("L")
                P.emit("mov.u64", lane, 0)
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", lane, lane, o64)
            res.append(lane)
        P.emit("assert_nc", None)
        for k in range(N2):
            th = P.r64("th")
            P.emit("pack", th, T[N + 2 * k], T[N + 2 * k + 1])
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", res[k], res[k], th)
        P.emit("assert_nc", None)
        self._final_sub(P, res, outs)
        return P, outs, ain + bin_

    def gen_mul2(self):
        """Fused dual product (a*b + c*d) * R^-1 mod p with ONE reduction: 3N^2 + N wide multiplies instead of
        2(2N^2 + N) for two Montgomery products and an addition (BLS12-381 Fq: 444 instead of 600).  Row i adds
        a*b_i and c*d_i to the lanes, then the usual reduction row.  Needs 3p < 2^(32N-1) so that the running
        total a*(b_0..b_i) + c*(d_0..d_i) + sum m_k p 2^(32k) < 3p 2^(32(i+1)) keeps the top lane of a row
        (i+N-1) below 2^63 and the carry lane (i+N) below 2^31; the result is < (2p^2 + Rp)/R < 2p.  Inputs may be
        any values <= p (so a caller can pass p - x for a difference of products)."""
        N, N2 = self.N, self.N2
        assert 3 * self.p < (1 << (32 * N - 1))
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        names = ["a", "b", "c", "d"]
        ins = [f"{nm}{k}" for nm in names for k in range(N2)]
        limbs = {}
        for nm in names:
            v = []
            for k in range(N2):
                lo, hi = P.r32(nm + "l"), P.r32(nm + "h")
                P.emit("unpack", (lo, hi), f"{nm}{k}")
                v += [lo, hi]
            limbs[nm] = v
        a, b, c, d = (limbs[nm] for nm in names)
        L = {}
        x = None

        def chain(cols, prods, carry_lane, first_addend=None):
            carrying = False
            for idx, (col, (u, v)) in enumerate(zip(cols, prods)):
                t = P.r64("t")
                if idx == 0 and first_addend is not None:
                    x64 = P.r64("x")
                    P.emit("pack", x64, first_addend, 0)
                    P.emit("mad.wide.u32", t, u, v, x64)
                else:
                    P.emit("mul.wide.u32", t, u, v)
                if col in L:
                    P.emit("addc.cc.u64" if carrying else "add.cc.u64", L[col], L[col], t)
                    carrying = True
                else:
                    if carrying:
                        L[col] = P.r64("L")
                        P.emit("addc.cc.u64", L[col], t, 0)
                    else:
                        L[col] = t
            if carry_lane is None:
                if carrying:
                    P.emit("assert_nc", None)
            elif carrying:
                if carry_lane in L:
                    P.emit("addc.cc.u64", L[carry_lane], L[carry_lane], 0)
                    P.emit("assert_nc", None)
                else:
                    L[carry_lane] = P.r64("L")
                    P.emit("addc.u64", L[carry_lane], 0, 0)

        ev = list(range(0, N, 2))
        od = list(range(1, N, 2))
        for i in range(N):
            chain([i + j for j in ev], [(a[j], b[i]) for j in ev], i + N, first_addend=x)
            chain([i + j for j in od], [(a[j], b[i]) for j in od], None)
            chain([i + j for j in ev], [(c[j], d[i]) for j in ev], i + N)
            chain([i + j for j in od], [(c[j], d[i]) for j in od], None)
            lo, hi = P.r32("lo"), P.r32("hi")
            P.emit("unpack", (lo, hi), L[i])
            m = P.r32("m")
            P.emit("mul.lo.u32", m, lo, self.m32)
            chain([i + j for j in ev], [(m, self.pl[j]) for j in ev], i + N)
            chain([i + j for j in od], [(m, self.pl[j]) for j in od], None)
            P.emit("assert_lo0", None, L[i])
            lo2, hi2 = P.r32("z"), P.r32("x")
            P.emit("unpack", (lo2, hi2), L[i])
            x = hi2
            del L[i]
        odd32 = {N: x}
        for col in range(N + 1, 2 * N, 2):
            if col in L:
                lo, hi = P.r32("ol"), P.r32("oh")
                P.emit("unpack", (lo, hi), L[col])
                odd32[col] = lo
                if col + 1 < 2 * N:
                    odd32[col + 1] = hi
                else:
                    P.emit("assert_hi0", None, hi)
        res = []
        for k in range(N2):
            col = N + 2 * k
            o64 = P.r64("o")
            P.emit("pack", o64, odd32.get(col, 0), odd32.get(col + 1, 0))
            lane = L.get(col)
            if lane is None:
                lane = P.r64("L")
                P.emit("mov.u64", lane, 0)
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", lane, lane, o64)
            res.append(lane)
        P.emit("assert_nc", None)
        self._final_sub(P, res, outs)
        return P, outs, ins

    def gen_sqr(self):
        """Dedicated Montgomery squaring: every off-diagonal product a_i*a_k (i < k) is taken once against
        the doubled operand, N(N+1)/2 + N^2 + N wide multiplies instead of 2N^2 + N.
        Row i (interleaved with its reduction row exactly as in gen_mul) adds
            a_i * a_i                      at lane 2i
            a_i * E_k,  k > i              at lane i + k,   E = 2 * (a with limbs <= i cleared):
                E_(i+1) = (a_(i+1) << 1) mod 2^32,   E_k = (a_k << 1 | a_(k-1) >> 31) for k >= i + 2
        (2a < 2^(32N) needs one spare top bit; the carry invariants of gen_mul additionally need
        2a + p < 2^(32N), i.e. TWO spare bits -- checked in __init__ via `self.has_sqr`).
        Column i is complete before reduction row i: a pair (j, k), j < k, j + k = i was added in row j < i."""
        N, N2 = self.N, self.N2
        assert self.has_sqr
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        a = []
        for k in range(N2):
            lo, hi = P.r32("al"), P.r32("ah")
            P.emit("unpack", (lo, hi), ain[k])
            a += [lo, hi]
        D, Dn = {}, {}          # D[k] = (a_k << 1 | a_(k-1) >> 31), Dn[k] = a_k << 1
        for k in range(1, N):
            Dn[k] = P.r32("dn")
            P.emit("shl.b32", Dn[k], a[k], 1)
            D[k] = P.r32("d")
            P.emit("shf.l.wrap.b32", D[k], a[k - 1], a[k], 1)
        L = {}
        x = None

        def chain(cols, prods, carry_lane, first_addend=None):
            carrying = False
            for idx, (c, (u, v)) in enumerate(zip(cols, prods)):
                t = P.r64("t")
                if idx == 0 and first_addend is not None:
                    x64 = P.r64("x")
                    P.emit("pack", x64, first_addend, 0)
                    P.emit("mad.wide.u32", t, u, v, x64)
                else:
                    P.emit("mul.wide.u32", t, u, v)
                if c in L:
                    P.emit("addc.cc.u64" if carrying else "add.cc.u64", L[c], L[c], t)
                    carrying = True
                else:
                    if carrying:
                        L[c] = P.r64("L")
                        P.emit("addc.cc.u64", L[c], t, 0)
                    else:
                        L[c] = t
            if carry_lane is None:
                if carrying:
                    P.emit("assert_nc", None)
            elif carrying:
                if carry_lane in L:
                    P.emit("addc.u64", L[carry_lane], L[carry_lane], 0)
                else:
                    L[carry_lane] = P.r64("L")
                    P.emit("addc.u64", L[carry_lane], 0, 0)

        def carry_lane_for(i, js):
            # the chain's last lane is i + js[-1]; a carry out of it lands two lanes up.  Lanes up to i + N - 1
            # can be live in row i, so only a chain ending at i + N - 2 may carry (into lane i + N); one ending at
            # lane i + N - 1 cannot, because the running total stays below 2^(32 (i + N + 1)) (see docstring).
            return i + N if js[-1] == N - 2 else None

        for i in range(N):
            opnd = {i: a[i]}
            for k in range(i + 1, N):
                opnd[k] = Dn[k] if k == i + 1 else D[k]
            same = [j for j in range(i, N) if (j - i) % 2 == 0]        # lanes 2i, 2i+2, ...: the family of lane i
            other = [j for j in range(i + 1, N) if (j - i) % 2 == 1]
            chain([i + j for j in same], [(a[i], opnd[j]) for j in same], carry_lane_for(i, same))
            if other:
                chain([i + j for j in other], [(a[i], opnd[j]) for j in other], carry_lane_for(i, other))
            # column i = lo32(L[i]) + x  (x: high half of the lane retired by the previous row)
            lo, hi = P.r32("lo"), P.r32("hi")
            P.emit("unpack", (lo, hi), L[i])
            m = P.r32("m")
            if x is None:
                P.emit("mul.lo.u32", m, lo, self.m32)
            else:
                sx = P.r32("sx")
                P.emit("add.u32", sx, lo, x)
                P.emit("mul.lo.u32", m, sx, self.m32)
            ev = list(range(0, N, 2))
            od = list(range(1, N, 2))
            chain([i + j for j in ev], [(m, self.pl[j]) for j in ev], i + N, first_addend=x)
            chain([i + j for j in od], [(m, self.pl[j]) for j in od], None)
            P.emit("assert_lo0", None, L[i])
            lo2, hi2 = P.r32("z"), P.r32("x")
            P.emit("unpack", (lo2, hi2), L[i])
            x = hi2
            del L[i]
        odd32 = {N: x}
        for c in range(N + 1, 2 * N, 2):
            if c in L:
                lo, hi = P.r32("ol"), P.r32("oh")
                P.emit("unpack", (lo, hi), L[c])
                odd32[c] = lo
                if c + 1 < 2 * N:
                    odd32[c + 1] = hi
                else:
                    P.emit("assert_hi0", None, hi)
        res = []
        for k in range(N2):
            c = N + 2 * k
            o64 = P.r64("o")
            P.emit("pack", o64, odd32.get(c, 0), odd32.get(c + 1, 0))
            lane = L.get(c)
            if lane is None:
                lane = P.r64("L")
                P.emit("mov.u64", lane, 0)
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", lane, lane, o64)
            res.append(lane)
        P.emit("assert_nc", None)
        self._final_sub(P, res, outs)
        return P, outs, ain

    def gen_add(self):
        N2 = self.N2
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        bin_ = [f"b{k}" for k in range(N2)]
        s = [P.r64("s") for _ in range(N2)]
        for k in range(N2):
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", s[k], ain[k], bin_[k])
        P.emit("assert_nc", None)
        self._final_sub(P, s, outs)
        return P, outs, ain + bin_

    def gen_sub(self):
        N2 = self.N2
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        bin_ = [f"b{k}" for k in range(N2)]
        d = [P.r64("d") for _ in range(N2)]
        for k in range(N2):
            P.emit("sub.cc.u64" if k == 0 else "subc.cc.u64", d[k], ain[k], bin_[k])
        bw = P.r32("bw")
        P.emit("subc.u32", bw, 0, 0)
        pr = P.pred()
        P.emit("setp.ne.u32", pr, bw, 0)
        # add back p when a < b
        for k in range(N2):
            pm = P.r64("pm")
            P.emit("selp.b64", pm, self.p64[k], 0, pr)
            P.emit("add.cc.u64" if k == 0 else "addc.cc.u64", outs[k], d[k], pm)
        return P, outs, ain + bin_

    def gen_negp(self):
        """p - a without reduction (a <= p): the negated operand of gen_mul2."""
        N2 = self.N2
        P = Prog()
        outs = [f"r{k}" for k in range(N2)]
        ain = [f"a{k}" for k in range(N2)]
        for k in range(N2):
            P.emit("sub.cc.u64" if k == 0 else "subc.cc.u64", outs[k], self.p64[k], ain[k])
        P.emit("assert_nc", None)
        return P, outs, ain

    # ---------- emit C++ ----------
    def cpp_fn(self, fname, P, outs, ins, out_early_clobber=False):
        omap = {}
        k = 0
        for o in outs:
            omap[o] = f"%{k}"
            k += 1
        for i_ in ins:
            omap[i_] = f"%{k}"
            k += 1
        lines = P.ptx(omap)
        body = "\n".join('        "' + ln + '\\n\\t"' for ln in ["{"] + lines + ["}"])
        n2 = self.N2
        cons = "=&l" if out_early_clobber else "=l"
        outc = ", ".join(f'"{cons}"(r.v[{i}])' for i in range(n2))
        nin = len(ins) // n2
        names = ["a", "b", "c", "d"][:nin]
        inc = ", ".join(f'"l"({nm}.v[{i}])' for nm in names for i in range(n2))
        args = ", ".join(f"const {self.name}_t& {nm}" for nm in names)
        # HK_MUL_CALL=1 emits the big multiplies as real functions (one copy of the ~7 KB body per kernel
        # instead of one per call site) -- an instruction-cache experiment, see DESIGN.md
        attr = "__noinline__" if (fname == "mul" and os.environ.get("HK_MUL_CALL") == "1") else "__forceinline__"
        return (f"__device__ {attr} void {self.name}_{fname}({self.name}_t& r, {args}) {{\n"
                f"    asm(\n{body}\n        : {outc}\n        : {inc});\n}}\n")


def field_list():
    # the product's own copy of the curve constants (never the oracle's: a wrong constant there would otherwise be
    # baked into the kernels AND the checker)
    from hekaton_system_b200.curves import BLS12_381, BN254
    return [
        ("bls12_381_fq", BLS12_381.p),
        ("bls12_381_fr", BLS12_381.r),
        ("bn254_fq", BN254.p),
        ("bn254_fr", BN254.r),
    ]


def extra_constants():
    """Fr-only constants (Montgomery form): 2-adic root of unity and F::GENERATOR
    (ark-ff FftField::TWO_ADIC_ROOT_OF_UNITY / GENERATOR)."""
    from hekaton_system_b200.curves import BLS12_381, BN254
    out = {}
    for c in (BLS12_381, BN254):
        R = 1 << 256
        two_adic_root = pow(c.fr_generator, (c.r - 1) >> c.two_adicity, c.r)
        out[c.name + "_fr"] = [
            ("ROOT", two_adic_root * R % c.r, "TWO_ADIC_ROOT_OF_UNITY * R mod r"),
            ("GEN", c.fr_generator * R % c.r, "F::GENERATOR * R mod r"),
        ]
    return out


def c_array64(x, n2):
    return "{" + ", ".join(f"0x{v:016x}ull" for v in lanes64(x, n2)) + "}"


def emit_header(path):
    out = []
    out.append("// GENERATED by tools/gen_field.py -- do not edit.  Regenerate: python tools/gen_field.py\n"
               "// Montgomery arithmetic for sm_100a as fully unrolled PTX; see the generator for the\n"
               "// algorithm and tests/test_field_gen.py for the instruction-level verification.\n"
               "#pragma once\n#include <cstdint>\n\nnamespace hk {\n")
    for name, p in field_list():
        fg = FieldGen(name, p)
        n2 = fg.N2
        R = fg.R
        out.append(f"// ---------------- {name}: {p.bit_length()} bits, {fg.N} x 32-bit limbs ----------------\n")
        out.append(f"struct alignas(16) {name}_t {{ uint64_t v[{n2}]; }};\n")
        out.append(f"struct {name}_params {{\n"
                   f"    static constexpr int N64 = {n2};\n"
                   f"    static constexpr int BITS = {p.bit_length()};\n"
                   f"    static constexpr uint32_t M32 = 0x{fg.m32:08x}u;\n"
                   f"}};\n")
        out.append(f"__device__ __constant__ const uint64_t {name}_MOD[{n2}] = {c_array64(p, n2)};\n")
        out.append(f"__device__ __constant__ const uint64_t {name}_ONE[{n2}] = {c_array64(R % p, n2)};   // R mod p\n")
        out.append(f"__device__ __constant__ const uint64_t {name}_R2[{n2}] = {c_array64(R * R % p, n2)};    // R^2 mod p\n")
        for cname, val, doc in extra_constants().get(name, []):
            out.append(f"__device__ __constant__ const uint64_t {name}_{cname}[{n2}] = {c_array64(val, n2)};   // {doc}\n")
        out.append(f"#define HK_HAS_SQR_{name} {1 if fg.has_sqr else 0}\n")
        gens = [("mul", fg.gen_mul), ("mulk", lambda: fg.gen_mul_kara("dep")), ("add", fg.gen_add), ("sub", fg.gen_sub)]
        if fg.has_sqr:
            gens.append(("sqr", fg.gen_sqr))
        for fname, gen in gens:
            P, outs, ins = gen()
            out.append(fg.cpp_fn(fname, P, outs, ins))
        has_mul2 = 3 * p < (1 << (32 * fg.N - 1))
        out.append(f"#define HK_HAS_MUL2_{name} {1 if has_mul2 else 0}\n")
        if has_mul2:
            P, outs, ins = fg.gen_mul2()
            out.append(fg.cpp_fn("mul2", P, outs, ins))
            P, outs, ins = fg.gen_negp()
            out.append(fg.cpp_fn("negp", P, outs, ins))
        if not has_mul2:
            out.append(f"// no fused dual product: 3p >= 2^{32 * fg.N - 1}\n"
                       f"__device__ __forceinline__ void {name}_mul2({name}_t& r, const {name}_t& a, const {name}_t& b, "
                       f"const {name}_t& c, const {name}_t& d) "
                       f"{{ {name}_t t, u; {name}_mul(t, a, b); {name}_mul(u, c, d); {name}_add(r, t, u); }}\n"
                       f"__device__ __forceinline__ void {name}_negp({name}_t& r, const {name}_t& a) "
                       f"{{ {name}_t z = {{}}; {name}_sub(r, z, a); }}\n")
        if not fg.has_sqr:
            out.append(f"// no dedicated squaring: 3p >= 2^{32 * fg.N} (the carry invariants need two spare bits)\n"
                       f"__device__ __forceinline__ void {name}_sqr({name}_t& r, const {name}_t& a) "
                       f"{{ {name}_mul(r, a, a); }}\n")
        out.append("\n")
    out.append("}  // namespace hk\n")
    with open(path, "w") as f:
        f.write("".join(out))


def selftest(iters=200, seed=1):
    rng = random.Random(seed)
    for name, p in field_list():
        fg = FieldGen(name, p)
        n2 = fg.N2
        R = fg.R
        Rinv = pow(R, -1, p)
        progs = {k: g() for k, g in (("mul", fg.gen_mul), ("add", fg.gen_add), ("sub", fg.gen_sub))}
        if fg.has_sqr:
            progs["sqr"] = fg.gen_sqr()
        edge = [0, 1, 2, p - 1, p - 2, R % p, (R * R) % p, (p - 1) // 2, (1 << (p.bit_length() - 1)),
                (1 << (p.bit_length() - 1)) - 1, p - (1 << 64), (1 << 64) - 1, (1 << 32) - 1]
        edge = [e % p for e in edge]
        cases = [(x, y) for x in edge for y in edge]
        cases += [(rng.randrange(p), rng.randrange(p)) for _ in range(iters)]
        for x, y in cases:
            env = {}
            for k in range(n2):
                env[f"a{k}"] = lanes64(x, n2)[k]
                env[f"b{k}"] = lanes64(y, n2)[k]
            for op, ref in (("mul", x * y * Rinv % p), ("add", (x + y) % p), ("sub", (x - y) % p)):
                P, outs, ins = progs[op]
                Rr = P.run(env)
                got = sum(Rr[o] << (64 * k) for k, o in enumerate(outs))
                assert got == ref, (name, op, hex(x), hex(y), hex(got), hex(ref))
        if fg.has_sqr:
            P, outs, ins = progs["sqr"]
            top = [p - 1, p - 2, (p - 1) // 2, (1 << (p.bit_length() - 1)), (1 << (p.bit_length() - 1)) - 1,
                   p - (1 << 64), p - (1 << 32) - 1, int("f" * (p.bit_length() // 4 - 1), 16) % p]
            for x in edge + top + [rng.randrange(p) for _ in range(3 * iters)] + \
                    [p - 1 - rng.randrange(1 << 40) for _ in range(iters)]:
                env = {f"a{k}": lanes64(x, n2)[k] for k in range(n2)}
                Rr = P.run(env)
                got = sum(Rr[o] << (64 * k) for k, o in enumerate(outs))
                assert got == x * x * Rinv % p, (name, "sqr", hex(x), hex(got))
            nsq = sum(1 for op, _, _ in P.ins if op in ("mul.wide.u32", "mad.wide.u32"))
            print(f"{name}: sqr ok ({nsq} wide multiplies)")
        # Karatsuba product + separate reduction
        P, outs, ins = fg.gen_mul_kara("dep")
        half = 1 << (16 * fg.N)
        kedge = edge + [(half - 1) % p, half % p, ((half - 1) * half) % p, (half + 1) % p]
        for x, y in [(x, y) for x in kedge for y in kedge] + [(rng.randrange(p), rng.randrange(p)) for _ in range(iters)]:
            env = {}
            for k in range(n2):
                env[f"a{k}"] = lanes64(x, n2)[k]
                env[f"b{k}"] = lanes64(y, n2)[k]
            Rr = P.run(env)
            got = sum(Rr[o] << (64 * k) for k, o in enumerate(outs))
            assert got == x * y * Rinv % p, (name, "mulk", hex(x), hex(y))
        nk = sum(1 for op, _, _ in P.ins if op in ("mul.wide.u32", "mad.wide.u32"))
        print(f"{name}: mulk ok ({nk} wide multiplies)")
        # fused dual product (operands up to and including p)
        if 3 * p < (1 << (32 * fg.N - 1)):
            P, outs, ins = fg.gen_mul2()
            Pn, outs_n, _ = fg.gen_negp()
            e4 = [0, 1, p, p - 1, R % p, (p - 1) // 2]
            quads = [(w, x, y, z) for w in e4 for x in e4 for y in e4 for z in e4]
            quads += [tuple(rng.randrange(p + 1) for _ in range(4)) for _ in range(iters)]
            # adversarial carries: operands made of all-ones / zero / one limbs, clipped to <= p
            sat = [p, p - 1, p - (1 << 32), p - (1 << 64) + 1]
            for _ in range(40):
                v = 0
                for i in range(fg.N):
                    v |= rng.choice([M32, M32, 0, 1, rng.getrandbits(32)]) << (32 * i)
                sat.append(min(v & ((1 << p.bit_length()) - 1), p))
            quads += [tuple(rng.choice(sat) for _ in range(4)) for _ in range(3 * iters)]
            for w, x, y, z in quads:
                env = {}
                for nm, v in zip("abcd", (w, x, y, z)):
                    for k in range(n2):
                        env[f"{nm}{k}"] = lanes64(v, n2)[k]
                Rr = P.run(env)
                got = sum(Rr[o] << (64 * k) for k, o in enumerate(outs))
                assert got == (w * x + y * z) * Rinv % p, (name, "mul2", hex(w), hex(x), hex(y), hex(z))
                Rn = Pn.run({f"a{k}": lanes64(w, n2)[k] for k in range(n2)})
                assert sum(Rn[o] << (64 * k) for k, o in enumerate(outs_n)) == p - w
            n2w = sum(1 for op, _, _ in P.ins if op in ("mul.wide.u32", "mad.wide.u32"))
            print(f"{name}: mul2 ok ({n2w} wide multiplies)")
        nmul = sum(1 for op, _, _ in progs["mul"][0].ins if op in ("mul.wide.u32", "mad.wide.u32"))
        print(f"{name}: ok  ({len(cases)} cases; mul = {nmul} wide multiplies, "
              f"{len(progs['mul'][0].ins)} PTX instructions)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..",
                                                  "hekaton_system_b200", "csrc", "field_gen.cuh"))
    ap.add_argument("--selftest", action="store_true")
    args = ap.parse_args()
    if args.selftest:
        selftest()
    else:
        emit_header(args.out)
        print("wrote", os.path.abspath(args.out))
