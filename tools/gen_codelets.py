#!/usr/bin/env python3
"""Generates go-dsp_b200/csrc/fft_codelets.cuh: straight-line, FMA-form complex128 DFT codelets in registers.

The butterflies of the reference (fft/radix2.go:104-121: t = w * x[odd]; x[even] +/- t, one complex multiply and two
complex additions = 10 flops per radix-2 butterfly) are restated here so that every constant multiplication disappears
into a fused multiply-add: each real value carries a compile-time scale factor that is only applied when the value is
next combined with another one (a + s*b = fma(s, b, a)), which turns a twiddle w = c*(1 + i*t) into two FMAs (the
"tangent" form of Linzer and Feig) and a twiddled radix-4 butterfly into FMAs only. Every output ends with scale 1, so
nothing is left to multiply at the end. Radix-4 decimation in time, trivial twiddles (1, -i, -1, i) cost nothing.

    python tools/gen_codelets.py            # rewrites go-dsp_b200/csrc/fft_codelets.cuh, prints the operation counts
"""
import cmath
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "go-dsp_b200", "csrc", "fft_codelets.cuh")


class Builder:
    def __init__(self):
        self.lines, self.count, self.nops, self.cse = [], 0, {"add": 0, "fma": 0, "mul": 0}, {}

    def new(self, expr, kind):
        key = expr
        if key in self.cse:
            return self.cse[key]
        name = "t%d" % self.count
        self.count += 1
        self.lines.append("    const double %s = %s;" % (name, expr))
        self.nops[kind] += 1
        self.cse[key] = name
        return name

    @staticmethod
    def lit(v):
        return repr(float(v))

    def lin2(self, A, B):
        """A + B for scaled values (scale, var); the result has a positive scale."""
        (sa, na), (sb, nb) = A, B
        # base = the operand with |scale| == 1 if there is one, else the larger |scale| (keeps |ratio| <= 1)
        if abs(abs(sa) - 1.0) > 1e-15 and (abs(abs(sb) - 1.0) <= 1e-15 or abs(sb) > abs(sa)):
            (sa, na), (sb, nb) = (sb, nb), (sa, na)
        s = abs(sa)
        r = sb / s                                  # result = s * (sign(sa)*na + r*nb)
        neg = sa < 0
        if abs(r - 1.0) <= 1e-15:
            expr, kind = ("%s - %s" % (nb, na) if neg else "%s + %s" % (na, nb)), "add"
        elif abs(r + 1.0) <= 1e-15:
            expr, kind = ("-%s - %s" % (na, nb) if neg else "%s - %s" % (na, nb)), "add"
        else:
            expr, kind = "fma(%s, %s, %s%s)" % (self.lit(r), nb, "-" if neg else "", na), "fma"
        return (s, self.new(expr, kind))

    # complex values are pairs of scaled reals
    def cadd(self, x, y):
        return (self.lin2(x[0], y[0]), self.lin2(x[1], y[1]))

    def csub(self, x, y):
        return (self.lin2(x[0], (-y[0][0], y[0][1])), self.lin2(x[1], (-y[1][0], y[1][1])))

    def cmulc(self, z, w):
        """z * w, w a unit-modulus compile-time constant"""
        c, s = w.real, w.imag
        (sr, nr), (si, ni) = z
        if abs(s) < 1e-15:
            return ((c * sr, nr), (c * si, ni))
        if abs(c) < 1e-15:
            return ((-s * si, ni), (s * sr, nr))
        re = self.lin2((c * sr, nr), (-s * si, ni))
        im = self.lin2((c * si, ni), (s * sr, nr))
        return (re, im)

    @staticmethod
    def mul_mi(z):                                   # z * (-i) = (im, -re)
        (sr, nr), (si, ni) = z
        return ((si, ni), (-sr, nr))

    def finish(self, z):
        out = []
        for (s, n) in z:
            if abs(s - 1.0) <= 1e-15:
                out.append(n)
            elif abs(s + 1.0) <= 1e-15:
                out.append("-" + n)
            else:
                out.append(self.new("%s * %s" % (self.lit(s), n), "mul"))
        return out


def root(n, e):
    """exp(-2 pi i e / n), exact on the axes and diagonals"""
    e %= n
    if (8 * e) % n == 0:
        k = 8 * e // n
        h = math.sqrt(0.5)
        return [1, complex(h, -h), -1j, complex(-h, -h), -1, complex(-h, h), 1j, complex(h, h)][k] + 0j
    return cmath.exp(-2j * math.pi * e / n)


def dft(b, x, plan):
    n = len(x)
    if n == 1:
        return x
    r = plan[0] if plan else (4 if n % 4 == 0 else 2)
    m = n // r
    subs = [dft(b, x[j::r], plan[1:]) for j in range(r)]
    out = [None] * n
    for k in range(m):
        t = [subs[0][k]] + [b.cmulc(subs[j][k], root(n, j * k)) for j in range(1, r)]
        if r == 2:
            out[k], out[k + m] = b.cadd(t[0], t[1]), b.csub(t[0], t[1])
        elif r == 4:
            a0, a1 = b.cadd(t[0], t[2]), b.csub(t[0], t[2])
            a2, a3 = b.cadd(t[1], t[3]), b.mul_mi(b.csub(t[1], t[3]))
            out[k], out[k + 2 * m] = b.cadd(a0, a2), b.csub(a0, a2)
            out[k + m], out[k + 3 * m] = b.cadd(a1, a3), b.csub(a1, a3)
        else:
            raise ValueError(r)
    return out


def codelet(n, plan, name, strided):
    b = Builder()
    idx = (lambda i: "v[%d * S]" % i) if strided else (lambda i: "v[%d]" % i)
    x = [((1.0, "%s.x" % idx(i)), (1.0, "%s.y" % idx(i))) for i in range(n)]
    y = dft(b, x, plan)
    fin = [b.finish(z) for z in y]
    head = []
    if strided:
        head.append("template <int S>")
        head.append("__device__ __forceinline__ void %s(cpx* v) {" % name)
    else:
        head.append("__device__ __forceinline__ void %s(cpx (&v)[%d]) {" % (name, n))
    body = list(b.lines)
    for i, (re, im) in enumerate(fin):
        body.append("    %s = make_double2(%s, %s);" % (idx(i), re, im))
    total = sum(b.nops.values())
    doc = "// forward %d-point DFT, natural order in and out: %d FP64 instructions (%d add, %d fma, %d mul), plan %s" % (
        n, total, b.nops["add"], b.nops["fma"], b.nops["mul"], "x".join(map(str, plan)))
    return "\n".join([doc] + head + body + ["}"]), total, b.nops


def check(n, plan):
    """numerical self-check of the generated DAG in Python against a direct DFT"""
    import random
    b = Builder()
    x = [((1.0, "x%dr" % i), (1.0, "x%di" % i)) for i in range(n)]
    y = dft(b, x, plan)
    fin = [b.finish(z) for z in y]
    env = {"fma": lambda a, c, d: a * c + d}
    vals = [complex(random.uniform(-1, 1), random.uniform(-1, 1)) for _ in range(n)]
    for i, v in enumerate(vals):
        env["x%dr" % i], env["x%di" % i] = v.real, v.imag
    for ln in b.lines:
        name, expr = ln.strip()[len("const double "):-1].split(" = ", 1)
        env[name] = eval(expr, {}, env)
    err = 0.0
    for k in range(n):
        got = complex(eval(fin[k][0], {}, env), eval(fin[k][1], {}, env))
        want = sum(vals[j] * root(n, j * k) for j in range(n))
        err = max(err, abs(got - want))
    return err


PLANS = {4: ([4], "dft4_fma"), 8: ([2, 4], "dft8_fma"), 16: ([4, 4], "dft16_fma"), 32: ([4, 2, 4], "dft32_fma")}


def main():
    best = {}
    for n, cands in {8: [[2, 4], [4, 2]], 16: [[4, 4], [2, 2, 4], [2, 4, 2]], 32: [[2, 4, 4], [4, 2, 4], [4, 4, 2]]}.items():
        for p in cands:
            _, tot, _ = codelet(n, p, "x", False)
            if n not in best or tot < best[n][0]:
                best[n] = (tot, p)
            print("n=%d plan %s: %d instructions" % (n, p, tot))
    parts = ["// GENERATED by tools/gen_codelets.py -- do not edit. FMA-form complex128 DFT codelets (see the generator's header).",
             "// Included by fft_core.cuh (after gd::cpx is defined); do not include directly.",
             "#pragma once", "", "namespace gd {", ""]
    for n in (4, 8, 16, 32):
        plan = best[n][1] if n in best else PLANS[n][0]
        name = PLANS[n][1]
        err = check(n, plan)
        assert err < 1e-13 * n, (n, err)
        for strided in ((True,) if n < 32 else (False,)):
            src, tot, nops = codelet(n, plan, name, strided)
            parts += [src, ""]
            print("%s: %d FP64 instructions %r, self-check error %.2e" % (name, tot, nops, err))
    parts += ["}  // namespace gd", ""]
    with open(OUT, "w") as f:
        f.write("\n".join(parts))
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
