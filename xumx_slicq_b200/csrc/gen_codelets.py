#!/usr/bin/env python3
"""Generate straight-line, register-resident DFT codelets for the sliCQT kernels.

    python xumx_slicq_b200/csrc/gen_codelets.py            # rewrites dft_codelets.cuh
    python xumx_slicq_b200/csrc/gen_codelets.py --check    # numerically validates every codelet

Two families are emitted into ``dft_codelets.cuh`` (plus ``fft_sizes.inc``, the plan per bin length M):

* ``dft<N, INV>(cpx (&v)[N])``   in-place complex DFT of compile-time size N with
  natural-order output, written in PACKED complex operations (slicq_cpx.cuh: one FADD2 / FMUL2 /
  FFMA2 per complex add, real scaling or multiply-accumulate; multiplications by +-i and negations
  are carried as lazy tags and folded into the consuming operation's lane modifiers).
  Built recursively: Good-Thomas prime-factor split where
  the factors are coprime (no twiddles), Cooley-Tukey with constant twiddles
  inside prime powers, radix-2/4 butterflies, and -- for odd primes -- the
  symmetric direct form  X[k], X[p-k] = x0 + sum a_n cos(.) -/+ i sum b_n sin(.)
  with a_n = x[n]+x[p-n], b_n = x[n]-x[p-n]  ((p-1)^2 real FMAs, all twiddles
  immediates).
* ``Dftp<P, NP, INV>``  DFT of prime length P in the same symmetric direct form, STREAMED over the inputs
  (``src.ld(n)``, one pair (n, P-n) at a time) and split BY OUTPUTS over NP threads ("parts"): part q accumulates
  X[k], X[P-k] for its range of k (part 0 also X[0]); the 2 * pairs + 1 packed accumulators are the only long-lived
  registers, every twiddle is an immediate of an FFMA2.  Used for the radix-43 pass of the slice FFT (NP = 2) and the
  bin lengths M = 4 P / 8 P with P = 29 ... 73 (NP = 2 ... 4).

All constants are evaluated in float64 (mpmath-free, math.cos/sin of exact
rational angles reduced to the first octant) and rounded once to fp32.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
from typing import Dict, List, Tuple

# sizes emitted
COMPLEX_SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 22, 23, 24, 28, 32]
REAL_SYM_PRIMES = [29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73]
# streamed prime transforms split by outputs over NP threads: (P, NP)
PRIME_PARTS = [(43, 2), (29, 2), (31, 2), (37, 2), (41, 2), (43, 3), (47, 3), (53, 3), (59, 3), (61, 3),
               (67, 4), (71, 4), (73, 4)]


def cospi2(num: int, den: int) -> float:
    """cos(2*pi*num/den) with exact symmetry reduction."""
    num %= den
    # reduce to [0, den/2]
    if 2 * num > den:
        num = den - num
    # cos(pi - x) = -cos(x)
    if 4 * num > den:
        return -cospi2_q(den - 2 * num, 2 * den)
    return cospi2_q(num, den)


def cospi2_q(num: int, den: int) -> float:
    """cos(2 pi num/den) for 0 <= num/den <= 1/4."""
    if num == 0:
        return 1.0
    if 4 * num == den:
        return 0.0
    if 8 * num > den:  # use sin of complement for accuracy
        return math.sin(2.0 * math.pi * (den - 4 * num) / (4.0 * den))
    return math.cos(2.0 * math.pi * num / den)


def sinpi2(num: int, den: int) -> float:
    """sin(2*pi*num/den) = cos(2 pi (num/den - 1/4))."""
    return cospi2(4 * num - den, 4 * den)


def factorize(n: int) -> Dict[int, int]:
    f: Dict[int, int] = {}
    d = 2
    while d * d <= n:
        while n % d == 0:
            f[d] = f.get(d, 0) + 1
            n //= d
        d += 1
    if n > 1:
        f[n] = f.get(n, 0) + 1
    return f


class Emitter:
    """Scalar SSA builder (real values) used by the real symmetric half transforms."""

    def __init__(self):
        self.ops: List[Tuple[str, str, tuple]] = []  # (dst, op, args)
        self.cnt = 0
        self.flops = 0

    def _new(self) -> str:
        self.cnt += 1
        return f"t{self.cnt}"

    def add(self, a, b):
        d = self._new(); self.ops.append((d, "add", (a, b))); self.flops += 1; return d

    def sub(self, a, b):
        d = self._new(); self.ops.append((d, "sub", (a, b))); self.flops += 1; return d

    def mulc(self, a, c: float):
        d = self._new(); self.ops.append((d, "mulc", (a, c))); self.flops += 1; return d

    def fmac(self, a, c: float, acc):
        """a*c + acc"""
        if c == 0.0:
            return acc
        d = self._new(); self.ops.append((d, "fmac", (a, c, acc))); self.flops += 1; return d

    def render(self, indent="    ") -> List[str]:
        out = []
        for d, op, a in self.ops:
            if op == "add":
                out.append(f"{indent}const float {d} = {a[0]} + {a[1]};")
            elif op == "sub":
                out.append(f"{indent}const float {d} = {a[0]} - {a[1]};")
            elif op == "mulc":
                out.append(f"{indent}const float {d} = {a[0]} * {lit(a[1])};")
            elif op == "fmac":
                out.append(f"{indent}const float {d} = fmaf({a[0]}, {lit(a[1])}, {a[2]});")
        return out


def _scalar_evaluate(e: "Emitter", env: Dict[str, float]) -> Dict[str, float]:
    import numpy as np
    f = np.float32
    for d, op, a in e.ops:
        if op == "add":
            env[d] = f(env[a[0]] + env[a[1]])
        elif op == "sub":
            env[d] = f(env[a[0]] - env[a[1]])
        elif op == "mulc":
            env[d] = f(env[a[0]] * f(a[1]))
        elif op == "fmac":
            env[d] = f(np.float64(env[a[0]]) * np.float64(f(a[1])) + np.float64(env[a[2]]))
    return env


def lit(c: float) -> str:
    import struct
    f32 = struct.unpack("f", struct.pack("f", c))[0]
    s = f"{f32:.9g}"
    if "e" not in s and "." not in s and "inf" not in s and "nan" not in s:
        s += ".0"
    return s + "f"


class CV:
    """A complex SSA value times i^rot (lazy multiplication by +-1, +-i)."""
    __slots__ = ("name", "rot")

    def __init__(self, name: str, rot: int = 0):
        self.name = name
        self.rot = rot & 3

    def times_i(self, k: int = 1) -> "CV":
        return CV(self.name, self.rot + k)


class CEmitter:
    """Complex SSA builder: every op is one packed instruction (slicq_cpx.cuh)."""

    def __init__(self):
        self.ops: List[Tuple[str, str, tuple]] = []
        self.cnt = 0
        self.flops = 0     # packed instructions

    def _emit(self, op: str, args: tuple) -> str:
        self.cnt += 1
        d = f"t{self.cnt}"
        self.ops.append((d, op, args))
        self.flops += 1
        return d

    def add(self, a: CV, b: CV) -> CV:
        op = ("cadd", "caddi", "csub", "csubi")[(b.rot - a.rot) & 3]
        return CV(self._emit(op, (a.name, b.name)), a.rot)

    def sub(self, a: CV, b: CV) -> CV:
        return self.add(a, b.times_i(2))

    def mulr(self, a: CV, c: float) -> CV:
        if c == 1.0:
            return a
        if c == -1.0:
            return a.times_i(2)
        return CV(self._emit("cmulr", (a.name, c)), a.rot)

    def fmar(self, a: CV, c: float, acc: CV) -> CV:
        """acc + c * a"""
        if c == 0.0:
            return acc
        d = (a.rot - acc.rot) & 3
        if d == 0:
            return CV(self._emit("cfmar", (a.name, c, acc.name)), acc.rot)
        if d == 2:
            return CV(self._emit("cfmar", (a.name, -c, acc.name)), acc.rot)
        if d == 1:
            return CV(self._emit("cfmai", (a.name, c, acc.name)), acc.rot)
        return CV(self._emit("cfmai", (a.name, -c, acc.name)), acc.rot)

    def mulw(self, a: CV, c: float, s: float) -> CV:
        """a * (c + i s)"""
        if s == 0.0:
            return self.mulr(a, c)
        if c == 0.0:
            return self.mulr(a.times_i(1), s)
        t = self._emit("cmulr", (a.name, c))
        return CV(self._emit("cfmai", (a.name, s, t)), a.rot)

    def plain(self, a: CV) -> str:
        """materialise the lazy rotation"""
        if a.rot == 0:
            return a.name
        if a.rot == 2:
            return self._emit("cmulr", (a.name, -1.0))
        return self._emit("cmulir", (a.name, 1.0 if a.rot == 1 else -1.0))

    def render(self, indent="    ") -> List[str]:
        out = []
        for d, op, a in self.ops:
            if op in ("cadd", "csub", "caddi", "csubi"):
                out.append(f"{indent}const cpx {d} = {op}({a[0]}, {a[1]});")
            elif op in ("cmulr", "cmulir"):
                out.append(f"{indent}const cpx {d} = {op}({a[0]}, {lit(a[1])});")
            else:
                out.append(f"{indent}const cpx {d} = {op}({a[0]}, {lit(a[1])}, {a[2]});")
        return out

    def evaluate(self, env: Dict[str, complex]) -> Dict[str, complex]:
        """float32 model of the packed operations (fma = one rounding)"""
        import numpy as np
        f = np.float32

        def c32(re, im):
            return complex(f(re), f(im))

        def fma(x, y, z):
            return f(np.float64(x) * np.float64(y) + np.float64(z))

        for d, op, a in self.ops:
            x = env[a[0]]
            if op == "cadd":
                y = env[a[1]]; env[d] = c32(f(x.real) + f(y.real), f(x.imag) + f(y.imag))
            elif op == "csub":
                y = env[a[1]]; env[d] = c32(f(x.real) - f(y.real), f(x.imag) - f(y.imag))
            elif op == "caddi":
                y = env[a[1]]; env[d] = c32(f(x.real) - f(y.imag), f(x.imag) + f(y.real))
            elif op == "csubi":
                y = env[a[1]]; env[d] = c32(f(x.real) + f(y.imag), f(x.imag) - f(y.real))
            elif op == "cmulr":
                env[d] = c32(f(x.real) * f(a[1]), f(x.imag) * f(a[1]))
            elif op == "cmulir":
                env[d] = c32(-f(x.imag) * f(a[1]), f(x.real) * f(a[1]))
            elif op == "cfmar":
                z = env[a[2]]; env[d] = c32(fma(x.real, f(a[1]), z.real), fma(x.imag, f(a[1]), z.imag))
            elif op == "cfmai":
                z = env[a[2]]; env[d] = c32(fma(-f(x.imag), f(a[1]), z.real), fma(x.real, f(a[1]), z.imag))
        return env


def dft_prime_sym(e: CEmitter, x: List[CV], sign: int) -> List[CV]:
    p = len(x)
    h = (p - 1) // 2
    a = [None] + [e.add(x[n], x[p - n]) for n in range(1, h + 1)]
    b = [None] + [e.sub(x[n], x[p - n]) for n in range(1, h + 1)]
    out: List[CV] = [None] * p
    s = x[0]
    for n in range(1, h + 1):
        s = e.add(s, a[n])
    out[0] = s
    for k in range(1, h + 1):
        A = x[0]
        B = None
        for n in range(1, h + 1):
            A = e.fmar(a[n], cospi2(n * k, p), A)
            sv = sinpi2(n * k, p)
            B = e.mulr(b[n], sv) if B is None else e.fmar(b[n], sv, B)
        # forward (sign=-1): X[k] = A - iB ; X[p-k] = A + iB
        if sign < 0:
            out[k] = e.add(A, B.times_i(3)); out[p - k] = e.add(A, B.times_i(1))
        else:
            out[k] = e.add(A, B.times_i(1)); out[p - k] = e.add(A, B.times_i(3))
    return out


def dft(e: CEmitter, x: List[CV], sign: int) -> List[CV]:
    """DFT of the list x with kernel exp(sign*2*pi*i*n*k/N); natural order in and out."""
    n = len(x)
    if n == 1:
        return list(x)
    if n == 2:
        return [e.add(x[0], x[1]), e.sub(x[0], x[1])]
    if n == 4:
        s02 = e.add(x[0], x[2]); d02 = e.sub(x[0], x[2])
        s13 = e.add(x[1], x[3]); d13 = e.sub(x[1], x[3])
        jd = d13.times_i(1 if sign > 0 else 3)  # sign*i*(x1-x3)
        return [e.add(s02, s13), e.add(d02, jd), e.sub(s02, s13), e.sub(d02, jd)]
    fac = factorize(n)
    if len(fac) == 1 and list(fac.values())[0] == 1:
        return dft_prime_sym(e, x, sign)
    if len(fac) > 1:
        # Good-Thomas: n1 = one prime power, n2 = the rest (coprime)
        p = max(fac, key=lambda q: q ** fac[q])  # largest prime-power first
        n1 = p ** fac[p]
        n2 = n // n1
        inner = []
        for b in range(n2):
            inner.append(dft(e, [x[(a * n2 + b * n1) % n] for a in range(n1)], sign))
        out: List[CV] = [None] * n
        i1 = pow(n2, -1, n1)
        i2 = pow(n1, -1, n2)
        for k1 in range(n1):
            col = dft(e, [inner[b][k1] for b in range(n2)], sign)
            for k2 in range(n2):
                out[(k1 * n2 * i1 + k2 * n1 * i2) % n] = col[k2]
        return out
    # prime power: Cooley-Tukey, n = n1*n2 with constant twiddles
    p = list(fac)[0]
    if p == 2:
        n1 = 4 if n >= 16 else 2
        if n == 8:
            n1 = 2
    else:
        n1 = p
    n2 = n // n1
    inner = [dft(e, [x[n2 * a + b] for a in range(n1)], sign) for b in range(n2)]
    out = [None] * n
    for k1 in range(n1):
        col_in = []
        for b in range(n2):
            v = inner[b][k1]
            if b * k1 != 0:
                v = e.mulw(v, cospi2(b * k1, n), sign * sinpi2(b * k1, n))
            col_in.append(v)
        col = dft(e, col_in, sign)
        for k2 in range(n2):
            out[k1 + n1 * k2] = col[k2]
    return out


def gen_complex(n: int, inv: bool) -> Tuple[List[str], int]:
    e = CEmitter()
    pre = [f"    const cpx x{i} = v[{i}];" for i in range(n)]
    y = dft(e, [CV(f"x{i}") for i in range(n)], +1 if inv else -1)
    outs = [e.plain(v) for v in y]
    body = pre + e.render()
    for i in range(n):
        body.append(f"    v[{i}] = {outs[i]};")
    return body, e.flops


def gen_real_sym(p: int) -> Tuple[List[str], int]:
    e = Emitter()
    h = (p - 1) // 2
    x = [f"x[{i}]" for i in range(p)]
    a = [None] + [e.add(x[n], x[p - n]) for n in range(1, h + 1)]
    b = [None] + [e.sub(x[n], x[p - n]) for n in range(1, h + 1)]
    lines: List[str] = []
    s = x[0]
    for n in range(1, h + 1):
        s = e.add(s, a[n])
    stores = [(0, s)]
    for k in range(1, h + 1):
        ak = x[0]
        bk = None
        for n in range(1, h + 1):
            ak = e.fmac(a[n], cospi2(n * k, p), ak)
            sv = sinpi2(n * k, p)
            bk = e.mulc(b[n], sv) if bk is None else e.fmac(b[n], sv, bk)
        stores.append((k, ak))
        stores.append((p - k, bk))
    # interleave stores right after the op that defines them to keep live ranges short
    rendered = e.render()
    defs = {}
    for idx, (d, _, _) in enumerate(e.ops):
        defs[d] = idx
    by_pos: Dict[int, List[str]] = {}
    for (k, name) in stores:
        pos = defs.get(name, -1)
        by_pos.setdefault(pos, []).append(f"    out[{k} * stride] = {name};")
    for idx, line in enumerate(rendered):
        lines.append(line)
        for st in by_pos.get(idx, []):
            lines.append(st)
    for st in by_pos.get(-1, []):
        lines.append(st)
    return lines, e.flops


def part_bounds(p: int, nparts: int) -> List[int]:
    """pairs k = 1..(p-1)/2 dealt to nparts consecutive groups: part q owns k in [kb[q], kb[q+1])"""
    h = (p - 1) // 2
    base, extra = divmod(h, nparts)
    kb = [1]
    for q in range(nparts):
        kb.append(kb[-1] + base + (1 if q < extra else 0))
    assert kb[-1] == h + 1
    return kb


def gen_prime_part(p: int, nparts: int, part: int, inv: bool) -> Tuple[List[str], int, int]:
    """Streamed symmetric direct form of the DFT-p, outputs k in this part's range (and p-k; part 0 also
    k = 0).  Inputs come one pair (n, p-n) at a time from `s.ld(n)`; the accumulators are the only
    long-lived registers.  Returns (body lines, packed instruction count, number of outputs)."""
    h = (p - 1) // 2
    kb = part_bounds(p, nparts)
    ks = list(range(kb[part], kb[part + 1]))
    L: List[str] = []
    ops = 0
    L.append("    const cpx x0 = s.ld(0);")
    if part == 0:
        L.append("    cpx S = x0;")
    for k in ks:
        L.append(f"    cpx A{k} = x0, B{k};")
    for n in range(1, h + 1):
        L.append("    {")
        L.append(f"        const cpx xu = s.ld({n}), xd = s.ld({p - n});")
        L.append("        const cpx a = cadd(xu, xd), b = csub(xu, xd);")
        ops += 2
        if part == 0:
            L.append("        S = cadd(S, a);")
            ops += 1
        for k in ks:
            c = cospi2(n * k, p)
            sv = sinpi2(n * k, p)
            L.append(f"        A{k} = cfmar(a, {lit(c)}, A{k});")
            if n == 1:
                L.append(f"        B{k} = cmulr(b, {lit(sv)});")
            else:
                L.append(f"        B{k} = cfmar(b, {lit(sv)}, B{k});")
            ops += 2
        L.append("    }")
    o = 0
    if part == 0:
        L.append("    o[0] = S;")
        o = 1
    for k in ks:
        # forward: X[k] = A - iB, X[p-k] = A + iB ; inverse: the opposite
        lo, hi = ("caddi", "csubi") if inv else ("csubi", "caddi")
        L.append(f"    o[{o}] = {lo}(A{k}, B{k});")
        L.append(f"    o[{o + 1}] = {hi}(A{k}, B{k});")
        o += 2
        ops += 2
    return L, ops, o


def gen_prime_parts(p: int, nparts: int) -> List[str]:
    out: List[str] = []
    kb = part_bounds(p, nparts)
    nout = max(2 * (kb[q + 1] - kb[q]) + (1 if q == 0 else 0) for q in range(nparts))
    for inv in (False, True):
        tag = "i" if inv else "f"
        for q in range(nparts):
            body, ops, _ = gen_prime_part(p, nparts, q, inv)
            out.append(f"// DFT-{p} {'inverse' if inv else 'forward'}, part {q} of {nparts}: {ops} packed instructions")
            out.append(f"template <class SRC> SLICQ_DEVFN void dftp_{p}_{nparts}_{q}_{tag}(SRC& s, cpx (&o)[{nout}]) {{")
            out.extend(body)
            out.append("}\n")
        out.append(f"template <> struct Dftp<{p}, {nparts}, {'true' if inv else 'false'}> {{")
        out.append(f"    static constexpr int NOUT = {nout};")
        out.append("    template <class SRC> static SLICQ_DEVFN void run(int part, SRC& s, cpx (&o)[NOUT]) {")
        out.append("        switch (part) {")
        for q in range(nparts):
            out.append(f"            case {q}: dftp_{p}_{nparts}_{q}_{tag}(s, o); break;")
        out.append("            default: break;")
        out.append("        }")
        out.append("    }")
        out.append("    // d.st(k, value): output index k of the transform")
        out.append("    template <class DST> static SLICQ_DEVFN void store(int part, DST& d, const cpx (&o)[NOUT]) {")
        out.append("        switch (part) {")
        for q in range(nparts):
            sts = []
            o = 0
            if q == 0:
                sts.append("d.st(0, o[0]);")
                o = 1
            for k in range(kb[q], kb[q + 1]):
                sts.append(f"d.st({k}, o[{o}]); d.st({p - k}, o[{o + 1}]);")
                o += 2
            out.append(f"            case {q}: " + " ".join(sts) + " break;")
        out.append("            default: break;")
        out.append("        }")
        out.append("    }")
        out.append("};\n")
    return out


HEADER = '''// GENERATED by gen_codelets.py -- do not edit by hand.
// Register-resident DFT codelets for the sliCQT kernels (see gen_codelets.py docstring).
#pragma once
#include "slicq_cpx.cuh"

// dft<N, INV>(v): in-place complex DFT, kernel exp(-/+ 2 pi i nk/N), natural order, unnormalised;
// packed complex arithmetic (one FADD2 / FMUL2 / FFMA2 per operation).
template <int N, bool INV> SLICQ_DEVFN void dft(cpx (&v)[N]);
// the same on float2 registers (conversions are register renames)
template <int N, bool INV> SLICQ_DEVFN void dft(float2 (&v)[N]) {
    cpx c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) c[i] = cpx_from(v[i]);
    dft<N, INV>(c);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = cpx_to(c[i]);
}

// Dftp<P, NP, INV>: DFT of prime length P streamed from `SRC::ld(n)` and split BY OUTPUTS over NP threads
// ("parts"; part q accumulates the outputs k and P-k of its pair range, part 0 also k = 0):
//   run(part, src, o)   reads all P inputs (one pair (n, P-n) at a time), leaves the part's outputs in o[]
//   store(part, dst, o) hands them to `DST::st(k, value)`
// Twiddles are immediates, the accumulators the only long-lived registers (2 * pairs + 1 complex).
template <int P, int NP, bool INV> struct Dftp;

template <> SLICQ_DEVFN void dft<1, false>(cpx (&)[1]) {}
template <> SLICQ_DEVFN void dft<1, true>(cpx (&)[1]) {}
'''


def generate() -> str:
    parts = [HEADER]
    for n in COMPLEX_SIZES:
        for inv in (False, True):
            body, flops = gen_complex(n, inv)
            parts.append(f"// DFT-{n} {'inverse' if inv else 'forward'}: {flops} packed instructions")
            parts.append(f"template <> SLICQ_DEVFN void dft<{n}, {'true' if inv else 'false'}>(cpx (&v)[{n}]) {{")
            parts.extend(body)
            parts.append("}\n")
    for (p, nparts) in PRIME_PARTS:
        parts.extend(gen_prime_parts(p, nparts))
    return "\n".join(parts)


def codelet_flops(n: int) -> int:
    e = CEmitter()
    y = dft(e, [CV(f"x{i}") for i in range(n)], -1)
    for v in y:
        e.plain(v)
    return e.flops


def choose_split(M: int):
    """FFT plan for one coefficient length M (multiple of 4).
    kind 1: one thread does the whole DFT-M in registers.
    kind 2: two passes A x B through shared memory (Cooley-Tukey, table twiddles), A >= B.
    kind 3: M = P * R with a prime P >= 29: Dftp<P, NP> (streamed, split by outputs) + DFT-R."""
    fac = factorize(M)
    big = [p for p in fac if p >= 29]
    if big:
        p = big[0]
        r = M // p
        assert p in REAL_SYM_PRIMES and r in COMPLEX_SIZES, M
        return (3, p, r)
    if M in COMPLEX_SIZES:
        return (1, M, 1)
    # two passes A x B through shared memory: B even (the centring shift by M/2 becomes a rotation of pass 2),
    # both factors register friendly; cost = packed instructions of the codelets + twiddle multiplications
    best = None
    for a in COMPLEX_SIZES:
        if M % a or a > 23:
            continue
        b = M // a
        if b not in COMPLEX_SIZES or b % 2 or b > 20 or b < 4:
            continue
        cost = codelet_flops(a) * b + codelet_flops(b) * a + 2 * (a - 1) * (b - 1) + 3 * M
        if a < b:          # pass 2 would run with idle threads (A tasks per transform against B in pass 1)
            cost += 100000
        if best is None or cost < best[0]:
            best = (cost, a, b)
    assert best is not None, M
    return (2, best[1], best[2])


def generate_sizes() -> str:
    lines = ["// GENERATED by gen_codelets.py -- FFT plans per coefficient length M.",
             "// SLICQ_FFT_SIZE(M, KIND, A, B, COST): kind 1 single-thread, 2 two-pass AxB, 3 prime P=A times R=B;",
             "// COST = instructions per transform of the tiled kernels (packed arithmetic + ~14 per element of data movement)"]
    for M in range(16, SLICQ_MAX_M + 1, 4):
        k, a, b = choose_split(M)
        if k == 1:
            cost = codelet_flops(M) + 12 * M
        elif k == 2:
            cost = codelet_flops(a) * b + codelet_flops(b) * a + 3 * M + 14 * M
        else:
            nparts = dict(PRIME_PARTS[1:])[a]
            cost = b * sum(gen_prime_part(a, nparts, q, False)[1] + 2 * a for q in range(nparts)) + a * codelet_flops(b) + 3 * M + 14 * M
        lines.append(f"SLICQ_FFT_SIZE({M}, {k}, {a}, {b}, {cost})")
    return "\n".join(lines) + "\n"


SLICQ_MAX_M = 292


def check() -> None:
    import numpy as np
    rs = np.random.RandomState(0)
    worst = 0.0
    for n in COMPLEX_SIZES:
        for inv in (False, True):
            e = CEmitter()
            y = dft(e, [CV(f"x{i}") for i in range(n)], +1 if inv else -1)
            outs = [e.plain(v) for v in y]
            vals = (rs.randn(n) + 1j * rs.randn(n)).astype(np.complex64)
            env = {f"x{i}": complex(vals[i]) for i in range(n)}
            env = e.evaluate(env)
            got = np.asarray([env[o] for o in outs])
            v64 = vals.astype(np.complex128)
            ref = np.fft.ifft(v64) * n if inv else np.fft.fft(v64)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            worst = max(worst, err)
            assert err < 2e-6, (n, inv, err)
    for p in REAL_SYM_PRIMES:
        e = Emitter()
        h = (p - 1) // 2
        lines, _ = gen_real_sym(p)  # exercise rendering
        # numeric model of rdft_sym through a fresh emitter
        e = Emitter()
        x = [f"x{i}" for i in range(p)]
        a = [None] + [e.add(x[n], x[p - n]) for n in range(1, h + 1)]
        b = [None] + [e.sub(x[n], x[p - n]) for n in range(1, h + 1)]
        vals = rs.randn(p).astype(np.float32)
        env = {f"x{i}": vals[i] for i in range(p)}
        outs = {}
        for k in range(1, h + 1):
            ak = x[0]; bk = None
            for n in range(1, h + 1):
                ak = e.fmac(a[n], cospi2(n * k, p), ak)
                sv = sinpi2(n * k, p)
                bk = e.mulc(b[n], sv) if bk is None else e.fmac(b[n], sv, bk)
            outs[k] = (ak, bk)
        env = _scalar_evaluate(e, env)
        F = np.fft.fft(vals.astype(np.float64))
        for k in range(1, h + 1):
            A = env[outs[k][0]]; B = env[outs[k][1]]
            # X[k] = A - iB for real input
            err = abs(complex(A, -B) - F[k]) / np.abs(F).max()
            worst = max(worst, err)
            assert err < 3e-6, (p, k, err)
    print(f"all codelets OK, worst relative error {worst:.3g}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("-o", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "dft_codelets.cuh"))
    args = ap.parse_args()
    if args.check:
        check()
        return
    src = generate()
    with open(args.o, "w") as f:
        f.write(src)
    print(f"wrote {args.o}: {len(src.splitlines())} lines")
    inc = os.path.join(os.path.dirname(args.o), "fft_sizes.inc")
    with open(inc, "w") as f:
        f.write(generate_sizes())
    print(f"wrote {inc}")


if __name__ == "__main__":
    main()
