#!/usr/bin/env python3
"""Generate straight-line, register-resident DFT codelets for the sliCQT kernels.

    python xumx_slicq_b200/csrc/gen_codelets.py            # rewrites dft_codelets.cuh
    python xumx_slicq_b200/csrc/gen_codelets.py --check    # numerically validates every codelet

Two families are emitted into ``dft_codelets.cuh``:

* ``dft<N, INV>(float2 (&v)[N])``   in-place complex DFT of compile-time size N with
  natural-order output.  Built recursively: Good-Thomas prime-factor split where
  the factors are coprime (no twiddles), Cooley-Tukey with constant twiddles
  inside prime powers, radix-2/4 butterflies, and -- for odd primes -- the
  symmetric direct form  X[k], X[p-k] = x0 + sum a_n cos(.) -/+ i sum b_n sin(.)
  with a_n = x[n]+x[p-n], b_n = x[n]-x[p-n]  ((p-1)^2 real FMAs, all twiddles
  immediates).
* ``rdft_sym<P>(const float (&x)[P], float* out, int stride)``  the *real* half of
  that symmetric form for large primes (29..73): given P reals it writes
  out[0] = sum x, out[k] = x0 + sum a_n cos(2 pi n k / P), out[P-k] = sum b_n sin(..)
  (k = 1..(P-1)/2).  A complex DFT-P is two of these (real and imaginary parts,
  run by two threads) plus a combine step done by the consumer (see slicq_fft.cuh).

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
    """Tiny SSA builder; values are variable names (strings) or numeric evaluation."""

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

    def neg(self, a):
        d = self._new(); self.ops.append((d, "neg", (a,))); return d

    def mulc(self, a, c: float):
        if c == 1.0:
            return a
        if c == -1.0:
            return self.neg(a)
        d = self._new(); self.ops.append((d, "mulc", (a, c))); self.flops += 1; return d

    def fmac(self, a, c: float, acc):
        """a*c + acc"""
        if c == 0.0:
            return acc
        d = self._new(); self.ops.append((d, "fmac", (a, c, acc))); self.flops += 1; return d

    # ---- rendering -------------------------------------------------------
    @staticmethod
    def _lit(c: float) -> str:
        s = repr(float.fromhex(float(c).hex()))
        import struct
        f32 = struct.unpack("f", struct.pack("f", c))[0]
        s = f"{f32:.9g}"
        if "e" not in s and "." not in s:
            s += ".0"
        return s + "f"

    def render(self, indent="    ") -> List[str]:
        out = []
        for d, op, a in self.ops:
            if op == "add":
                out.append(f"{indent}const float {d} = {a[0]} + {a[1]};")
            elif op == "sub":
                out.append(f"{indent}const float {d} = {a[0]} - {a[1]};")
            elif op == "neg":
                out.append(f"{indent}const float {d} = -{a[0]};")
            elif op == "mulc":
                out.append(f"{indent}const float {d} = {a[0]} * {self._lit(a[1])};")
            elif op == "fmac":
                out.append(f"{indent}const float {d} = fmaf({a[0]}, {self._lit(a[1])}, {a[2]});")
        return out

    def evaluate(self, env: Dict[str, float]) -> Dict[str, float]:
        import numpy as np
        f = np.float32
        for d, op, a in self.ops:
            if op == "add":
                env[d] = f(env[a[0]] + env[a[1]])
            elif op == "sub":
                env[d] = f(env[a[0]] - env[a[1]])
            elif op == "neg":
                env[d] = f(-env[a[0]])
            elif op == "mulc":
                env[d] = f(env[a[0]] * f(a[1]))
            elif op == "fmac":
                env[d] = f(np.float64(env[a[0]]) * np.float64(f(a[1])) + np.float64(env[a[2]]))
        return env


C = Tuple[str, str]  # complex value = (re name, im name)


def cadd(e: Emitter, a: C, b: C) -> C:
    return (e.add(a[0], b[0]), e.add(a[1], b[1]))


def csub(e: Emitter, a: C, b: C) -> C:
    return (e.sub(a[0], b[0]), e.sub(a[1], b[1]))


def cmul_const(e: Emitter, a: C, c: float, s: float) -> C:
    """a * (c + i s)"""
    if s == 0.0:
        return (e.mulc(a[0], c), e.mulc(a[1], c))
    if c == 0.0:
        # a * (i s) = (-a.im*s, a.re*s)
        return (e.mulc(a[1], -s), e.mulc(a[0], s))
    re = e.fmac(a[1], -s, e.mulc(a[0], c))
    im = e.fmac(a[1], c, e.mulc(a[0], s))
    return (re, im)


def mul_i(e: Emitter, a: C, sign: int) -> C:
    """a * (sign * i)"""
    if sign > 0:
        return (e.neg(a[1]), a[0])
    return (a[1], e.neg(a[0]))


def dft_prime_sym(e: Emitter, x: List[C], sign: int) -> List[C]:
    p = len(x)
    h = (p - 1) // 2
    a = [None] + [cadd(e, x[n], x[p - n]) for n in range(1, h + 1)]
    b = [None] + [csub(e, x[n], x[p - n]) for n in range(1, h + 1)]
    out: List[C] = [None] * p
    sr, si = x[0]
    for n in range(1, h + 1):
        sr = e.add(sr, a[n][0]); si = e.add(si, a[n][1])
    out[0] = (sr, si)
    for k in range(1, h + 1):
        ar, ai = x[0]
        br = bi = None
        for n in range(1, h + 1):
            c = cospi2(n * k, p)
            s = sinpi2(n * k, p)
            ar = e.fmac(a[n][0], c, ar)
            ai = e.fmac(a[n][1], c, ai)
            br = e.mulc(b[n][0], s) if br is None else e.fmac(b[n][0], s, br)
            bi = e.mulc(b[n][1], s) if bi is None else e.fmac(b[n][1], s, bi)
        # forward (sign=-1): X[k] = A - iB ; X[p-k] = A + iB  (B complex)
        if sign < 0:
            out[k] = (e.add(ar, bi), e.sub(ai, br))
            out[p - k] = (e.sub(ar, bi), e.add(ai, br))
        else:
            out[k] = (e.sub(ar, bi), e.add(ai, br))
            out[p - k] = (e.add(ar, bi), e.sub(ai, br))
    return out


def dft(e: Emitter, x: List[C], sign: int) -> List[C]:
    """DFT of the list x with kernel exp(sign*2*pi*i*n*k/N); natural order in and out."""
    n = len(x)
    if n == 1:
        return list(x)
    if n == 2:
        return [cadd(e, x[0], x[1]), csub(e, x[0], x[1])]
    if n == 4:
        s02 = cadd(e, x[0], x[2]); d02 = csub(e, x[0], x[2])
        s13 = cadd(e, x[1], x[3]); d13 = csub(e, x[1], x[3])
        jd = mul_i(e, d13, sign)  # sign*i*(x1-x3)
        return [cadd(e, s02, s13), cadd(e, d02, jd), csub(e, s02, s13), csub(e, d02, jd)]
    fac = factorize(n)
    if len(fac) == 1 and list(fac.values())[0] == 1:
        return dft_prime_sym(e, x, sign)
    if len(fac) > 1:
        # Good-Thomas: n1 = one prime power, n2 = the rest (coprime)
        p = max(fac, key=lambda q: q ** fac[q])  # largest prime-power first
        n1 = p ** fac[p]
        n2 = n // n1
        # input (Ruritanian): x2[a][b] = x[(a*n2 + b*n1) % n]
        inner = []
        for b in range(n2):
            inner.append(dft(e, [x[(a * n2 + b * n1) % n] for a in range(n1)], sign))
        out: List[C] = [None] * n
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
                v = cmul_const(e, v, cospi2(b * k1, n), sign * sinpi2(b * k1, n))
            col_in.append(v)
        col = dft(e, col_in, sign)
        for k2 in range(n2):
            out[k1 + n1 * k2] = col[k2]
    return out


def gen_complex(n: int, inv: bool) -> Tuple[List[str], int]:
    e = Emitter()
    x = [(f"v[{i}].x", f"v[{i}].y") for i in range(n)]
    # read inputs into named scalars first so that in-place writes are safe
    pre = [f"    const float xr{i} = v[{i}].x, xi{i} = v[{i}].y;" for i in range(n)]
    x = [(f"xr{i}", f"xi{i}") for i in range(n)]
    y = dft(e, x, +1 if inv else -1)
    body = pre + e.render()
    for i in range(n):
        body.append(f"    v[{i}] = make_float2({y[i][0]}, {y[i][1]});")
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


HEADER = '''// GENERATED by gen_codelets.py -- do not edit by hand.
// Register-resident DFT codelets for the sliCQT kernels (see gen_codelets.py docstring).
#pragma once

#ifndef SLICQ_DEVFN
#define SLICQ_DEVFN __device__ __forceinline__
#endif

// dft<N, INV>(v): in-place complex DFT, kernel exp(-/+ 2 pi i nk/N), natural order, unnormalised.
template <int N, bool INV> SLICQ_DEVFN void dft(float2 (&v)[N]);
// rdft_sym<P>(x, out, stride): real symmetric half-transform for odd prime P:
//   out[0] = sum_n x[n];  out[k*stride] = x0 + sum a_n cos(2 pi nk/P);  out[(P-k)*stride] = sum b_n sin(2 pi nk/P)
template <int P> SLICQ_DEVFN void rdft_sym(const float (&x)[P], float* out, int stride);

template <> SLICQ_DEVFN void dft<1, false>(float2 (&)[1]) {}
template <> SLICQ_DEVFN void dft<1, true>(float2 (&)[1]) {}
'''


def generate() -> str:
    parts = [HEADER]
    for n in COMPLEX_SIZES:
        for inv in (False, True):
            body, flops = gen_complex(n, inv)
            parts.append(f"// DFT-{n} {'inverse' if inv else 'forward'}: {flops} flops")
            parts.append(f"template <> SLICQ_DEVFN void dft<{n}, {'true' if inv else 'false'}>(float2 (&v)[{n}]) {{")
            parts.extend(body)
            parts.append("}\n")
    for p in REAL_SYM_PRIMES:
        body, flops = gen_real_sym(p)
        parts.append(f"// real symmetric half-DFT, prime {p}: {flops} flops")
        parts.append(f"template <> SLICQ_DEVFN void rdft_sym<{p}>(const float (&x)[{p}], float* out, int stride) {{")
        parts.extend(body)
        parts.append("}\n")
    return "\n".join(parts)


def codelet_flops(n: int) -> int:
    e = Emitter()
    dft(e, [(f"r{i}", f"i{i}") for i in range(n)], -1)
    return e.flops


def choose_split(M: int):
    """FFT plan for one coefficient length M (multiple of 4).
    kind 1: one thread does the whole DFT-M in registers.
    kind 2: two passes A x B through shared memory (Cooley-Tukey, table twiddles), A >= B.
    kind 3: M = P * R with a prime P >= 29: rdft_sym<P> real/imag split + DFT-R."""
    fac = factorize(M)
    big = [p for p in fac if p >= 29]
    if big:
        p = big[0]
        r = M // p
        assert p in REAL_SYM_PRIMES and r in COMPLEX_SIZES, M
        return (3, p, r)
    if M in COMPLEX_SIZES:
        return (1, M, 1)
    best = None
    for a in COMPLEX_SIZES:
        if M % a:
            continue
        b = M // a
        if b not in COMPLEX_SIZES or b > a:
            continue
        cost = codelet_flops(a) * b + codelet_flops(b) * a
        if best is None or cost < best[0]:
            best = (cost, a, b)
    assert best is not None, M
    return (2, best[1], best[2])


def generate_sizes() -> str:
    lines = ["// GENERATED by gen_codelets.py -- FFT plans per coefficient length M.",
             "// SLICQ_FFT_SIZE(M, KIND, A, B): kind 1 single-thread, 2 two-pass AxB, 3 prime P=A times R=B"]
    for M in range(16, SLICQ_MAX_M + 1, 4):
        k, a, b = choose_split(M)
        lines.append(f"SLICQ_FFT_SIZE({M}, {k}, {a}, {b})")
    return "\n".join(lines) + "\n"


SLICQ_MAX_M = 292


def check() -> None:
    import numpy as np
    rs = np.random.RandomState(0)
    worst = 0.0
    for n in COMPLEX_SIZES:
        for inv in (False, True):
            e = Emitter()
            x = [(f"xr{i}", f"xi{i}") for i in range(n)]
            y = dft(e, x, +1 if inv else -1)
            vals = rs.randn(n) + 1j * rs.randn(n)
            env = {}
            for i in range(n):
                env[f"xr{i}"] = np.float32(vals[i].real)
                env[f"xi{i}"] = np.float32(vals[i].imag)
            env = e.evaluate(env)
            got = np.asarray([complex(env[r], env[i]) for r, i in y])
            v32 = np.asarray([complex(env[f"xr{i}"], env[f"xi{i}"]) for i in range(n)])
            ref = np.fft.ifft(v32) * n if inv else np.fft.fft(v32)
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
        env = e.evaluate(env)
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
