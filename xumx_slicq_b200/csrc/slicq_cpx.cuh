// Packed complex arithmetic for the sliCQT kernels: one complex64 value = one 64-bit register pair,
// every operation = ONE Blackwell packed-fp32 instruction (FADD2 / FMUL2 / FFMA2).
//
// sm_100a executes `add/mul/fma.rn.f32x2` on both halves of a register pair with a single issue slot,
// and the SASS forms carry per-operand modifiers: swap of the two lanes (`.LO_HI`), per-lane negation
// (`.NP` / `.PN`) and broadcast of a 32-bit register or immediate.  ptxas derives those modifiers from
// `mov.b64` packs / unpacks around the PTX instruction, so that with (re, im) in the (lo, hi) lane
//     a + b, a - b, a + i b, a - i b, conj        cost one FADD2,
//     a * r,  acc + a * r,  acc + i r a           cost one FMUL2 / FFMA2 (r real: register or immediate),
//     a * w  (w complex)                          costs FMUL2 + FFMA2
// -- half the issue slots of the scalar formulation (the FP32 lane rate is the same; the slice and bin
// kernels are bound by issue slots and latency, not by the FP32 pipe: profiles/r1_experiments.txt).
//
// -DSLICQ_EMU (tests only) maps the same operations to plain float arithmetic on the host.
#pragma once
#include "slicq_common.cuh"

#ifdef SLICQ_EMU
struct cpx { float x, y; };
#define CPX_FN static inline __attribute__((always_inline))
CPX_FN cpx cpx_make(float re, float im) { cpx r; r.x = re; r.y = im; return r; }
CPX_FN float cpx_re(cpx a) { return a.x; }
CPX_FN float cpx_im(cpx a) { return a.y; }
CPX_FN cpx cadd(cpx a, cpx b) { return cpx_make(a.x + b.x, a.y + b.y); }
CPX_FN cpx csub(cpx a, cpx b) { return cpx_make(a.x - b.x, a.y - b.y); }
CPX_FN cpx caddi(cpx a, cpx b) { return cpx_make(a.x - b.y, a.y + b.x); }      // a + i b
CPX_FN cpx csubi(cpx a, cpx b) { return cpx_make(a.x + b.y, a.y - b.x); }      // a - i b
CPX_FN cpx caddc(cpx a, cpx b) { return cpx_make(a.x + b.x, a.y - b.y); }      // a + conj(b)
CPX_FN cpx csubc(cpx a, cpx b) { return cpx_make(a.x - b.x, a.y + b.y); }      // a - conj(b)
CPX_FN cpx cmulr(cpx a, float r) { return cpx_make(a.x * r, a.y * r); }
CPX_FN cpx cmulir(cpx a, float r) { return cpx_make(-a.y * r, a.x * r); }      // i r a
CPX_FN cpx cfmar(cpx a, float r, cpx acc) { return cpx_make(fmaf(a.x, r, acc.x), fmaf(a.y, r, acc.y)); }    // acc + r a
CPX_FN cpx cfmai(cpx a, float r, cpx acc) { return cpx_make(fmaf(-a.y, r, acc.x), fmaf(a.x, r, acc.y)); }   // acc + i r a
CPX_FN cpx cmul2(cpx a, cpx b) { return cpx_make(a.x * b.x, a.y * b.y); }      // lane-wise product
CPX_FN cpx cfma2(cpx a, cpx b, cpx c) { return cpx_make(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
struct cpx { unsigned long long v; };
#define CPX_FN __device__ __forceinline__
CPX_FN cpx cpx_make(float re, float im) { cpx r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im)); return r; }
CPX_FN float cpx_re(cpx a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
CPX_FN float cpx_im(cpx a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }
CPX_FN cpx cpx_add2_(cpx a, cpx b) { cpx r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
CPX_FN cpx cpx_mul2_(cpx a, cpx b) { cpx r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
CPX_FN cpx cpx_fma2_(cpx a, cpx b, cpx c) { cpx r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
CPX_FN cpx cadd(cpx a, cpx b) { return cpx_add2_(a, b); }
CPX_FN cpx csub(cpx a, cpx b) { cpx r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
CPX_FN cpx caddi(cpx a, cpx b) { return cpx_add2_(a, cpx_make(-cpx_im(b), cpx_re(b))); }
CPX_FN cpx csubi(cpx a, cpx b) { return cpx_add2_(a, cpx_make(cpx_im(b), -cpx_re(b))); }
CPX_FN cpx caddc(cpx a, cpx b) { return cpx_add2_(a, cpx_make(cpx_re(b), -cpx_im(b))); }
CPX_FN cpx csubc(cpx a, cpx b) { return cpx_add2_(a, cpx_make(-cpx_re(b), cpx_im(b))); }
CPX_FN cpx cmulr(cpx a, float r) { return cpx_mul2_(a, cpx_make(r, r)); }
CPX_FN cpx cmulir(cpx a, float r) { return cpx_mul2_(cpx_make(-cpx_im(a), cpx_re(a)), cpx_make(r, r)); }
CPX_FN cpx cfmar(cpx a, float r, cpx acc) { return cpx_fma2_(a, cpx_make(r, r), acc); }
CPX_FN cpx cfmai(cpx a, float r, cpx acc) { return cpx_fma2_(cpx_make(-cpx_im(a), cpx_re(a)), cpx_make(r, r), acc); }
CPX_FN cpx cmul2(cpx a, cpx b) { return cpx_mul2_(a, b); }
CPX_FN cpx cfma2(cpx a, cpx b, cpx c) { return cpx_fma2_(a, b, c); }
#endif

CPX_FN cpx cpx_zero() { return cpx_make(0.f, 0.f); }
CPX_FN cpx cpx_from(float2 a) { return cpx_make(a.x, a.y); }
CPX_FN float2 cpx_to(cpx a) { return make_float2(cpx_re(a), cpx_im(a)); }
CPX_FN cpx cconjp(cpx a) { return cpx_make(cpx_re(a), -cpx_im(a)); }
CPX_FN cpx cnegp(cpx a) { return cpx_make(-cpx_re(a), -cpx_im(a)); }
CPX_FN cpx cswap(cpx a) { return cpx_make(cpx_im(a), cpx_re(a)); }
// a * w and a * conj(w), w = (c, s) complex
CPX_FN cpx cmulw(cpx a, float2 w) { return cfmai(a, w.y, cmulr(a, w.x)); }
CPX_FN cpx cmulwc(cpx a, float2 w) { return cfmai(a, -w.y, cmulr(a, w.x)); }
// 64-bit loads / stores of complex64 data (memory layout = float2)
CPX_FN cpx cpx_ld(const float2* p) {
#ifdef SLICQ_EMU
    return cpx_make(p->x, p->y);
#else
    cpx r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r;
#endif
}
CPX_FN void cpx_st(float2* p, cpx a) {
#ifdef SLICQ_EMU
    p->x = a.x; p->y = a.y;
#else
    *reinterpret_cast<unsigned long long*>(p) = a.v;
#endif
}
