// Stage 1 (Tukey slicing + slice FFT), stage 3b (spectrum gather + slice IFFT) and stage 4
// (50 % overlap-add) of the sliCQT path.
//
// The length-L real slice transform (L = 18060 = 2 * 43*15*14 for the pretrained Bark
// parameters) runs as ONE complex FFT of length N = L/2 = 9030 held in shared memory, computed
// with the Good-Thomas prime-factor algorithm on the 3-D index space 43 x 15 x 14: the three
// factors are pairwise coprime, so there are NO twiddle multiplications between the passes --
// input index n lives at (n mod 43, n mod 15, n mod 14), output index
// k = (210 k1 + 602 k2 + 645 k3) mod 9030 lives at (k1, k2, k3).  The load / store phases visit
// indices tid + i * blockDim and keep the three residues of the index in registers (Pfa3::Walk).
//   pass A  43-point symmetric real half-transforms down the first axis (rdft_sym<43>), the real
//           and the imaginary part of a column on separate threads, in place
//   pass B  combine (A_k -/+ i B_k) fused with the 15-point codelet, rows k1 and 43-k1 together
//   pass C  14-point codelet along the contiguous axis
// followed (forward) / preceded (inverse) by the even/odd split that turns the N-point complex
// transform into the L-point real one.
//
// Shared-memory layout: element (a, b, c) at a*SA + b*SB + c (float2 units) with SB = 15 and
// SA = 227: SA odd makes the "lane <-> a" task orders of passes B and C conflict-free for 64-bit
// accesses, and SA + SB + 1 odd does the same for the permuted load / store phases.
//
// reference: nsgt/slicing.py:7-72 + nsgt/nsgtf.py:40 (forward), nsgt/nsigtf.py:93-103 +
//            nsgt/unslicing.py:33-69 + nsgt/slicq.py:207-230 (inverse); closed forms in DESIGN.md.
#include "slicq_common.cuh"
#include "dft_codelets.cuh"

// optional per-phase timing (tuning builds only, -DSLICQ_PHASE_TIMING): thread 0 of every CTA stores
// clock64() at the phase boundaries into a buffer registered with slicq_debug_set_timing()
// (tools/phase_timing.py).  Compiled out of the product build.
#if defined(SLICQ_PHASE_TIMING) && !defined(SLICQ_EMU)
__device__ long long* g_phase_buf = nullptr;
#define PHASE_MARK(i) do { if (threadIdx.x == 0 && g_phase_buf) g_phase_buf[(long long)blockIdx.x * 16 + (i)] = clock64(); } while (0)
#define PHASE_CLOCK() clock64()
#define PHASE_STORE_T(t, i, v) do { if (threadIdx.x == (t) && g_phase_buf) g_phase_buf[(long long)blockIdx.x * 16 + (i)] = (v); } while (0)
#define PHASE_STORE(i, v) do { if (threadIdx.x == 0 && g_phase_buf) g_phase_buf[(long long)blockIdx.x * 16 + (i)] = (v); } while (0)
extern "C" int slicq_debug_set_timing(long long* buf) { return (int)cudaMemcpyToSymbol(g_phase_buf, &buf, sizeof buf); }
#else
#define PHASE_MARK(i) do {} while (0)
#define PHASE_CLOCK() 0LL
#define PHASE_STORE(i, v) do {} while (0)
#define PHASE_STORE_T(t, i, v) do {} while (0)
extern "C" int slicq_debug_set_timing(long long*) { return -1; }
#endif

#ifndef SLICQ_GATHER_UG
#define SLICQ_GATHER_UG 3   // spectrum pairs per thread and round of the synthesis gather
#endif
#ifndef SLICQ_FWD_LOADS
#define SLICQ_FWD_LOADS 6
#endif
#ifndef SLICQ_SLICE_THREADS
#define SLICQ_SLICE_THREADS 384
#endif

namespace {

constexpr int cmodinv(int a, int m) {
    a %= m;
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 0;
}

template <int P1_, int P2_, int P3_>
struct Pfa3 {
    static constexpr int P1 = P1_, P2 = P2_, P3 = P3_;
    static constexpr int N = P1 * P2 * P3;
    static constexpr int SB = (P3 % 2 == 0) ? P3 + 1 : P3 + 2;              // odd
    static constexpr int SA = (P2 * SB) % 2 == 1 ? P2 * SB + 2 : P2 * SB + 1;  // odd, > P2*SB
    static constexpr int SMEM_ELEMS = P1 * SA;
    static constexpr int I1 = cmodinv(N / P1, P1), I2 = cmodinv(N / P2, P2), I3 = cmodinv(N / P3, P3);
    // shared-memory slot of FFT input index n / output index k: a few integer multiplies (constant
    // divisors) instead of a table lookup, so the permuted phases have no dependent memory access
    static __host__ SLICQ_DEVFN int pos_in(int n) { return (n % P1) * SA + (n % P2) * SB + (n % P3); }
    static __host__ SLICQ_DEVFN int pos_out(int k) { return ((k * I1) % P1) * SA + ((k * I2) % P2) * SB + ((k * I3) % P3); }
    // Residue walker: the load / store phases visit indices tid + i * blockDim, so they keep the three
    // residues of the index and advance them by a constant (one add and one conditional subtract
    // each) instead of dividing three times per element.
    struct Walk { int a, b, c; };
    static SLICQ_DEVFN Walk walk_in(int n) { Walk w; w.a = n % P1; w.b = n % P2; w.c = n % P3; return w; }
    static SLICQ_DEVFN Walk walk_out(int k) { Walk w; w.a = (k * I1) % P1; w.b = (k * I2) % P2; w.c = (k * I3) % P3; return w; }
    static SLICQ_DEVFN Walk add_res(Walk w, int da, int db, int dc) {   // da < P1, db < P2, dc < P3
        w.a += da; if (w.a >= P1) w.a -= P1;
        w.b += db; if (w.b >= P2) w.b -= P2;
        w.c += dc; if (w.c >= P3) w.c -= P3;
        return w;
    }
    static SLICQ_DEVFN Walk step_in(Walk w, int d) { return add_res(w, d % P1, d % P2, d % P3); }     // index + d, d >= 0
    static SLICQ_DEVFN Walk step_out(Walk w, int d) { return add_res(w, ((d % P1) * I1) % P1, ((d % P2) * I2) % P2, ((d % P3) * I3) % P3); }
    static SLICQ_DEVFN Walk neg(Walk w) {   // residues of N - index
        w.a = w.a ? P1 - w.a : 0; w.b = w.b ? P2 - w.b : 0; w.c = w.c ? P3 - w.c : 0;
        return w;
    }
    static SLICQ_DEVFN int slot(Walk w) { return w.a * SA + w.b * SB + w.c; }
};
template <class PF, bool INV>
SLICQ_DEVFN void pfa_passes(float2* Z) {
    constexpr int P1 = PF::P1, P2 = PF::P2, P3 = PF::P3, SA = PF::SA, SB = PF::SB;
    float* Zf = reinterpret_cast<float*>(Z);
    // ---- pass A: lane <-> (column, re|im): consecutive words of one row
    for (int t = threadIdx.x; t < P2 * P3 * 2; t += blockDim.x) {
        const int c2 = t & 1, col = t >> 1;
        const int b = col / P3, c = col - b * P3;
        float* base = Zf + (b * SB + c) * 2 + c2;
        float x[P1];
#pragma unroll
        for (int a = 0; a < P1; ++a) x[a] = base[a * SA * 2];
        rdft_sym<P1>(x, base, SA * 2);
    }
    __syncthreads();
    PHASE_MARK(2);
    // ---- pass B: lane <-> kk (row pair), pitch SA odd
    constexpr int H1 = (P1 - 1) / 2;
    for (int t = threadIdx.x; t < (H1 + 1) * P3; t += blockDim.x) {
        const int c = t / (H1 + 1), kk = t - c * (H1 + 1);
        float2* rp = Z + kk * SA + c;
        float2* rm = Z + (P1 - kk) * SA + c;
        float2 xp[P2], xm[P2];
#pragma unroll
        for (int b = 0; b < P2; ++b) {
            const float2 a = rp[b * SB];
            if (kk == 0) {
                xp[b] = a;
            } else {
                const float2 q = rm[b * SB];  // (B_re, B_im)
                if (!INV) {
                    xp[b] = make_float2(a.x + q.y, a.y - q.x);
                    xm[b] = make_float2(a.x - q.y, a.y + q.x);
                } else {
                    xp[b] = make_float2(a.x - q.y, a.y + q.x);
                    xm[b] = make_float2(a.x + q.y, a.y - q.x);
                }
            }
        }
        dft<P2, INV>(xp);
#pragma unroll
        for (int b = 0; b < P2; ++b) rp[b * SB] = xp[b];
        if (kk != 0) {
            dft<P2, INV>(xm);
#pragma unroll
            for (int b = 0; b < P2; ++b) rm[b * SB] = xm[b];
        }
    }
    __syncthreads();
    PHASE_MARK(3);
    // ---- pass C: lane <-> a, pitch SA odd
    for (int t = threadIdx.x; t < P1 * P2; t += blockDim.x) {
        const int b = t / P1, a = t - b * P1;
        float2* r = Z + a * SA + b * SB;
        float2 v[P3];
#pragma unroll
        for (int c = 0; c < P3; ++c) v[c] = r[c];
        dft<P3, INV>(v);
#pragma unroll
        for (int c = 0; c < P3; ++c) r[c] = v[c];
    }
    __syncthreads();
}

}  // namespace

typedef Pfa3<43, 15, 14> Pfa9030;

// ------------------------------------------------------------------------------------------
// stage 1: one CTA per (row, slice).  x -> padded half spectrum H_ext[pad_l + f], f in [-pad_l, N + pad_r]
template <class PF>
__global__ void __launch_bounds__(SLICQ_SLICE_THREADS, 2) slice_fft_fwd_kernel(const __grid_constant__ SlicqSliceParams p) {
    SLICQ_DYN_SMEM(float2, Z);
    constexpr int N = PF::N;
    const int rsl = blockIdx.x;
    const int rs = p.rs0 + rsl;
    const int row = rs / p.S, k = rs - row * p.S;
    const long long s0 = (p.k0 + k - 1) * (long long)p.t.hop - p.t0;  // x index of slice sample 0
    PHASE_MARK(0);
    const float* __restrict__ xr = p.x + row * p.x_row_stride;
    const float2* __restrict__ tw2 = reinterpret_cast<const float2*>(p.t.tukey);
    const int e_lo = p.t.tw_lo >> 1, e_hi = (p.t.tw_hi + 1) >> 1;
    // zero part of the window: no loads
    constexpr int NT = SLICQ_SLICE_THREADS;
    {
        typename PF::Walk wz = PF::walk_in(threadIdx.x);
        for (int e = threadIdx.x; e < e_lo; e += NT) { Z[PF::slot(wz)] = make_float2(0.f, 0.f); wz = PF::step_in(wz, NT); }
        wz = PF::walk_in(e_hi + threadIdx.x);
        for (int e = e_hi + threadIdx.x; e < N; e += NT) { Z[PF::slot(wz)] = make_float2(0.f, 0.f); wz = PF::step_in(wz, NT); }
    }
    PHASE_MARK(7);
    const long long sa = s0 + 2 * e_lo, sb = s0 + 2 * e_hi;   // x range touched by the window support
    const bool interior = sa >= 0 && sb <= p.T;
    const bool vec = ((reinterpret_cast<uintptr_t>(xr + s0) & 7) == 0);
    constexpr int U = SLICQ_FWD_LOADS;   // sample pairs in flight per thread: the phase costs one memory round trip per U * NT pairs
    typename PF::Walk wl = PF::walk_in(e_lo + threadIdx.x);
    if (interior && vec) {
        const float2* __restrict__ x2 = reinterpret_cast<const float2*>(xr + s0);
        for (int e0 = e_lo + threadIdx.x; e0 < e_hi; e0 += U * NT) {
            float2 v[U], w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * NT;
                if (e < e_hi) { w[u] = __ldg(tw2 + e); v[u] = __ldg(x2 + e); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * NT;
                if (e < e_hi) Z[PF::slot(PF::step_in(wl, u * NT))] = make_float2(v[u].x * w[u].x, v[u].y * w[u].y);
            }
            wl = PF::step_in(wl, U * NT);
        }
    } else {
        for (int e = e_lo + threadIdx.x; e < e_hi; e += NT) {
            const long long s = s0 + 2 * e;
            const float2 w = __ldg(tw2 + e);
            float a = 0.f, b = 0.f;
            if (s >= 0 && s < p.T) a = __ldg(xr + s) * w.x;
            if (s + 1 >= 0 && s + 1 < p.T) b = __ldg(xr + s + 1) * w.y;
            Z[PF::slot(wl)] = make_float2(a, b);
            wl = PF::step_in(wl, NT);
        }
    }
    PHASE_MARK(4);
    __syncthreads();
    PHASE_MARK(1);
    pfa_passes<PF, false>(Z);
    PHASE_MARK(5);
    // even/odd split: H[k] = E + w^k O, H[N-k] = conj(E - w^k O); mirrored margins for the bins
    // that reach below DC / above Nyquist (Hermitian symmetry of a real slice)
    float2* __restrict__ H = p.spec + (long long)rsl * p.spec_stride + p.t.pad_l;
    const int pad_l = p.t.pad_l, pad_r = p.t.pad_r;
    const float sc = p.t.spec_scale, se = p.t.ends_scale;
    const float mir = p.t.adjoint ? 0.f : 1.f;     // adjoint mode: positions outside [0, N] read as zero
    constexpr int UP = 4;
    typename PF::Walk wp = PF::walk_out(threadIdx.x);
    for (int kk0 = threadIdx.x; kk0 <= N / 2; kk0 += UP * NT) {
        int pk[UP], pn[UP];
        float2 wk[UP];
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            const int kk = kk0 + u * NT;
            if (kk <= N / 2) {
                const typename PF::Walk wq = PF::step_out(wp, u * NT);
                pk[u] = PF::slot(wq); pn[u] = PF::slot(PF::neg(wq));    // slots of output kk and N - kk (kk = 0: both slot 0)
                wk[u] = __ldg(p.t.post_tw + kk);
            }
        }
        wp = PF::step_out(wp, UP * NT);
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            const int kk = kk0 + u * NT;
            if (kk > N / 2) continue;
            const float2 zk = Z[pk[u]];
            if (kk == 0) {
                H[0] = make_float2((zk.x + zk.y) * se, 0.f);
                H[N] = make_float2((zk.x - zk.y) * se, 0.f);
            } else {
                const float2 zn = Z[pn[u]];
                const float2 E = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                const float2 O = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));  // -(i/2)(zk - conj zn)
                const float2 t = cmul(wk[u], O);
                const float2 hk = make_float2((E.x + t.x) * sc, (E.y + t.y) * sc);
                const float2 hn = make_float2((E.x - t.x) * sc, -(E.y - t.y) * sc);
                H[kk] = hk;
                H[N - kk] = hn;
                if (kk <= pad_l) H[-kk] = make_float2(hk.x * mir, -hk.y * mir);
                if (kk <= pad_r) H[N + kk] = make_float2(hn.x * mir, -hn.y * mir);
            }
        }
    }
    PHASE_MARK(6);
}

// ------------------------------------------------------------------------------------------
// stage 3b + 4: one CTA per (row, slice).  packed windowed bin spectra T -> slice signal u[L],
// overlap-added straight into y:  y[(k-1)*hop + p] += u_k[p].  Every output sample is the sum of
// exactly two slices, one even and one odd: the launch with parity 0 STORES the even slices, the
// launch with parity 1 (stream-ordered after it) ADDS the odd ones with fire-and-forget reductions
// (red.global.add: every location receives exactly one addition, so no ordering question arises) --
// no intermediate slice buffer, and a two-term sum is order independent (bitwise reproducible).
// (reference: nsgt/unslicing.py:33-69, nsgt/slicq.py:207-230)
template <class PF>
__global__ void __launch_bounds__(SLICQ_SLICE_THREADS, 2) slice_fft_inv_kernel(const __grid_constant__ SlicqSliceParams p) {
    SLICQ_DYN_SMEM(float2, Z);
    constexpr int N = PF::N;
    // CTA i of the launch = i-th slice of this launch's parity at or after unit rs0
    const int pi = p.par_base + blockIdx.x;
    const int row = pi / p.par_cs, k = 2 * (pi - row * p.par_cs) + p.parity;
    const int rsl = row * p.S + k - p.rs0;
    PHASE_MARK(0);
    // ---- gather of the windowed bin spectra + Hermitian pre-processing.
    // Spectrum position f is the sum over the bins j = jlo .. jlo + n (n <= 3) covering it of
    // T[f + gd[j]], always in the order (T_jlo + T_jlo+1) + (T_jlo+2 + T_jlo+3): deterministic.
    // Under load a dependent global access costs a full L2/HBM round trip (~1000 cycles), so the phase
    // is organised in as few round trips as possible:
    //   1. tables: gd -> shared memory, this thread's NW pair descriptors and its "extra" entry -> registers
    //   2. the rare third/fourth terms (gx list, < 5 % of the positions) are summed into Z / RN,
    //      overlapped with the loads of round 0
    //   3. NW / UG rounds: the first two terms of UG pairs (k, N-k) and their twiddle in flight together
    int* gd = reinterpret_cast<int*>(Z + ((PF::SMEM_ELEMS + 1) & ~1));
    float2* RN = reinterpret_cast<float2*>(gd + ((p.t.n_bins + 4 + 3) & ~3));
    const int tid = threadIdx.x;
    constexpr int NT = SLICQ_SLICE_THREADS, NW = (N / 2 + NT) / NT, UG = SLICQ_GATHER_UG, NR = (NW + UG - 1) / UG;
#ifdef SLICQ_DEBUG_TWRAP
    const float2* __restrict__ Trow = p.spec + (long long)(rsl % SLICQ_DEBUG_TWRAP) * p.spec_stride;   // tuning experiment, see slicq_fft_tile.cuh
#else
    const float2* __restrict__ Trow = p.spec + (long long)rsl * p.spec_stride;
#endif
    for (int j = tid; j < p.t.n_bins + 4; j += NT) gd[j] = j < p.t.n_bins ? __ldg(p.t.gd + j) : 0;   // 4 pad entries
    unsigned dsc[NR * UG];
#pragma unroll
    for (int u = 0; u < NR * UG; ++u) {
        const int kk = tid + u * NT;
        dsc[u] = (kk <= N / 2) ? __ldg(p.t.gjp + kk) : 0u;
    }
    const int nx = p.t.n_gx;
    int4 xe = make_int4(-1, 0, -1, 0);
    if (tid < nx) xe = __ldg(p.t.gx + tid);
    __syncthreads();
    float2 a[UG][4], w[UG];
    const typename PF::Walk wk0 = PF::walk_in(tid);
    auto load_round = [&](int r) {
#pragma unroll
        for (int u = 0; u < UG; ++u) {
            const int kk = tid + (r * UG + u) * NT;
            if (kk > N / 2) continue;
            const unsigned d = dsc[r * UG + u];
            const int* dk = gd + (d & 0x3fffu);
            const int* dn = gd + ((d >> 16) & 0x3fffu);
            a[u][0] = __ldg(Trow + kk + dk[0]);
            if (((d >> 14) & 3u) >= 1u) a[u][1] = __ldg(Trow + kk + dk[1]);
            a[u][2] = __ldg(Trow + (N - kk) + dn[0]);
            if ((d >> 30) >= 1u) a[u][3] = __ldg(Trow + (N - kk) + dn[1]);
            w[u] = __ldg(p.t.post_tw + kk);
        }
    };
    auto consume_round = [&](int r) {
#pragma unroll
        for (int u = 0; u < UG; ++u) {
            const int kk = tid + (r * UG + u) * NT;
            if (kk > N / 2) continue;
            const unsigned d = dsc[r * UG + u];
            const unsigned nk = (d >> 14) & 3u, nn = d >> 30;
            const typename PF::Walk wkk = PF::step_in(wk0, (r * UG + u) * NT);
            const int pk = PF::slot(wkk), pn = PF::slot(PF::neg(wkk));      // slots of kk and N - kk (kk = 0: both slot 0)
            float2 rk = a[u][0], rn = a[u][2];
            if (nk >= 1u) { rk.x += a[u][1].x; rk.y += a[u][1].y; }
            if (nn >= 1u) { rn.x += a[u][3].x; rn.y += a[u][3].y; }
            if (nk >= 2u) { const float2 e = Z[pk]; rk.x += e.x; rk.y += e.y; }
            if (nn >= 2u) { const float2 e = (kk == 0) ? *RN : Z[pn]; rn.x += e.x; rn.y += e.y; }
            if (kk == 0) {
                // imaginary parts of DC / Nyquist are ignored by a C2R transform (nsigtf.py:103)
                Z[pk] = make_float2(rk.x + rn.x, rk.x - rn.x);
            } else {
                const float2 E = make_float2(rk.x + rn.x, rk.y - rn.y);   // rk + conj(rn)
                const float2 O = make_float2(rk.x - rn.x, rk.y + rn.y);   // rk - conj(rn)
                const float2 t = cmul_conj(O, w[u]);                      // conj(w^k) O
                Z[pk] = make_float2(E.x - t.y, E.y + t.x);       // E + i t
                Z[pn] = make_float2(E.x + t.y, t.x - E.y);       // conj(E - i t)
            }
        }
    };
    // extras of this thread (entry tid of gx; lists longer than the CTA are finished below)
    float2 x2 = make_float2(0.f, 0.f), x3 = make_float2(0.f, 0.f);
    if (xe.x >= 0) { x2 = __ldg(Trow + xe.y); if (xe.z >= 0) x3 = __ldg(Trow + xe.z); }
    load_round(0);
    if (xe.x >= 0) {
        const float2 e = make_float2(x2.x + x3.x, x2.y + x3.y);
        if (xe.x < N) Z[PF::pos_in(xe.x)] = e; else *RN = e;
    }
    for (int i = tid + NT; i < nx; i += NT) {
        const int4 q = __ldg(p.t.gx + i);
        float2 e = __ldg(Trow + q.y);
        if (q.z >= 0) { const float2 v = __ldg(Trow + q.z); e.x += v.x; e.y += v.y; }
        if (q.x < N) Z[PF::pos_in(q.x)] = e; else *RN = e;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        consume_round(r);
        if (r + 1 < NR) load_round(r + 1);
    }
    __syncthreads();
    PHASE_MARK(1);
    pfa_passes<PF, true>(Z);
    PHASE_MARK(5);
    const float scale = 1.0f / (float)(2 * N);
    // slice sample p = 2n, 2n+1 goes to y index tb + p;  first half (n < N/2) = hop k-1, second = hop k
    const long long tb = (p.k0 + k - 1) * (long long)p.t.hop - p.t0;
    float* __restrict__ yr = p.x + row * p.x_row_stride;
    const bool accumulate = p.parity != 0;
    // the last slice of the call has no right neighbour: its second half is stored even when odd
    const bool second_store = !accumulate || (k + 1 >= p.S);
    const bool first_to_halo = (k == 0);      // hop -1: other shard (halo) or before the signal (dropped)
    float* __restrict__ halo = (p.halo_out != nullptr && p.k0 > 0) ? p.halo_out + (long long)row * p.t.hop : nullptr;
    const bool vec = ((reinterpret_cast<uintptr_t>(yr + tb) & 7) == 0) && tb >= 0 && tb + 2 * N <= p.T && !first_to_halo;
    // The odd slices add into hops an even slice has stored (previous launch): reductions
    // (red.global.add, no return value) instead of load + add + store, so that this phase has no
    // global round trip.  y = even + odd either way: bitwise the same two-term sum.
    typename PF::Walk wo = PF::walk_out(threadIdx.x);
    if (vec) {
        const bool add1 = accumulate, add2 = accumulate && !second_store;
        float* __restrict__ yo = yr + tb;
        for (int n = threadIdx.x; n < N; n += SLICQ_SLICE_THREADS) {
            const float2 z0 = Z[PF::slot(wo)];
            wo = PF::step_out(wo, SLICQ_SLICE_THREADS);
            const float2 z = make_float2(z0.x * scale, z0.y * scale);
            if ((n < N / 2) ? add1 : add2) slicq_red_add2(yo + 2 * n, z); else *reinterpret_cast<float2*>(yo + 2 * n) = z;
        }
    } else {
        for (int n = threadIdx.x; n < N; n += SLICQ_SLICE_THREADS) {
            const float2 z0 = Z[PF::slot(wo)];
            wo = PF::step_out(wo, SLICQ_SLICE_THREADS);
            const float2 z = make_float2(z0.x * scale, z0.y * scale);
            const bool first = n < N / 2;
            const bool add = accumulate && (first || !second_store);
            if (first && first_to_halo) {
                if (halo) { halo[2 * n] = z.x; halo[2 * n + 1] = z.y; }
            } else {
                const long long t = tb + 2 * n;
                if (t >= 0 && t < p.T) { if (add) slicq_red_add(yr + t, z.x); else yr[t] = z.x; }
                if (t + 1 >= 0 && t + 1 < p.T) { if (add) slicq_red_add(yr + t + 1, z.y); else yr[t + 1] = z.y; }
            }
        }
    }
    PHASE_MARK(6);
}

// host-side helpers / launchers ---------------------------------------------------------------
extern "C" int slicq_slice_smem_bytes(int L) {
    if (L == 2 * Pfa9030::N) return (int)(Pfa9030::SMEM_ELEMS * sizeof(float2));
    return -1;
}

// dynamic shared memory of slice_fft_inv_kernel: Z, per-bin gather offsets (+ 4 pad entries), Nyquist value
static int slice_inv_smem_bytes(const SlicqDeviceTables& t) {
    size_t b = (size_t)((Pfa9030::SMEM_ELEMS + 1) & ~1) * sizeof(float2);
    b += (size_t)((t.n_bins + 4 + 3) & ~3) * sizeof(int);
    b += 2 * sizeof(float2);
    return (int)b;
}

extern "C" int slicq_launch_slice_fwd(const SlicqSliceParams* p, cudaStream_t s) {
    if (p->n_rs <= 0) return 0;
    if (p->t.L != 2 * Pfa9030::N) return -2;
    const int smem = (int)(Pfa9030::SMEM_ELEMS * sizeof(float2));
    static int attr_done = 0;
    if (!attr_done) { SLICQ_SET_SMEM(slice_fft_fwd_kernel<Pfa9030>, smem); attr_done = 1; }
    SLICQ_LAUNCH(slice_fft_fwd_kernel<Pfa9030>, dim3(p->n_rs), dim3(SLICQ_SLICE_THREADS), smem, s, *p);
    return (int)cudaGetLastError();
}

extern "C" int slicq_launch_slice_inv(const SlicqSliceParams* p, cudaStream_t s) {
    if (p->n_rs <= 0) return 0;
    if (p->t.L != 2 * Pfa9030::N) return -2;
    const int smem = slice_inv_smem_bytes(p->t);
    static int attr_done = 0;
    if (attr_done < smem) { SLICQ_SET_SMEM(slice_fft_inv_kernel<Pfa9030>, smem); attr_done = smem; }
    // slices of parity q among units [0, u) of rows of S slices
    const int S = p->S, q = p->parity, cs = (S + 1 - q) / 2;
    auto count = [&](long long u) { return (u / S) * cs + ((u % S) + 1 - q) / 2; };
    const long long c0 = count(p->rs0), c1 = count((long long)p->rs0 + p->n_rs);
    if (c1 <= c0) return 0;
    SlicqSliceParams sp = *p;
    sp.par_cs = cs; sp.par_base = (int)c0;
    SLICQ_LAUNCH(slice_fft_inv_kernel<Pfa9030>, dim3((unsigned)(c1 - c0)), dim3(SLICQ_SLICE_THREADS), smem, s, sp);
    return (int)cudaGetLastError();
}

