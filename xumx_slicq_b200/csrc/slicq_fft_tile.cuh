// Batched small-FFT engine shared by the analysis (per-bin IFFT) and synthesis (per-bin FFT)
// kernels.  One CTA processes a tile of `nf` independent length-M transforms (all bins of one
// bucket x G (row,slice) units).  Three compile-time plans per M (fft_sizes.inc):
//   kind 1  one thread owns a whole DFT-M in registers,
//   kind 2  two passes A x B through shared memory (Cooley-Tukey, table twiddles),
//   kind 3  M = P * R with a prime P >= 29: the DFT-P runs as two real symmetric
//           half-transforms (rdft_sym<P>, real and imaginary part on separate threads)
//           followed by a combine step; the DFT-R is a register codelet.
// Loaders / storers are functors with   Ctx begin(int i)   and  get<M>(ctx, m) / put(ctx, k, v).
// Every loop is a block-stride task loop so the code is blockDim agnostic (see slicq_common.cuh).
#pragma once
#include "slicq_common.cuh"
#include "dft_codelets.cuh"

// ---------------------------------------------------------------------------------------
// kind 1
template <int M, bool INV, class Load, class Store>
SLICQ_DEVFN void fft_tile_single(int nf, const Load& ld, const Store& st) {
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        float2 v[M];
        typename Load::Ctx lc = ld.begin(i);
#pragma unroll
        for (int m = 0; m < M; ++m) v[m] = ld.template get<M>(lc, m);
        dft<M, INV>(v);
        typename Store::Ctx sc = st.begin(i);
#pragma unroll
        for (int k = 0; k < M; ++k) st.put(sc, k, v[k]);
    }
}

// ---------------------------------------------------------------------------------------
// kind 2 : x[B*n1 + n2] -> X[k1 + A*k2]
template <int A, int B>
struct TwoPassLayout {
    static constexpr int BP = (B % 2 == 0) ? B + 1 : B;  // odd row pitch: conflict-free column reads
    static constexpr int PER_FFT = A * BP;               // float2 elements of scratch per transform
};

template <int M, int A, int B, bool INV, class Load, class Store>
SLICQ_DEVFN void fft_tile_two_pass(int nf, float2* sm, const float2* __restrict__ tw, const Load& ld,
                                   const Store& st) {
    static_assert(A * B == M, "bad split");
    typedef TwoPassLayout<A, B> Lay;
    // pass 1: DFT-A over n1 for every (transform, n2), twiddle, park in shared memory
    for (int t = threadIdx.x; t < nf * B; t += blockDim.x) {
        const int i = t / B, n2 = t - i * B;
        float2 v[A];
        typename Load::Ctx lc = ld.begin(i);
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) v[n1] = ld.template get<M>(lc, B * n1 + n2);
        dft<A, INV>(v);
        float2* dst = sm + i * Lay::PER_FFT + n2;
        dst[0] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < A; ++k1) {
            const float2 w = __ldg(tw + n2 * k1);
            dst[k1 * Lay::BP] = INV ? cmul_conj(v[k1], w) : cmul(v[k1], w);
        }
    }
    __syncthreads();
    // pass 2: DFT-B over n2 for every (transform, k1)
    for (int t = threadIdx.x; t < nf * A; t += blockDim.x) {
        const int i = t / A, k1 = t - i * A;
        float2 v[B];
        const float2* src = sm + i * Lay::PER_FFT + k1 * Lay::BP;
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) v[n2] = src[n2];
        dft<B, INV>(v);
        typename Store::Ctx sc = st.begin(i);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) st.put(sc, k1 + A * k2, v[k2]);
    }
}

// ---------------------------------------------------------------------------------------
// kind 3, prime first (analysis side):  x[R*n1 + n2], n1 in [0,P)  ->  X[k1 + P*k2]
//   pass 1: rdft_sym<P> on (transform, n2, re|im)         tasks nf*R*2
//   pass 2: combine + twiddle + DFT-R on (transform, k1)   tasks nf*P   (output runs of P)
template <int P, int R>
struct PrimeLayout {
    static constexpr int PER_FFT = R * 2 * P;  // floats of scratch per transform
};

template <int M, int P, int R, bool INV, class Load, class Store>
SLICQ_DEVFN void fft_tile_prime_first(int nf, float* sm, const float2* __restrict__ tw, const Load& ld,
                                      const Store& st) {
    static_assert(P * R == M, "bad split");
    constexpr int H = (P - 1) / 2;
    for (int t = threadIdx.x; t < nf * R * 2; t += blockDim.x) {
        const int c = t & 1;
        const int u = t >> 1;
        const int i = u / R, n2 = u - i * R;
        float x[P];
        typename Load::Ctx lc = ld.begin(i);
#pragma unroll
        for (int n1 = 0; n1 < P; ++n1) {
            const float2 v = ld.template get<M>(lc, R * n1 + n2);
            x[n1] = c ? v.y : v.x;
        }
        rdft_sym<P>(x, sm + (size_t)(u * 2 + c) * P, 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nf * P; t += blockDim.x) {
        const int i = t / P, k1 = t - i * P;
        const int kk = k1 <= H ? k1 : P - k1;
        float2 v[R];
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) {
            const float* re = sm + (size_t)((i * R + n2) * 2) * P;
            const float* im = re + P;
            float2 y = make_float2(re[kk], im[kk]);
            if (k1 != 0) {
                float br = re[P - kk], bi = im[P - kk];
                // forward: X[kk] = (Ar + Bi, Ai - Br), X[P-kk] = (Ar - Bi, Ai + Br); inverse: swapped
                const bool plus = (k1 <= H) != INV;
                if (!plus) { br = -br; bi = -bi; }
                y.x += bi;
                y.y -= br;
            }
            if (n2 != 0 && k1 != 0) {
                const float2 w = __ldg(tw + n2 * k1);
                y = INV ? cmul_conj(y, w) : cmul(y, w);
            }
            v[n2] = y;
        }
        dft<R, INV>(v);
        typename Store::Ctx sc = st.begin(i);
#pragma unroll
        for (int k2 = 0; k2 < R; ++k2) st.put(sc, k1 + P * k2, v[k2]);
    }
}

// kind 3, prime last (synthesis side):  x[P*n1 + n2], n1 in [0,R)  ->  X[k1 + R*k2], k2 in [0,P)
//   pass 1: DFT-R + twiddle on (transform, n2)             tasks nf*P   (input runs of P)
//   pass 2: rdft_sym<P> in place on (transform, k1, re|im)  tasks nf*R*2
//   pass 3: combine + store on (transform, k2, k1)          tasks nf*P*R
template <int M, int P, int R, bool INV, class Load, class Store>
SLICQ_DEVFN void fft_tile_prime_last(int nf, float* sm, const float2* __restrict__ tw, const Load& ld,
                                     const Store& st) {
    static_assert(P * R == M, "bad split");
    constexpr int H = (P - 1) / 2;
    for (int t = threadIdx.x; t < nf * P; t += blockDim.x) {
        const int i = t / P, n2 = t - i * P;
        float2 v[R];
        typename Load::Ctx lc = ld.begin(i);
#pragma unroll
        for (int n1 = 0; n1 < R; ++n1) v[n1] = ld.template get<M>(lc, P * n1 + n2);
        dft<R, INV>(v);
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
            float2 y = v[k1];
            if (k1 != 0 && n2 != 0) {
                const float2 w = __ldg(tw + n2 * k1);
                y = INV ? cmul_conj(y, w) : cmul(y, w);
            }
            float* re = sm + (size_t)((i * R + k1) * 2) * P;
            re[n2] = y.x;
            re[P + n2] = y.y;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nf * R * 2; t += blockDim.x) {
        float* row = sm + (size_t)t * P;
        float x[P];
#pragma unroll
        for (int n = 0; n < P; ++n) x[n] = row[n];
        rdft_sym<P>(x, row, 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nf * M; t += blockDim.x) {
        const int i = t / M, r = t - i * M;
        const int k2 = r / R, k1 = r - k2 * R;
        const int kk = k2 <= H ? k2 : P - k2;
        const float* re = sm + (size_t)((i * R + k1) * 2) * P;
        const float* im = re + P;
        float2 y = make_float2(re[kk], im[kk]);
        if (k2 != 0) {
            float br = re[P - kk], bi = im[P - kk];
            const bool plus = (k2 <= H) != INV;
            if (!plus) { br = -br; bi = -bi; }
            y.x += bi;
            y.y -= br;
        }
        typename Store::Ctx sc = st.begin(i);
        st.put(sc, k1 + R * k2, y);
    }
}
