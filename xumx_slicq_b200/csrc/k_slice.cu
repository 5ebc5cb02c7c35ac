// Stage 1 (Tukey slicing + slice FFT), stage 3b (spectrum gather + slice IFFT) and stage 4
// (50 % overlap-add) of the sliCQT path.
//
// The length-L real slice transform (L = 18060 = 2 * 43*15*14 for the pretrained Bark
// parameters) runs as ONE complex FFT of length N = L/2 = 9030 held in shared memory, computed
// with the Good-Thomas prime-factor algorithm on the 3-D index space 43 x 15 x 14: the three
// factors are pairwise coprime, so there are NO twiddle multiplications between the passes --
// input index n lives at (n mod 43, n mod 15, n mod 14), output index
// k = (210 k1 + 602 k2 + 645 k3) mod 9030 lives at (k1, k2, k3).
//   pass A  43-point symmetric real half-transforms down the first axis (rdft_sym<43>), the real
//           and the imaginary part of a column on separate threads, in place
//   pass B  combine (A_k -/+ i B_k) fused with the 15-point codelet, rows k1 and 43-k1 together
//   pass C  14-point codelet along the contiguous axis
// followed (forward) / preceded (inverse) by the even/odd split that turns the N-point complex
// transform into the L-point real one.
//
// reference: nsgt/slicing.py:7-72 + nsgt/nsgtf.py:40 (forward), nsgt/nsigtf.py:93-103 +
//            nsgt/unslicing.py:33-69 + nsgt/slicq.py:207-230 (inverse); closed forms in DESIGN.md.
#include "slicq_common.cuh"
#include "dft_codelets.cuh"

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
    static constexpr int SB = P3 + 1;  // pitch of axis 2 (float2): odd -> conflict-free pass C
    static constexpr int SA_MIN = P2 * SB;
    // pitch of axis 1: == 14 (mod 16) keeps the (k1, n3) task order of pass B conflict-free
    static constexpr int SA = SA_MIN + ((14 - SA_MIN % 16) + 16) % 16;
    static constexpr int SMEM_ELEMS = P1 * SA;
    static constexpr int I1 = cmodinv(N / P1, P1), I2 = cmodinv(N / P2, P2), I3 = cmodinv(N / P3, P3);
    static SLICQ_DEVFN int pos_in(int n) { return (n % P1) * SA + (n % P2) * SB + (n % P3); }
    static SLICQ_DEVFN int pos_out(int k) {
        return ((k * I1) % P1) * SA + ((k * I2) % P2) * SB + ((k * I3) % P3);
    }
};

template <class PF, bool INV>
SLICQ_DEVFN void pfa_passes(float2* Z) {
    constexpr int P1 = PF::P1, P2 = PF::P2, P3 = PF::P3, SA = PF::SA, SB = PF::SB;
    float* Zf = reinterpret_cast<float*>(Z);
    // ---- pass A
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
    // ---- pass B
    constexpr int H1 = (P1 - 1) / 2;
    for (int t = threadIdx.x; t < (H1 + 1) * P3; t += blockDim.x) {
        const int kk = t / P3, c = t - kk * P3;
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
    // ---- pass C
    for (int t = threadIdx.x; t < P1 * P2; t += blockDim.x) {
        const int a = t / P2, b = t - a * P2;
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

// sum of the windowed bin spectra covering position f (fixed bin order -> deterministic)
SLICQ_DEVFN float2 gather_spectrum(const SlicqDeviceTables& t, const float2* __restrict__ row, int f) {
    const int j0 = t.jlo[f];
    const int cnt = t.jcnt[f];
    float2 acc = make_float2(0.f, 0.f);
    for (int j = j0; j < j0 + cnt; ++j) {
        const int M = __ldg(t.bin_M + j);
        int d = f - __ldg(t.bin_pos + j);
        if (d < 0) d += M;
        const float2 v = row[__ldg(t.bin_coff + j) + d];
        acc.x += v.x;
        acc.y += v.y;
    }
    return acc;
}

}  // namespace

typedef Pfa3<43, 15, 14> Pfa9030;

// ------------------------------------------------------------------------------------------
// stage 1: one CTA per (row, slice).  x -> H (half spectrum, N+1 complex bins)
template <class PF>
__global__ void __launch_bounds__(256, 2) slice_fft_fwd_kernel(const __grid_constant__ SlicqSliceParams p) {
    SLICQ_DYN_SMEM(float2, Z);
    constexpr int N = PF::N;
    const int rsl = blockIdx.x;
    const int rs = p.rs0 + rsl;
    const int row = rs / p.S, k = rs - row * p.S;
    const long long s0 = (p.k0 + k - 1) * (long long)p.t.hop - p.t0;  // x index of slice sample 0
    const float* __restrict__ xr = p.x + row * p.x_row_stride;
    const float* __restrict__ tw = p.t.tukey;
    for (int e = threadIdx.x; e < N; e += blockDim.x) {
        const long long s = s0 + 2 * e;
        const float w0 = __ldg(tw + 2 * e), w1 = __ldg(tw + 2 * e + 1);
        float a = 0.f, b = 0.f;
        if (w0 != 0.f && s >= 0 && s < p.T) a = __ldg(xr + s) * w0;
        if (w1 != 0.f && s + 1 >= 0 && s + 1 < p.T) b = __ldg(xr + s + 1) * w1;
        Z[PF::pos_in(e)] = make_float2(a, b);
    }
    __syncthreads();
    pfa_passes<PF, false>(Z);
    // even/odd split: H[k] = E + w^k O, H[N-k] = conj(E - w^k O)
    float2* __restrict__ H = p.spec + (long long)rsl * p.spec_stride;
    for (int kk = threadIdx.x; kk <= N / 2; kk += blockDim.x) {
        const float2 zk = Z[PF::pos_out(kk)];
        if (kk == 0) {
            H[0] = make_float2(zk.x + zk.y, 0.f);
            H[N] = make_float2(zk.x - zk.y, 0.f);
        } else {
            const float2 zn = Z[PF::pos_out(N - kk)];
            const float2 E = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
            // O = -(i/2) (zk - conj(zn))
            const float2 O = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
            const float2 t = cmul(__ldg(p.t.post_tw + kk), O);
            H[kk] = make_float2(E.x + t.x, E.y + t.y);
            H[N - kk] = make_float2(E.x - t.x, -(E.y - t.y));
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 3b: one CTA per (row, slice).  packed windowed bin spectra T -> slice signal u [L]
template <class PF>
__global__ void __launch_bounds__(256, 2) slice_fft_inv_kernel(const __grid_constant__ SlicqSliceParams p) {
    SLICQ_DYN_SMEM(float2, Z);
    constexpr int N = PF::N;
    const int rsl = blockIdx.x;
    const float2* __restrict__ Trow = p.spec + (long long)rsl * p.spec_stride;
    for (int kk = threadIdx.x; kk <= N / 2; kk += blockDim.x) {
        const float2 rk = gather_spectrum(p.t, Trow, kk);
        const float2 rn = gather_spectrum(p.t, Trow, N - kk);
        if (kk == 0) {
            // imaginary parts of DC / Nyquist are ignored by a C2R transform (nsigtf.py:103)
            Z[PF::pos_in(0)] = make_float2(rk.x + rn.x, rk.x - rn.x);
        } else {
            const float2 E = make_float2(rk.x + rn.x, rk.y - rn.y);   // rk + conj(rn)
            const float2 O = make_float2(rk.x - rn.x, rk.y + rn.y);   // rk - conj(rn)
            const float2 t = cmul_conj(O, __ldg(p.t.post_tw + kk));   // conj(w^k) O
            Z[PF::pos_in(kk)] = make_float2(E.x - t.y, E.y + t.x);       // E + i t
            Z[PF::pos_in(N - kk)] = make_float2(E.x + t.y, t.x - E.y);   // conj(E - i t)
        }
    }
    __syncthreads();
    pfa_passes<PF, true>(Z);
    const float scale = 1.0f / (float)(2 * N);
    float2* __restrict__ U = reinterpret_cast<float2*>(p.u + (long long)rsl * (2 * N));
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float2 z = Z[PF::pos_out(n)];
        U[n] = make_float2(z.x * scale, z.y * scale);
    }
}

// ------------------------------------------------------------------------------------------
// stage 4: overlap-add of the chunk's slices into y.
//   out hop h (samples [h*hop, (h+1)*hop) of the padded signal) = second half of slice h
//                                                                + first half of slice h+1.
//   grid.x = n_rs units, grid.y = 2: y==0 -> "store" job of unit (its hop), y==1 -> "carry"
//   job (first half of the chunk's first slice of a row, whose partner was written earlier).
__global__ void __launch_bounds__(256) overlap_add_kernel(const __grid_constant__ SlicqOlaParams p) {
    const int rsl = blockIdx.x;
    const int rs = p.rs0 + rsl;
    const int row = rs / p.S, k = rs - row * p.S;
    const float* __restrict__ u = p.u + (long long)rsl * p.L;
    float* __restrict__ yr = p.y + row * p.y_row_stride;
    if (blockIdx.y == 0) {
        // hop h = k: u_k[hop + q] (+ u_{k+1}[q] when that slice is in this chunk)
        const bool has_next = (k + 1 < p.S) && (rsl + 1 < p.n_rs);
        const float* __restrict__ un = u + p.L;
        const long long tb = (p.k0 + k) * (long long)p.hop - p.t0;
        for (int q = threadIdx.x; q < p.hop; q += blockDim.x) {
            const long long t = tb + q;
            if (t < 0 || t >= p.length) continue;
            float v = u[p.hop + q];
            if (has_next) v += un[q];
            yr[t] = v;
        }
    } else {
        // first half of slice k belongs to hop k-1: only needed when slice k-1 is NOT in this chunk
        const bool prev_in_chunk = (k > 0) && (rsl > 0);
        if (prev_in_chunk) return;
        if (k == 0) {
            if (p.halo_out == nullptr || p.k0 == 0) return;  // hop -1 of the whole signal: dropped
            float* __restrict__ h = p.halo_out + (long long)row * p.hop;
            for (int q = threadIdx.x; q < p.hop; q += blockDim.x) h[q] = u[q];
            return;
        }
        const long long tb = (p.k0 + k - 1) * (long long)p.hop - p.t0;
        for (int q = threadIdx.x; q < p.hop; q += blockDim.x) {
            const long long t = tb + q;
            if (t < 0 || t >= p.length) continue;
            yr[t] += u[q];
        }
    }
}

// host-side launchers -----------------------------------------------------------------------
extern "C" int slicq_slice_smem_bytes(int L) {
    if (L == 2 * Pfa9030::N) return (int)(Pfa9030::SMEM_ELEMS * sizeof(float2));
    return -1;
}

extern "C" int slicq_launch_slice_fwd(const SlicqSliceParams* p, cudaStream_t s) {
    if (p->n_rs <= 0) return 0;
    if (p->t.L != 2 * Pfa9030::N) return -2;
    const int smem = (int)(Pfa9030::SMEM_ELEMS * sizeof(float2));
    static int attr_done = 0;
    if (!attr_done) { SLICQ_SET_SMEM(slice_fft_fwd_kernel<Pfa9030>, smem); attr_done = 1; }
    SLICQ_LAUNCH(slice_fft_fwd_kernel<Pfa9030>, dim3(p->n_rs), dim3(256), smem, s, *p);
    return (int)cudaGetLastError();
}

extern "C" int slicq_launch_slice_inv(const SlicqSliceParams* p, cudaStream_t s) {
    if (p->n_rs <= 0) return 0;
    if (p->t.L != 2 * Pfa9030::N) return -2;
    const int smem = (int)(Pfa9030::SMEM_ELEMS * sizeof(float2));
    static int attr_done = 0;
    if (!attr_done) { SLICQ_SET_SMEM(slice_fft_inv_kernel<Pfa9030>, smem); attr_done = 1; }
    SLICQ_LAUNCH(slice_fft_inv_kernel<Pfa9030>, dim3(p->n_rs), dim3(256), smem, s, *p);
    return (int)cudaGetLastError();
}

extern "C" int slicq_launch_ola(const SlicqOlaParams* p, cudaStream_t s) {
    if (p->n_rs <= 0) return 0;
    SLICQ_LAUNCH(overlap_add_kernel, dim3(p->n_rs, 2), dim3(256), 0, s, *p);
    return (int)cudaGetLastError();
}
