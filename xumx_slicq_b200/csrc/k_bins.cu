// Stage 2 (analysis) and stage 3a (synthesis) of the sliCQT path: the ragged per-bin transforms.
//
//  bins_fwd_kernel  (reference: nsgt/nsgtf.py:50-81 + nsgt/slicq.py:13-33 `arrange`)
//      c_j[k, n] = IFFT_M( H_k[(pos_j + m~) ] * wf_j[m] ),  wf_j = g_j * (-1)^(pos_j/2) / M_j
//      H_k = half spectrum of slice k written by slice_fft_fwd_kernel; positions outside
//      [0, L/2] are taken from the Hermitian mirror.  Output goes straight into the caller's
//      ragged bucket tensors [row][bin][slice][M] (any strides, M contiguous).
//  bins_inv_kernel  (reference: nsgt/nsigtf.py:29-33, :82-92)
//      T_k[coff_j + m] = FFT_M( c_j[k, :] )[m] * wi_j[m],  wi_j = gd_j * M_j * (-1)^(pos_j/2)
//      T is the packed [sum_M] row consumed by slice_fft_inv_kernel's gather.
//
// Tiling: one CTA = (bucket, G consecutive (row,slice) units); all 70 buckets run in ONE launch,
// the bucket's compile-time FFT plan is selected by a CTA-uniform switch.
#include "slicq_fft_tile.cuh"

namespace {

struct HLoad {  // analysis: windowed gather from the half spectrum
    const float2* H;
    long long stride;
    const float* wf;
    const int* bin_pos;
    const int* bin_coff;
    int first_bin, g0, ng, N2, L;
    struct Ctx { const float2* row; const float* w; int pos; };
    SLICQ_DEVFN Ctx begin(int i) const {
        const int f = i / ng, g = i - f * ng;
        const int j = first_bin + f;
        Ctx c;
        c.row = H + (long long)(g0 + g) * stride;
        c.w = wf + __ldg(bin_coff + j);
        c.pos = __ldg(bin_pos + j);
        return c;
    }
    template <int M> SLICQ_DEVFN float2 get(const Ctx& c, int m) const {
        const int mt = (m < M / 2) ? m : m - M;
        int idx = c.pos + mt;
        float sgn = 1.f;
        if (idx < 0) { idx = -idx; sgn = -1.f; }
        else if (idx > N2) { idx = L - idx; sgn = -1.f; }
        const float2 v = c.row[idx];
        const float w = __ldg(c.w + m);
        return make_float2(v.x * w, v.y * (w * sgn));
    }
};

struct BucketIO {  // the caller's ragged bucket tensor, used as store (analysis) or load (synthesis)
    float2* ptr;
    long long s_row, s_bin, s_slice;
    int rs0, S, g0, ng;
    struct Ctx { float2* p; };
    SLICQ_DEVFN Ctx begin(int i) const {
        const int f = i / ng, g = i - f * ng;
        const int rs = rs0 + g0 + g;
        const int row = rs / S, k = rs - row * S;
        Ctx c;
        c.p = ptr + row * s_row + f * s_bin + k * s_slice;
        return c;
    }
    SLICQ_DEVFN void put(const Ctx& c, int k, float2 v) const { c.p[k] = v; }
    template <int M> SLICQ_DEVFN float2 get(const Ctx& c, int m) const { return c.p[m]; }
};

struct TStore {  // synthesis: dual-window multiply into the packed row
    float2* T;
    long long stride;
    const float* wi;
    const int* bin_coff;
    int first_bin, g0, ng;
    struct Ctx { float2* p; const float* w; };
    SLICQ_DEVFN Ctx begin(int i) const {
        const int f = i / ng, g = i - f * ng;
        const int off = __ldg(bin_coff + first_bin + f);
        Ctx c;
        c.p = T + (long long)(g0 + g) * stride + off;
        c.w = wi + off;
        return c;
    }
    SLICQ_DEVFN void put(const Ctx& c, int k, float2 v) const {
        const float w = __ldg(c.w + k);
        c.p[k] = make_float2(v.x * w, v.y * w);
    }
};

SLICQ_DEVFN int find_bucket(const SlicqBinsParams& p, int tile) {
    int lo = 0, hi = p.n_buckets - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (p.b[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

}  // namespace

// Two-pass instantiation needs dft<A>, dft<B> only when KIND==2 etc.; `if (KIND == ..)` on template
// constants would still instantiate the dead branches, so dispatch through partial specialisation.
namespace {
template <int M, int KIND, int A, int B, bool SYNTH> struct TileRunner;
template <int M, int A, int B, bool SYNTH> struct TileRunner<M, 1, A, B, SYNTH> {
    template <class Load, class Store>
    static SLICQ_DEVFN void run(int nf, unsigned char*, const float2*, const Load& ld, const Store& st) {
        fft_tile_single<M, !SYNTH>(nf, ld, st);
    }
};
template <int M, int A, int B> struct TileRunner<M, 2, A, B, false> {
    template <class Load, class Store>
    static SLICQ_DEVFN void run(int nf, unsigned char* sm, const float2* tw, const Load& ld, const Store& st) {
        fft_tile_two_pass<M, A, B, true>(nf, reinterpret_cast<float2*>(sm), tw, ld, st);
    }
};
template <int M, int A, int B> struct TileRunner<M, 2, A, B, true> {
    template <class Load, class Store>
    static SLICQ_DEVFN void run(int nf, unsigned char* sm, const float2* tw, const Load& ld, const Store& st) {
        fft_tile_two_pass<M, B, A, false>(nf, reinterpret_cast<float2*>(sm), tw, ld, st);
    }
};
template <int M, int A, int B> struct TileRunner<M, 3, A, B, false> {
    template <class Load, class Store>
    static SLICQ_DEVFN void run(int nf, unsigned char* sm, const float2* tw, const Load& ld, const Store& st) {
        fft_tile_prime_first<M, A, B, true>(nf, reinterpret_cast<float*>(sm), tw, ld, st);
    }
};
template <int M, int A, int B> struct TileRunner<M, 3, A, B, true> {
    template <class Load, class Store>
    static SLICQ_DEVFN void run(int nf, unsigned char* sm, const float2* tw, const Load& ld, const Store& st) {
        fft_tile_prime_last<M, A, B, false>(nf, reinterpret_cast<float*>(sm), tw, ld, st);
    }
};
}  // namespace

__global__ void __launch_bounds__(256) bins_fwd_kernel(const __grid_constant__ SlicqBinsParams p) {
    SLICQ_DYN_SMEM(unsigned char, smem);
    const int tile = blockIdx.x;
    const int bi = find_bucket(p, tile);
    const SlicqBucketArg& b = p.b[bi];
    const int g0 = (tile - b.tile_start) * b.G;
    const int ng = (p.n_rs - g0 < b.G) ? (p.n_rs - g0) : b.G;
    const int nf = b.n_bins * ng;
    HLoad ld;
    ld.H = p.spec; ld.stride = p.spec_stride; ld.wf = p.t.wf; ld.bin_pos = p.t.bin_pos; ld.bin_coff = p.t.bin_coff;
    ld.first_bin = b.first_bin; ld.g0 = g0; ld.ng = ng; ld.N2 = p.t.N2; ld.L = p.t.L;
    BucketIO st;
    st.ptr = b.ptr; st.s_row = b.s_row; st.s_bin = b.s_bin; st.s_slice = b.s_slice;
    st.rs0 = p.rs0; st.S = p.S; st.g0 = g0; st.ng = ng;
    const float2* tw = p.t.tw + b.tw_off;
    switch (b.M) {
#define SLICQ_FFT_SIZE(M_, K_, A_, B_) \
    case M_: TileRunner<M_, K_, A_, B_, false>::run(nf, smem, tw, ld, st); break;
#include "fft_sizes.inc"
#undef SLICQ_FFT_SIZE
        default: break;
    }
}

__global__ void __launch_bounds__(256) bins_inv_kernel(const __grid_constant__ SlicqBinsParams p) {
    SLICQ_DYN_SMEM(unsigned char, smem);
    const int tile = blockIdx.x;
    const int bi = find_bucket(p, tile);
    const SlicqBucketArg& b = p.b[bi];
    const int g0 = (tile - b.tile_start) * b.G;
    const int ng = (p.n_rs - g0 < b.G) ? (p.n_rs - g0) : b.G;
    const int nf = b.n_bins * ng;
    BucketIO ld;
    ld.ptr = b.ptr; ld.s_row = b.s_row; ld.s_bin = b.s_bin; ld.s_slice = b.s_slice;
    ld.rs0 = p.rs0; ld.S = p.S; ld.g0 = g0; ld.ng = ng;
    TStore st;
    st.T = p.spec; st.stride = p.spec_stride; st.wi = p.t.wi; st.bin_coff = p.t.bin_coff;
    st.first_bin = b.first_bin; st.g0 = g0; st.ng = ng;
    const float2* tw = p.t.tw + b.tw_off;
    switch (b.M) {
#define SLICQ_FFT_SIZE(M_, K_, A_, B_) \
    case M_: TileRunner<M_, K_, A_, B_, true>::run(nf, smem, tw, ld, st); break;
#include "fft_sizes.inc"
#undef SLICQ_FFT_SIZE
        default: break;
    }
}

// host-side launchers (called from slicq_api.cu) ------------------------------------------
extern "C" int slicq_launch_bins(const SlicqBinsParams* p, int n_tiles, int smem_bytes, int synth, cudaStream_t s) {
    if (n_tiles <= 0) return 0;
    static int attr_done = 0;
    if (!attr_done) {
        SLICQ_SET_SMEM(bins_fwd_kernel, 160 * 1024);
        SLICQ_SET_SMEM(bins_inv_kernel, 160 * 1024);
        attr_done = 1;
    }
    if (synth) {
        SLICQ_LAUNCH(bins_inv_kernel, dim3(n_tiles), dim3(256), smem_bytes, s, *p);
    } else {
        SLICQ_LAUNCH(bins_fwd_kernel, dim3(n_tiles), dim3(256), smem_bytes, s, *p);
    }
    return (int)cudaGetLastError();
}
