// Common definitions for the sliCQT sm_100a kernels.
//
// Build modes:
//   nvcc (product):  real CUDA, sm_100a.
//   g++ -DSLICQ_EMU (tests only): the *same kernel source* is compiled for the host and
//       every CTA is executed by ONE emulated thread (blockDim = 1).  All kernels in
//       this library are written as block-stride task loops separated by
//       __syncthreads(), with no warp-level primitives, so a 1-thread CTA computes
//       bit-for-bit the same task results.  The emulation library is test
//       infrastructure (tests/emu): the product loader never loads it.
#pragma once

#include <stdint.h>
#include <stddef.h>

#ifdef SLICQ_EMU
// ------------------------------------------------------------------ host emulation shim
#include <cmath>
#include <cstdlib>
#include <cstring>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
extern dim3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char* slicq_emu_smem;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
static inline void __syncthreads() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
#define SLICQ_SET_SMEM(kern, bytes) (0)
#define SLICQ_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(slicq_emu_smem)
#define SLICQ_LAUNCH(kern, grid, block, smem, stream, ...)                         \
    do {                                                                           \
        dim3 g_ = (grid);                                                          \
        gridDim = g_; blockDim = dim3(1, 1, 1); threadIdx = dim3(0, 0, 0);         \
        for (unsigned bz_ = 0; bz_ < g_.z; ++bz_)                                  \
            for (unsigned by_ = 0; by_ < g_.y; ++by_)                              \
                for (unsigned bx_ = 0; bx_ < g_.x; ++bx_) {                        \
                    blockIdx = dim3(bx_, by_, bz_);                                \
                    kern(__VA_ARGS__);                                             \
                }                                                                  \
    } while (0)
#else
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
#define SLICQ_SET_SMEM(kern, bytes) \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define SLICQ_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char slicq_smem_raw_[]; \
    type* name = reinterpret_cast<type*>(slicq_smem_raw_)
#define SLICQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#define SLICQ_DEVFN __device__ __forceinline__

#define SLICQ_MAX_BUCKETS 96
#define SLICQ_MAX_M 292

// ---------------------------------------------------------------------------------------
// complex helpers
SLICQ_DEVFN float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
SLICQ_DEVFN float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
SLICQ_DEVFN float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// ---------------------------------------------------------------------------------------
// device-side tables owned by the plan (all in global memory, < 1 MB, L2 resident)
struct SlicqDeviceTables {
    int L;          // slice length (sl_len)
    int N2;         // L / 2 : length of the complex FFT used for the real slice FFT
    int hop;        // L / 2 : slice advance (50 % overlap)
    int n_bins;     // J   (263)
    int n_buckets;  // 70
    int sum_M;      // 18640 coefficients per (row, slice)
    const float* tukey;     // [L]   slicing window
    const float* wf;        // [sum_M] analysis windows  g_j[m] * (-1)^(pos_j/2) / M_j
    const float* wi;        // [sum_M] synthesis windows gd_j[m] * M_j * (-1)^(pos_j/2)
    const int* bin_pos;     // [J]  centre position (rfbas_j)
    const int* bin_M;       // [J]
    const int* bin_coff;    // [J]  offset of bin j inside a packed [sum_M] row
    const float2* post_tw;  // [N2/2 + 1]  exp(-2 pi i k / L)
    const float2* tw;       // concatenated per-bucket twiddles exp(-2 pi i j / M_b), j in [0, M_b)
    const unsigned short* jlo;  // [N2 + 1] first bin covering spectrum position f
    const unsigned char* jcnt;  // [N2 + 1] number of consecutive bins covering f (<= 8)
};

// one bucket as a kernel sees it for one call (pointer + strides of the caller's tensor)
struct SlicqBucketArg {
    float2* ptr;          // base of the bucket tensor (complex64 elements)
    long long s_row;      // element stride between rows   (flattened batch*channel)
    long long s_bin;      // element stride between bins
    long long s_slice;    // element stride between slices ; M is contiguous (stride 1)
    int M;
    int first_bin;
    int n_bins;
    int G;                // (row,slice) units per tile
    int tw_off;           // offset into SlicqDeviceTables::tw
    int tile_start;       // first tile index of this bucket inside the launch
    int kind, A, B;       // FFT plan of this size (see fft_sizes.inc)
    int pad_;
};

struct SlicqBinsParams {
    SlicqDeviceTables t;
    float2* spec;          // forward: half spectra H [n_rs][spec_stride] ; inverse: packed T [n_rs][spec_stride]
    long long spec_stride;
    int n_rs;              // (row,slice) units in this chunk
    int rs0;               // flattened index (row * S + slice) of the first unit of the chunk
    int S;                 // slices per row in this call
    int n_buckets;
    SlicqBucketArg b[SLICQ_MAX_BUCKETS];
};

struct SlicqSliceParams {
    SlicqDeviceTables t;
    // forward: input signal ; inverse: unused
    const float* x;
    long long x_row_stride;
    long long T;           // valid samples in x (per row)
    long long t0;          // global sample index of x[.,0]
    long long k0;          // global slice index of local slice 0
    float2* spec;          // forward: H out ; inverse: T in
    long long spec_stride;
    float* u;              // inverse: slice time signals out [n_rs][L]
    int n_rs, rs0, S;
};

struct SlicqOlaParams {
    const float* u;        // [n_rs][L]
    int L, hop;
    int n_rs, rs0, S;
    float* y;              // [rows][y_row_stride]
    long long y_row_stride;
    long long length;      // valid output samples per row
    long long k0;          // global slice index of local slice 0
    long long t0;          // global sample index of y[.,0]
    float* halo_out;       // [rows][hop] or null: receives the first half of local slice 0 when k0 > 0
    float scale;
};
