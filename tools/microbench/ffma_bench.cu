// Microbenchmark: FP32 FMA issue rates on sm_100a -- scalar FFMA (register / immediate operand),
// packed FFMA2 (fma.rn.f32x2), and FFMA mixed with integer ALU work.  Prints FMA/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, int n) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    int acc = threadIdx.x;
    for (int it = 0; it < n; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f);            // reg operand
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], 0.999f, a[(i + 1) & 15]);  // imm operand
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                v = __ffma2_rn(v, make_float2(s, s), make_float2(0.5f, 0.25f));
                a[i] = v.x; a[i + 1] = v.y;
            }
        } else if (MODE == 3) {  // FFMA + equal number of integer ops
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = fmaf(a[i], s, 0.5f); acc = (acc ^ (acc >> 3)) + i; }
        } else if (MODE == 4) {  // FFMA2 + integer ops (same FMA count as mode 3)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = make_float2(a[i], a[i + 1]);
                v = __ffma2_rn(v, make_float2(s, s), make_float2(0.5f, 0.25f));
                a[i] = v.x; a[i + 1] = v.y;
                acc = (acc ^ (acc >> 3)) + i; acc = (acc ^ (acc >> 3)) + i + 1;
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
}

template <int MODE> void run(const char* name, float* d, int blocks) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 0.999f, 16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 0.999f, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * 256 * 16.0 * ITERS;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.2f TFMA/s  (~%.1f FMA/clk/SM at %d MHz nominal)\n", name, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    for (int rep = 0; rep < 2; ++rep) {
        run<0>("FFMA reg operand", d, 148 * 8);
        run<1>("FFMA imm operand", d, 148 * 8);
        run<2>("FFMA2 packed", d, 148 * 8);
        run<3>("FFMA + int ALU 1:2", d, 148 * 8);
        run<4>("FFMA2 + int ALU", d, 148 * 8);
    }
    return 0;
}
