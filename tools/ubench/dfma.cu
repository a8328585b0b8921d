// DFMA vs FFMA issue rate on sm_100a (how expensive are the decimator's fp64 centre taps?)
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
    float a[8]; double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; d[i] = (double)in[threadIdx.x + 8 + i]; }
    const float b = in[64], c = in[65]; const double bd = (double)in[66], cd = (double)in[67];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = fma(d[i], bd, cd);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + (float)d[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
    const int iters = 8192;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, in, iters); else k<1><<<148 * 8, 256>>>(out, in, iters);
            cudaEventRecord(e1); cudaDeviceSynchronize();
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fma = 148.0 * 8 * 256 * (double)iters * 8;
        printf("%s: %.3f ms  %.2f T%s/s\n", mode ? "DFMA" : "FFMA", ms, fma / ms / 1e9, mode ? "DFMA" : "FFMA");
    }
    return 0;
}
