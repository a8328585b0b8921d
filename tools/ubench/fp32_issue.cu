// Micro-benchmark: issue rate of scalar vs packed FP32 on sm_100a (FFMA vs FFMA2 / FADD2),
// alone and mixed with ALU / shared-memory instructions.  Prints warp-instructions per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, long long* cyc) {
    __shared__ float sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = in[i & 255];
    __syncthreads();
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 8 + i]);
    const float2 b = make_float2(in[threadIdx.x + 64], in[threadIdx.x + 65]);
    const float2 c = make_float2(in[threadIdx.x + 66], in[threadIdx.x + 67]);
    unsigned u = threadIdx.x * 7u + 1u;
    const float* sp = sh + (threadIdx.x & 31) * 2;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {              // 16 scalar FFMA (3 register operands)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
        } else if (MODE == 1) {       // 8 FFMA2 = the same flops
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
        } else if (MODE == 2) {       // 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fadd2_rn(a[i], b);
        } else if (MODE == 3) {       // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x += b.x; a[i].y += b.y; }
        } else if (MODE == 4) {       // 8 FFMA2 + 8 LOP3 (ALU pipe)
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = __ffma2_rn(a[i], b, c); u = (u ^ (u >> 3)) + 0x9e3779b9u * i; }
        } else if (MODE == 5) {       // 16 FFMA + 8 ALU
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); u = (u ^ (u >> 3)) + 0x9e3779b9u * i; }
        } else if (MODE == 6) {       // 8 FFMA2 + 4 LDS.64 (conflict-free)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
#pragma unroll
            for (int i = 0; i < 4; ++i) { float2 x = *reinterpret_cast<const float2*>(sp + 64 * ((i + it) & 15)); a[i].x += x.x; a[i + 4].y += x.y; }
        } else if (MODE == 7) {       // 8 FFMA2 with one operand negated + swapped-pair use (a, (b.y,b.x), c)
            const float2 bs = make_float2(-b.y, b.x);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], (i & 1) ? bs : b, c);
        } else if (MODE == 8) {       // 8 FMUL2
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmul2_rn(a[i], b);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)u;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int fp_instr, int other_instr, int ctas_per_sm, float* out, float* in, long long* cyc) {
    int grid = 148 * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, in, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, in, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double c = 0; for (int i = 0; i < grid; ++i) c += (double)h[i]; c /= grid;
    const double warps = 8.0 * ctas_per_sm;
    const double wi_fp = warps * ITER * fp_instr, wi_ot = warps * ITER * other_instr;
    printf("%-34s ctas/SM=%d  cycles=%.0f  fp-instr/clk/SM=%.2f  other/clk/SM=%.2f  total/clk/SM=%.2f  (%.3f ms)\n",
           name, ctas_per_sm, c, wi_fp / c, wi_ot / c, (wi_fp + wi_ot) / c, ms);
}

int main() {
    float *out, *in; long long* cyc;
    CK(cudaMalloc(&out, 148 * 8 * 256 * 4)); CK(cudaMalloc(&in, 4096 * 4)); CK(cudaMalloc(&cyc, 148 * 8 * 8));
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 0.5f + 1e-3f * (i % 17);
    CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    for (int c : {2, 4, 8}) {
        run<0>("16 FFMA (3-reg)", 16, 0, c, out, in, cyc);
        run<1>("8 FFMA2", 8, 0, c, out, in, cyc);
        run<2>("8 FADD2", 8, 0, c, out, in, cyc);
        run<3>("16 FADD", 16, 0, c, out, in, cyc);
        run<8>("8 FMUL2", 8, 0, c, out, in, cyc);
        run<4>("8 FFMA2 + ~24 ALU", 8, 24, c, out, in, cyc);
        run<5>("16 FFMA + ~24 ALU", 16, 24, c, out, in, cyc);
        run<6>("8 FFMA2 + 4 LDS.64 + 8 FADD", 16, 4, c, out, in, cyc);
        run<7>("8 FFMA2 (neg/swapped operand)", 8, 0, c, out, in, cyc);
    }
    return 0;
}
