// front_stage.cuh — helpers shared by the tile-staged front-end kernels (frontend.cu, classical.cu):
// dB, warp reductions, padded sample reads and the fp32 staging of a tile's samples.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace b2a {

constexpr int kThreads = 256;

__device__ __forceinline__ float db10(float s) {
    // 10*log10(max(amin, s)); lg2.approx abs error 2^-22 -> < 1e-5 dB, 1e-7 of the 80 dB range
    return 3.01029995663981195f * __log2f(fmaxf(s, 1e-10f));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool I16>
__device__ __forceinline__ float load_sample(const void* clip, int s, int n, int pad_mode) {
    if (s < 0 || s >= n) {
        if (pad_mode == 0) return 0.f;
        s = (s < 0) ? -s : 2 * (n - 1) - s;      // np.pad(mode="reflect")
        if (s < 0 || s >= n) return 0.f;
    }
    if (I16) return (float)((const int16_t*)clip)[s] * (1.0f / 32768.0f);
    return ((const float*)clip)[s];
}

// Stage samples [c0, c0+len) of one clip into smem as fp32.  16-byte global loads on the clip's own
// 16-byte grid: the first `head` samples (up to the next boundary) and the tail go one by one, so a
// clip whose start is not a multiple of 8 samples (110 250-sample clips: three out of four) still
// streams through vector loads — only the shared-memory stores narrow to the alignment that is left.
// The loads of a batch are issued before the first conversion (one L2/DRAM round trip per batch).
template <bool I16, int NT = kThreads>
__device__ __forceinline__ void stage_audio(float* __restrict__ dst, const void* __restrict__ clip,
                                            long long clip_elem0, int c0, int len, int n, int pad_mode,
                                            bool base_aligned) {
    constexpr int V = I16 ? 8 : 4;               // elements per 16-byte load
    constexpr int kBatch = 4;
    // samples [c0 + head, ...) start on a 16-byte boundary of the batch (base_aligned: the batch itself does)
    const int head = base_aligned ? (int)((V - ((clip_elem0 + c0) & (V - 1))) & (V - 1)) : len;
    const int hl = head < len ? head : len;
    for (int i = threadIdx.x; i < hl; i += NT) dst[i] = load_sample<I16>(clip, c0 + i, n, pad_mode);
    const int groups = len > hl ? (len - hl + V - 1) / V : 0;
#pragma unroll 1
    for (int g0 = threadIdx.x; g0 < groups; g0 += kBatch * NT) {
        int4 raw[kBatch];
        int state[kBatch];                       // 0: none, 1: vector load, 2: edge (clip boundary, padding, tail)
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int g = g0 + u * NT, i = hl + g * V, sidx = c0 + i;
            state[u] = g < groups ? ((sidx >= 0 && sidx + V <= n && i + V <= len) ? 1 : 2) : 0;
            if (state[u] == 1)
                raw[u] = __ldg(reinterpret_cast<const int4*>(I16 ? (const void*)((const int16_t*)clip + sidx)
                                                                 : (const void*)((const float*)clip + sidx)));
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (state[u] == 0) continue;
            const int g = g0 + u * NT, i = hl + g * V, sidx = c0 + i;
            if (state[u] == 2) {
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (i + e < len) dst[i + e] = load_sample<I16>(clip, sidx + e, n, pad_mode);
                continue;
            }
            float f[V];
            if constexpr (I16) {
                const int r[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    f[2 * e] = (float)(short)(r[e] & 0xffff) * (1.0f / 32768.0f);
                    f[2 * e + 1] = (float)(r[e] >> 16) * (1.0f / 32768.0f);
                }
            } else {
                f[0] = __int_as_float(raw[u].x); f[1] = __int_as_float(raw[u].y);
                f[2] = __int_as_float(raw[u].z); f[3] = __int_as_float(raw[u].w);
            }
            float* d = dst + i;
            if ((hl & 3) == 0) {
#pragma unroll
                for (int e = 0; e < V; e += 4) *reinterpret_cast<float4*>(d + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
            } else if ((hl & 1) == 0) {
#pragma unroll
                for (int e = 0; e < V; e += 2) *reinterpret_cast<float2*>(d + e) = make_float2(f[e], f[e + 1]);
            } else {
#pragma unroll
                for (int e = 0; e < V; ++e) d[e] = f[e];
            }
        }
    }
}

}  // namespace b2a
