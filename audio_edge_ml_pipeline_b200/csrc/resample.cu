// resample.cu — rational polyphase resampler (files whose rate differs from `sample_rate`).
//
// Stands in for the soxr_hq resampling librosa.load performs inside `_load_segment`
// (reference src/preprocessing/feature_extraction/audio/deep.py:44-50): zero-phase Kaiser-sinc
// low-pass at the decimator's specification (tables.h: design_resampler), zero-extended edges,
// output length ceil(n * target / orig) as librosa.resample fixes it.  libsoxr is not available
// offline, so parity against its exact output is unpinned; the oracle
// (oracle/librosa_restated.py: resample_restated) evaluates the same closed form through
// scipy.signal.resample_poly.  sm_100a only, no CPU path.
//
//   y[m] = sum_i poly[(m*down + half) % up][i] * x[(m*down + half) / up - i]
//
// Two kernels.  resample_phase_kernel (up >= 8, down odd — 44.1 -> 16 kHz is 160/441): the 32 lanes
// of a warp take outputs m, m + up, m + 2 up, ..., which all use the SAME phase row, so a 128-bit
// tap load is one broadcast wavefront instead of 32, and their input positions are `down` samples
// apart — odd, hence 32 distinct shared-memory banks.  A CTA stages the 32 * down + K input samples
// once and its eight warps take eight phases; blockIdx.y walks the rest.  resample_kernel (everything
// else: up < 8 shares rows between lanes anyway; an even `down` would put every lane on one bank):
// one CTA produces 256 consecutive outputs from a staged window, every thread walks its own phase
// row.  Both: int16 widened while staging, four independent accumulators, identical tap order.
#include "../../include/b2a.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "tables.h"

void b2a_internal_set_error(const char* msg);      // api.cu

namespace {

constexpr int kThreads = 256;

thread_local std::string g_rs_err;

// Batched launches (blockIdx.z = clip): clip c reads in + c * in_stride elements, in_len[c] of them (or
// n_in when in_len is NULL), and owns the row out + c * out_stride of which outputs [0, out_cap) are
// written: the resampled signal while it lasts (ceil(n_in * up / down) samples), zeros after it — the
// pad/trim to a fixed duration that follows librosa.load in the reference (deep.py:52-61).
struct RsBatch {
    long long in_stride, out_stride;
    const int* in_len;
};

template <bool I16>
__global__ void __launch_bounds__(kThreads) resample_kernel(const void* __restrict__ in_, long long n_in,
                                                            float* __restrict__ out_, long long n_out,
                                                            const float* __restrict__ poly, int up, int down,
                                                            int K, int half_len, int win, RsBatch bt) {
    extern __shared__ __align__(16) float s_x[];
    const void* in = (const unsigned char*)in_ + (size_t)blockIdx.z * bt.in_stride * (I16 ? 2 : 4);
    float* out = out_ + (size_t)blockIdx.z * bt.out_stride;
    long long n_real = n_out;                       // outputs the signal really has; [n_real, n_out) are zeros
    if (bt.in_len) {
        n_in = bt.in_len[blockIdx.z];
        n_real = (n_in * up + down - 1) / down;
        if ((long long)blockIdx.x * kThreads >= n_real) {       // whole block past the signal: just the padding
            const long long m = (long long)blockIdx.x * kThreads + threadIdx.x;
            if (m < n_out) out[m] = 0.f;
            return;
        }
    }
    const long long m0 = (long long)blockIdx.x * kThreads;
    // newest sample any output of this CTA touches, oldest = that of the first output minus K-1
    const long long k_first = (m0 * down + half_len) / up;
    const long long k_lo = k_first - (K - 1);
    for (int i = threadIdx.x; i < win; i += kThreads) {
        const long long k = k_lo + i;
        float v = 0.f;
        if (k >= 0 && k < n_in)
            v = I16 ? (float)((const int16_t*)in)[k] * (1.0f / 32768.0f) : ((const float*)in)[k];
        s_x[i] = v;
    }
    __syncthreads();
    const long long m = m0 + threadIdx.x;
    if (m >= n_out) return;
    if (m >= n_real) { out[m] = 0.f; return; }
    const long long u = m * down + half_len;
    const int p = (int)(u % up);
    const float* xs = s_x + (int)(u / up - k_lo);            // x[(u / up) - i] = xs[-i]
    const float4* h4 = reinterpret_cast<const float4*>(poly + (size_t)p * K);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int i = 0; i < K / 4; ++i) {
        const float4 h = __ldg(h4 + i);
        a0 = fmaf(h.x, xs[-4 * i], a0);
        a1 = fmaf(h.y, xs[-4 * i - 1], a1);
        a2 = fmaf(h.z, xs[-4 * i - 2], a2);
        a3 = fmaf(h.w, xs[-4 * i - 3], a3);
    }
    out[m] = (a0 + a1) + (a2 + a3);
}

constexpr int kPhWarps = kThreads / 32;

template <bool I16>
__global__ void __launch_bounds__(kThreads) resample_phase_kernel(const void* __restrict__ in_, long long n_in,
                                                                  float* __restrict__ out_, long long n_out,
                                                                  const float* __restrict__ poly, int up, int down,
                                                                  int K, int half_len, int win, RsBatch bt) {
    extern __shared__ __align__(16) float s_x[];
    const void* in = (const unsigned char*)in_ + (size_t)blockIdx.z * bt.in_stride * (I16 ? 2 : 4);
    float* out = out_ + (size_t)blockIdx.z * bt.out_stride;
    const long long mb = (long long)blockIdx.x * 32 * up;          // first output of this block of 32 * up
    long long n_real = n_out;
    if (bt.in_len) {
        n_in = bt.in_len[blockIdx.z];
        n_real = (n_in * up + down - 1) / down;
        if (mb >= n_real) {                                         // whole block past the signal: padding only
            const int j = blockIdx.y * kPhWarps + (threadIdx.x >> 5);
            const long long m = mb + j + (long long)up * (threadIdx.x & 31);
            if (j < up && m < n_out) out[m] = 0.f;
            return;
        }
    }
    const long long k_lo = (mb * down + half_len) / up - (K - 1);   // oldest sample the block touches
    constexpr int kBatch = 8;
#pragma unroll 1
    for (int i0 = threadIdx.x; i0 < win; i0 += kBatch * kThreads) {
        float v[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const long long k = k_lo + i0 + b * kThreads;
            v[b] = 0.f;
            if (i0 + b * kThreads < win && k >= 0 && k < n_in)
                v[b] = I16 ? (float)__ldg((const int16_t*)in + k) * (1.0f / 32768.0f) : __ldg((const float*)in + k);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b)
            if (i0 + b * kThreads < win) s_x[i0 + b * kThreads] = v[b];
    }
    __syncthreads();
    const int j = blockIdx.y * kPhWarps + (threadIdx.x >> 5);      // which of the block's `up` residues
    if (j >= up) return;
    const long long m = mb + j + (long long)up * (threadIdx.x & 31);
    if (m >= n_out) return;
    if (m >= n_real) { out[m] = 0.f; return; }
    const long long u = m * down + half_len;                        // u % up is the same for all 32 lanes
    const float* xs = s_x + (int)(u / up - k_lo);                   // lanes are `down` samples apart
    const float4* h4 = reinterpret_cast<const float4*>(poly + (size_t)(u % up) * K);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int i = 0; i < K / 4; ++i) {
        const float4 h = __ldg(h4 + i);
        a0 = fmaf(h.x, xs[-4 * i], a0);
        a1 = fmaf(h.y, xs[-4 * i - 1], a1);
        a2 = fmaf(h.z, xs[-4 * i - 2], a2);
        a3 = fmaf(h.w, xs[-4 * i - 3], a3);
    }
    out[m] = (a0 + a1) + (a2 + a3);
}

}  // namespace

struct b2a_resampler {
    int device = 0, orig = 0, target = 0;
    b2a::ResamplerDesign d;
    float* d_poly = nullptr;
    void* d_in = nullptr;
    float* d_out = nullptr;
    size_t cap_in = 0, cap_out = 0;          // bytes
    cudaStream_t stream = nullptr;
    std::mutex mu;                           // run_host works in the handle's own buffers and stream: one call at a time
};

namespace {

int rs_fail(int code, const std::string& msg) {
    g_rs_err = msg;
    b2a_internal_set_error(msg.c_str());          // b2a_last_error() sees it too
    return code;
}

#define RS_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return rs_fail(e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,     \
                           std::string(#expr) + ": " + cudaGetErrorString(e__));          \
    } while (0)

// One launch for `n_clips` signals.  Single-signal calls pass n_clips = 1, bt = {0, 0, NULL}, n_out = the
// signal's own length; batched calls pass the row capacity as n_out and per-clip lengths in bt.in_len.
int rs_launch(b2a_resampler* r, const void* d_in, int in_dtype, long long n_in, float* d_out, long long n_out,
              int n_clips, RsBatch bt, cudaStream_t st) {
    if (n_out <= 0 || n_clips <= 0) return B2A_OK;
    if (n_clips > 65535) return rs_fail(B2A_EINVAL, "at most 65535 clips per batched resampler launch");
    const int K = r->d.taps_per_phase;
    const int up = r->d.up, down = r->d.down;
    // phase-aligned kernel: 32 lanes x `down` samples + the filter must fit shared memory
    const long long win_ph = 32LL * down + K + 2;
    if (up >= 8 && (down & 1) && win_ph * (long long)sizeof(float) <= 200 * 1024) {
        const size_t smem = (size_t)win_ph * sizeof(float);
        const dim3 grid((unsigned)((n_out + 32LL * up - 1) / (32LL * up)), (unsigned)((up + kPhWarps - 1) / kPhWarps),
                        (unsigned)n_clips);
        if (in_dtype == B2A_IN_I16) {
            auto k = resample_phase_kernel<true>;
            RS_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, kThreads, smem, st>>>(d_in, n_in, d_out, n_out, r->d_poly, up, down, K, r->d.half_len, (int)win_ph, bt);
        } else {
            auto k = resample_phase_kernel<false>;
            RS_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, kThreads, smem, st>>>(d_in, n_in, d_out, n_out, r->d_poly, up, down, K, r->d.half_len, (int)win_ph, bt);
        }
        RS_TRY(cudaGetLastError());
        return B2A_OK;
    }
    // input samples spanned by 256 consecutive outputs, plus the filter length
    const int win = K + (int)(((long long)(kThreads - 1) * r->d.down + r->d.up - 1) / r->d.up) + 2;
    const size_t smem = (size_t)win * sizeof(float);
    const dim3 grid((unsigned)((n_out + kThreads - 1) / kThreads), 1, (unsigned)n_clips);
    if (in_dtype == B2A_IN_I16) {
        auto k = resample_kernel<true>;
        RS_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kThreads, smem, st>>>(d_in, n_in, d_out, n_out, r->d_poly, r->d.up, r->d.down, K, r->d.half_len, win, bt);
    } else {
        auto k = resample_kernel<false>;
        RS_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kThreads, smem, st>>>(d_in, n_in, d_out, n_out, r->d_poly, r->d.up, r->d.down, K, r->d.half_len, win, bt);
    }
    RS_TRY(cudaGetLastError());
    return B2A_OK;
}

}  // namespace

extern "C" {

const char* b2a_resampler_last_error(void) { return g_rs_err.c_str(); }

int b2a_resampler_create(int32_t orig_sr, int32_t target_sr, int32_t device, b2a_resampler** out) {
    if (!out) return rs_fail(B2A_EINVAL, "out is NULL");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return rs_fail(B2A_ENODEVICE, "no CUDA device: the resampler has no CPU path");
    }
    if (device < 0 || device >= n_dev) return rs_fail(B2A_EINVAL, "device index out of range");
    b2a_resampler* r = new (std::nothrow) b2a_resampler();
    if (!r) return rs_fail(B2A_ENOMEM, "out of host memory");
    const char* err = nullptr;
    if (!b2a::design_resampler(orig_sr, target_sr, &r->d, &err)) {
        delete r;
        return rs_fail(B2A_EINVAL, err ? err : "resampler design failed");
    }
    r->device = device; r->orig = orig_sr; r->target = target_sr;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_poly, r->d.poly.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(r->d_poly, r->d.poly.data(), r->d.poly.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        b2a_resampler_destroy(r);
        return rs_fail(e == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, cudaGetErrorString(e));
    }
    *out = r;
    return B2A_OK;
}

int b2a_resampler_destroy(b2a_resampler* r) {
    if (!r) return B2A_OK;
    cudaSetDevice(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    cudaFree(r->d_poly); cudaFree(r->d_in); cudaFree(r->d_out);
    delete r;
    return B2A_OK;
}

int64_t b2a_resampler_out_len(const b2a_resampler* r, int64_t n_in) {
    if (!r || n_in <= 0) return 0;
    // librosa.resample: int(np.ceil(n * target / orig)), with the ratio in lowest terms
    return (n_in * r->d.up + r->d.down - 1) / r->d.down;
}

int b2a_resampler_geometry(const b2a_resampler* r, int32_t* up, int32_t* down, int32_t* half_len,
                           int32_t* taps_per_phase, float* poly /* [up][taps_per_phase] or NULL */) {
    if (!r) return rs_fail(B2A_EINVAL, "resampler is NULL");
    if (up) *up = r->d.up;
    if (down) *down = r->d.down;
    if (half_len) *half_len = r->d.half_len;
    if (taps_per_phase) *taps_per_phase = r->d.taps_per_phase;
    if (poly) std::copy(r->d.poly.begin(), r->d.poly.end(), poly);
    return B2A_OK;
}

int b2a_resampler_design(int32_t orig_sr, int32_t target_sr, int32_t* up, int32_t* down, int32_t* half_len,
                         int32_t* taps_per_phase, float* poly, int64_t poly_capacity) {
    b2a::ResamplerDesign d;
    const char* err = nullptr;
    if (!b2a::design_resampler(orig_sr, target_sr, &d, &err)) return rs_fail(B2A_EINVAL, err ? err : "design failed");
    if (up) *up = d.up;
    if (down) *down = d.down;
    if (half_len) *half_len = d.half_len;
    if (taps_per_phase) *taps_per_phase = d.taps_per_phase;
    if (poly) {
        if (poly_capacity < (int64_t)d.poly.size()) return rs_fail(B2A_EINVAL, "poly buffer too small");
        std::copy(d.poly.begin(), d.poly.end(), poly);
    }
    return B2A_OK;
}

int b2a_resampler_run_device(b2a_resampler* r, const void* d_in, int32_t in_dtype, int64_t n_in,
                             float* d_out, void* stream) {
    if (!r) return rs_fail(B2A_EINVAL, "resampler is NULL");
    if (n_in < 0) return rs_fail(B2A_EINVAL, "n_in < 0");
    if (in_dtype != B2A_IN_I16 && in_dtype != B2A_IN_F32) return rs_fail(B2A_EINVAL, "unknown input dtype");
    if (n_in == 0) return B2A_OK;
    if (!d_in || !d_out) return rs_fail(B2A_EINVAL, "NULL buffer");
    RS_TRY(cudaSetDevice(r->device));
    return rs_launch(r, d_in, in_dtype, n_in, d_out, b2a_resampler_out_len(r, n_in), 1, RsBatch{0, 0, nullptr}, (cudaStream_t)stream);
}

int b2a_resampler_run_device_batch(b2a_resampler* r, const void* d_in, int32_t in_dtype, int64_t n_clips,
                                   int64_t in_stride, const int32_t* d_in_len, float* d_out, int64_t out_stride,
                                   int64_t out_cap, void* stream) {
    if (!r) return rs_fail(B2A_EINVAL, "resampler is NULL");
    if (n_clips < 0 || in_stride <= 0 || out_cap <= 0 || out_stride < out_cap) return rs_fail(B2A_EINVAL, "bad batch geometry");
    if (in_dtype != B2A_IN_I16 && in_dtype != B2A_IN_F32) return rs_fail(B2A_EINVAL, "unknown input dtype");
    if (n_clips == 0) return B2A_OK;
    if (!d_in || !d_out || !d_in_len) return rs_fail(B2A_EINVAL, "NULL buffer");
    RS_TRY(cudaSetDevice(r->device));
    for (int64_t c0 = 0; c0 < n_clips; c0 += 32768) {         // gridDim.z limit
        const int nb = (int)std::min<int64_t>(32768, n_clips - c0);
        const int rc = rs_launch(r, (const unsigned char*)d_in + (size_t)c0 * in_stride * (in_dtype == B2A_IN_I16 ? 2 : 4),
                                 in_dtype, in_stride, d_out + (size_t)c0 * out_stride, out_cap, nb,
                                 RsBatch{in_stride, out_stride, d_in_len + c0}, (cudaStream_t)stream);
        if (rc != B2A_OK) return rc;
    }
    return B2A_OK;
}

int b2a_resampler_rates(const b2a_resampler* r, int32_t* orig_sr, int32_t* target_sr, int32_t* device) {
    if (!r) return rs_fail(B2A_EINVAL, "resampler is NULL");
    if (orig_sr) *orig_sr = r->orig;
    if (target_sr) *target_sr = r->target;
    if (device) *device = r->device;
    return B2A_OK;
}

int b2a_resampler_run_host(b2a_resampler* r, const void* in, int32_t in_dtype, int64_t n_in, float* out) {
    if (!r) return rs_fail(B2A_EINVAL, "resampler is NULL");
    if (n_in < 0) return rs_fail(B2A_EINVAL, "n_in < 0");
    if (in_dtype != B2A_IN_I16 && in_dtype != B2A_IN_F32) return rs_fail(B2A_EINVAL, "unknown input dtype");
    if (n_in == 0) return B2A_OK;
    if (!in || !out) return rs_fail(B2A_EINVAL, "NULL buffer");
    // Callers decode files on a thread pool (extractors.py) and share one resampler per (orig, target,
    // device): serialise here rather than trusting every caller to.
    std::lock_guard<std::mutex> guard(r->mu);
    RS_TRY(cudaSetDevice(r->device));
    const size_t in_bytes = (size_t)n_in * (in_dtype == B2A_IN_I16 ? 2 : 4);
    const size_t out_bytes = (size_t)b2a_resampler_out_len(r, n_in) * sizeof(float);
    if (in_bytes > r->cap_in) {
        cudaFree(r->d_in); r->d_in = nullptr; r->cap_in = 0;
        RS_TRY(cudaMalloc(&r->d_in, in_bytes));
        r->cap_in = in_bytes;
    }
    if (out_bytes > r->cap_out) {
        cudaFree(r->d_out); r->d_out = nullptr; r->cap_out = 0;
        RS_TRY(cudaMalloc((void**)&r->d_out, out_bytes));
        r->cap_out = out_bytes;
    }
    RS_TRY(cudaMemcpyAsync(r->d_in, in, in_bytes, cudaMemcpyHostToDevice, r->stream));
    const int rc = rs_launch(r, r->d_in, in_dtype, n_in, r->d_out, b2a_resampler_out_len(r, n_in), 1, RsBatch{0, 0, nullptr}, r->stream);
    if (rc != B2A_OK) return rc;
    RS_TRY(cudaMemcpyAsync(out, r->d_out, out_bytes, cudaMemcpyDeviceToHost, r->stream));
    RS_TRY(cudaStreamSynchronize(r->stream));
    return B2A_OK;
}

}  // extern "C"
