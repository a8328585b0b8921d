// augment.cu — Stage-1b waveform augmentation on the device.
//
// Reference: src/preprocessing/augment.py — volume_scale :88-93, gaussian_noise :96-102, time_shift
// :121-126, polarity_inversion :129-132, _apply_augmentations :186-203 (every enabled step in sequence,
// fresh parameters per copy), _preserve_length :206-212, level match + per-file loop of run() :325-375.
// All of these are element-wise float32 arithmetic plus a cyclic index shift, so one kernel applies a
// whole chain per output sample.  The random numbers are drawn by the HOST with the reference's own
// generator (numpy default_rng(seed), consumed in the reference's order: augment.py:325) so that outputs
// are reproducible against the reference bit for bit; the host sends each copy's step list (gain, noise
// amplitude, shift) and its float32 noise rows — white for gaussian_noise, and for pdm_hiss (:135-167) the
// pink, notched, unit-RMS row the host synthesises from its draw with numpy's own FFT (noise synthesis stays
// with the generator; the device mixes).  time_stretch / pitch_shift (librosa phase vocoder) are not built.
//
// Arithmetic is kept operation-for-operation with NumPy's float32 (no FMA contraction):
//   gain      y * float32(gain)
//   noise     clip(y + (float32(noise) * float32(amp)), -1, 1)
//   roll      out[(j + shift) mod n] = y[j]
//   polarity  -y
// Output is float32, or int16 quantised the way soundfile writes PCM_16 (libsndfile f2s_clip_array with
// clipping on: lrintf(x * 32768) saturated) — what the reference's sf.write + librosa.load round trip
// between Stage 1b and Stage 2 does to the samples (soundfile is absent here: that rule is unpinned).
#include "../../include/b2a.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

void b2a_internal_set_error(const char* msg);      // api.cu

namespace {

constexpr int kAugThreads = 256;
constexpr int kAugPerThread = 4;

template <bool IN_I16, bool OUT_I16>
__global__ void __launch_bounds__(kAugThreads) augment_kernel(const void* __restrict__ src, const long long* __restrict__ src_off,
                                                              const int* __restrict__ lengths,
                                                              const long long* __restrict__ out_off,
                                                              const b2a_aug_step* __restrict__ steps, int max_steps,
                                                              const float* __restrict__ noise, void* __restrict__ out) {
    const long long row = blockIdx.y;
    const int n = lengths[row];
    const b2a_aug_step* st = steps + row * max_steps;
    // total cyclic shift of the chain: the sample that ends at index i started at (i - total) mod n
    long long total = 0;
    int ns = 0;
    for (; ns < max_steps && st[ns].op >= 0; ++ns)
        if (st[ns].op == B2A_AUG_ROLL) total += st[ns].shift;
    const long long so = src_off[row], oo = out_off[row];
    const int i0 = (blockIdx.x * kAugThreads + threadIdx.x) * kAugPerThread;
    if (n <= 0) return;
    int start = (int)(((-total) % n + n) % n);
#pragma unroll
    for (int e = 0; e < kAugPerThread; ++e) {
        const int i = i0 + e;
        if (i >= n) return;
        int cur = i + start;
        if (cur >= n) cur -= n;
        float v = IN_I16 ? __int2float_rn((int)((const int16_t*)src)[so + cur]) * (1.0f / 32768.0f)
                         : ((const float*)src)[so + cur];
        for (int k = 0; k < ns; ++k) {
            const b2a_aug_step s = st[k];
            switch (s.op) {
                case B2A_AUG_GAIN: v = __fmul_rn(v, s.a); break;
                case B2A_AUG_NOISE:
                    v = fminf(fmaxf(__fadd_rn(v, __fmul_rn(noise[s.noise_off + cur], s.a)), -1.0f), 1.0f);
                    break;
                case B2A_AUG_ROLL: {
                    const int sh = ((s.shift % n) + n) % n;
                    cur += sh;
                    if (cur >= n) cur -= n;
                    break;
                }
                case B2A_AUG_POLARITY: v = -v; break;
                default: break;
            }
        }
        if (OUT_I16) {
            const float sc = __fmul_rn(v, 32768.0f);
            short q;
            if (sc >= 32767.0f) q = 32767;
            else if (sc <= -32768.0f) q = -32768;
            else q = (short)__float2int_rn(sc);
            ((int16_t*)out)[oo + i] = q;
        } else {
            ((float*)out)[oo + i] = v;
        }
    }
}

int aug_fail(int code, const std::string& msg) {
    b2a_internal_set_error(msg.c_str());
    return code;
}

#define AUG_TRY(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return aug_fail(e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,                 \
                            std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
    } while (0)

}  // namespace

extern "C" {

int b2a_augment_device(const void* d_src, int32_t src_dtype, const int64_t* d_src_off, const int32_t* d_lengths,
                       const int64_t* d_out_off, int64_t n_out, int32_t max_len, const b2a_aug_step* d_steps,
                       int32_t max_steps, const float* d_noise, void* d_out, int32_t out_dtype, void* stream) {
    if (n_out < 0 || max_len < 0 || max_steps < 0) return aug_fail(B2A_EINVAL, "negative size");
    if ((src_dtype != B2A_IN_I16 && src_dtype != B2A_IN_F32) || (out_dtype != B2A_IN_I16 && out_dtype != B2A_IN_F32))
        return aug_fail(B2A_EINVAL, "dtype");
    if (n_out == 0 || max_len == 0) return B2A_OK;
    if (!d_src || !d_src_off || !d_lengths || !d_out_off || !d_out || (max_steps > 0 && !d_steps))
        return aug_fail(B2A_EINVAL, "NULL buffer");
    static_assert(sizeof(long long) == sizeof(int64_t), "offset type");
    cudaStream_t st = (cudaStream_t)stream;
    const int per_block = kAugThreads * kAugPerThread;
    for (int64_t r0 = 0; r0 < n_out; r0 += 65535) {
        const int nb = (int)(n_out - r0 < 65535 ? n_out - r0 : 65535);
        const dim3 grid((unsigned)((max_len + per_block - 1) / per_block), (unsigned)nb);
        const long long* so = (const long long*)d_src_off + r0;
        const long long* oo = (const long long*)d_out_off + r0;
        const int* ln = d_lengths + r0;
        const b2a_aug_step* sp = d_steps + r0 * max_steps;
        if (src_dtype == B2A_IN_I16) {
            if (out_dtype == B2A_IN_I16) augment_kernel<true, true><<<grid, kAugThreads, 0, st>>>(d_src, so, ln, oo, sp, max_steps, d_noise, d_out);
            else augment_kernel<true, false><<<grid, kAugThreads, 0, st>>>(d_src, so, ln, oo, sp, max_steps, d_noise, d_out);
        } else {
            if (out_dtype == B2A_IN_I16) augment_kernel<false, true><<<grid, kAugThreads, 0, st>>>(d_src, so, ln, oo, sp, max_steps, d_noise, d_out);
            else augment_kernel<false, false><<<grid, kAugThreads, 0, st>>>(d_src, so, ln, oo, sp, max_steps, d_noise, d_out);
        }
        AUG_TRY(cudaGetLastError());
    }
    return B2A_OK;
}

int b2a_augment_host(int32_t device, const void* src, int32_t src_dtype, int64_t src_elems, const int64_t* src_off,
                     const int32_t* lengths, const int64_t* out_off, int64_t n_out, const b2a_aug_step* steps,
                     int32_t max_steps, const float* noise, int64_t noise_elems, void* out, int32_t out_dtype,
                     int64_t out_elems) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return aug_fail(B2A_ENODEVICE, "no CUDA device visible: augmentation has no CPU path");
    }
    if (device < 0 || device >= ndev) return aug_fail(B2A_EINVAL, "device index out of range");
    if (n_out < 0 || src_elems < 0 || noise_elems < 0 || out_elems < 0 || max_steps < 0) return aug_fail(B2A_EINVAL, "negative size");
    if ((src_dtype != B2A_IN_I16 && src_dtype != B2A_IN_F32) || (out_dtype != B2A_IN_I16 && out_dtype != B2A_IN_F32))
        return aug_fail(B2A_EINVAL, "dtype");
    if (n_out == 0) return B2A_OK;
    if (!src || !src_off || !lengths || !out_off || !out || (max_steps > 0 && !steps)) return aug_fail(B2A_EINVAL, "NULL buffer");
    int32_t max_len = 0;
    for (int64_t r = 0; r < n_out; ++r) {
        const int64_t L = lengths[r];
        if (L < 0 || src_off[r] < 0 || src_off[r] + L > src_elems) return aug_fail(B2A_EINVAL, "source range out of bounds");
        if (out_off[r] < 0 || out_off[r] + L > out_elems) return aug_fail(B2A_EINVAL, "output range out of bounds");
        for (int k = 0; k < max_steps; ++k) {
            const b2a_aug_step& s = steps[r * max_steps + k];
            if (s.op < 0) break;
            if (s.op > B2A_AUG_POLARITY) return aug_fail(B2A_EINVAL, "unknown augmentation op");
            if (s.op == B2A_AUG_NOISE && (!noise || s.noise_off < 0 || s.noise_off + L > noise_elems))
                return aug_fail(B2A_EINVAL, "noise range out of bounds");
        }
        if (L > max_len) max_len = (int32_t)L;
    }
    AUG_TRY(cudaSetDevice(device));
    const size_t se = src_dtype == B2A_IN_I16 ? 2 : 4, oe = out_dtype == B2A_IN_I16 ? 2 : 4;
    void *d_src = nullptr, *d_out = nullptr, *d_meta = nullptr, *d_steps = nullptr;
    float* d_noise = nullptr;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_src); cudaFree(d_out); cudaFree(d_meta); cudaFree(d_steps); cudaFree(d_noise);
        if (st) cudaStreamDestroy(st);
    };
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    const size_t meta_bytes = (size_t)n_out * (8 + 8 + 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_src, (size_t)src_elems * se + 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)out_elems * oe + 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_meta, meta_bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_steps, (size_t)n_out * (max_steps > 0 ? max_steps : 1) * sizeof(b2a_aug_step));
    if (e == cudaSuccess && noise_elems > 0) e = cudaMalloc((void**)&d_noise, (size_t)noise_elems * 4);
    long long* d_so = (long long*)d_meta;
    long long* d_oo = d_so ? d_so + n_out : nullptr;
    int* d_len = d_oo ? (int*)(d_oo + n_out) : nullptr;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_src, src, (size_t)src_elems * se, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_so, src_off, (size_t)n_out * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_oo, out_off, (size_t)n_out * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_len, lengths, (size_t)n_out * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && max_steps > 0)
        e = cudaMemcpyAsync(d_steps, steps, (size_t)n_out * max_steps * sizeof(b2a_aug_step), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && noise_elems > 0) e = cudaMemcpyAsync(d_noise, noise, (size_t)noise_elems * 4, cudaMemcpyHostToDevice, st);
    int rc = B2A_OK;
    if (e == cudaSuccess)
        rc = b2a_augment_device(d_src, src_dtype, (const int64_t*)d_so, d_len, (const int64_t*)d_oo, n_out, max_len,
                                (const b2a_aug_step*)d_steps, max_steps, d_noise, d_out, out_dtype, st);
    if (e == cudaSuccess && rc == B2A_OK) e = cudaMemcpyAsync(out, d_out, (size_t)out_elems * oe, cudaMemcpyDeviceToHost, st);
    const cudaError_t es = st ? cudaStreamSynchronize(st) : cudaSuccess;
    cleanup();
    if (rc != B2A_OK) return rc;
    if (e != cudaSuccess) return aug_fail(e == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, std::string("augment: ") + cudaGetErrorString(e));
    if (es != cudaSuccess) return aug_fail(B2A_ECUDA, std::string("augment: ") + cudaGetErrorString(es));
    return B2A_OK;
}

}  // extern "C"
