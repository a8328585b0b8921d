// effects.cu — the two librosa-backed augmentors of Stage 1b on the device (SURVEY 8f N3 "later" part;
// reference: src/preprocessing/augment.py:105-118 -> librosa.effects.time_stretch / pitch_shift).  sm_100a only.
//
// time_stretch(y, rate) = istft(phase_vocoder(stft(y, n_fft 2048, hop 512), rate), length = round(n / rate)):
//   stft2048_kernel     complex64 STFT, zero-padded centre frames, periodic Hann in double (fft_core.cuh)
//   phase_vocoder_kernel  one thread per bin walks the output steps: float32 phase accumulator, float64 phase
//                       increments wrapped to (-pi, pi], linear interpolation of |D| between the two nearest frames,
//                       exactly the dtype sequence of librosa.phase_vocoder
//   istft_frames_kernel inverse real FFT of every output frame: the split step run backwards, then the forward
//                       complex FFT on the conjugate (irfft(X) = conj(fft(conj Z)) / NC, two samples per point)
//   overlap_add_kernel  window x frame, frames added in ascending order into a float32 accumulator, divided by the
//                       window's sum of squares accumulated the same way (librosa.istft / window_sumsquare)
// pitch_shift(y, sr, n_steps) = fix_length(resample(time_stretch(y, 2^(-n/12)), ratio), len(y)):
//   resample_arbitrary_kernel  the project's resampler specification (tables.cpp: design_resampler) as a continuous
//                       kernel, evaluated at the exact output instants m / ratio from a finely sampled table
//                       (1/2048 sample) with linear interpolation — the ratio 2^(n/12) is irrational, libsoxr's
//                       variable-rate path is what it stands in for (parity unpinned).
// Rows are processed in chunks sized to ~1 GB of spectra; there is no CPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b2a.h"
#include "fft_core.cuh"
#include "front_stage.cuh"
#include "tables.h"

void b2a_internal_set_error(const char* msg);      // api.cu

namespace b2a {
namespace {

constexpr int kLog2NC = 10;
using G = FftGeom<kLog2NC>;
constexpr int NC = G::NC, NFFT = G::NFFT, T = G::T, NB = NC + 1, kHop = 512;
constexpr int kF = 16;                          // frames per CTA
constexpr int kFR = kThreads / T;               // frames per FFT round (4)
constexpr int kResTab = 2048;                   // kernel table samples per unit of the lower rate
constexpr double kResHalf = 95.5;
constexpr double kTwoPi = 6.283185307179586476925286766559;

struct RowInfo {
    long long in_off;       // element offset of the row's samples
    int n;                  // samples
    int frames;             // 1 + n / hop
    int steps;              // phase-vocoder output frames
    int n_frames;           // frames the inverse transform uses
    int out_len;            // round(n / rate)
    double rate;
};

struct FftSmem {
    float* audio; float2* xch; float2* tw; float2* tw2; float2* twp;
};
__device__ __forceinline__ FftSmem carve_fft(unsigned char* smem, int cl) {
    FftSmem s;
    s.audio = reinterpret_cast<float*>(smem);
    s.xch = reinterpret_cast<float2*>(s.audio + cl);
    s.tw = s.xch + kFR * G::XSTRIDE;
    s.tw2 = s.tw + NC;
    s.twp = s.tw2 + NC / 2 + 1;
    return s;
}
constexpr int kClStft = (kHop * (kF - 1) + NFFT + 7) & ~7;
size_t fft_smem_bytes(int cl) {
    return (size_t)cl * 4 + (size_t)kFR * G::XSTRIDE * 8 + (size_t)NC * 8 + (size_t)(NC / 2 + 1) * 8 +
           (size_t)FftTwp<kLog2NC>::SIZE * 8 + 64;
}
__device__ __forceinline__ void fill_tw(const FftSmem& s, const float2* tw, const float2* tw2) {
    for (int i = threadIdx.x; i < NC; i += kThreads) s.tw[i] = tw[i];
    for (int i = threadIdx.x; i < NC / 2 + 1; i += kThreads) s.tw2[i] = tw2[i];
    FftTwp<kLog2NC>::fill(s.twp, tw, threadIdx.x, kThreads);
}

// ---- complex STFT: spec[(row * frames_stride + t) * NB + k] ---------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) stft2048_kernel(const float* __restrict__ src, const RowInfo* __restrict__ rows,
                                                               const float* __restrict__ window, const float2* __restrict__ tw,
                                                               const float2* __restrict__ tw2, float2* __restrict__ spec,
                                                               int frames_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem s = carve_fft(smem, kClStft);
    const RowInfo r = rows[blockIdx.y];
    const int t0 = blockIdx.x * kF;
    if (t0 >= r.frames) return;
    fill_tw(s, tw, tw2);
    const int j = threadIdx.x % T, slot = threadIdx.x / T;
    const float2* const win = reinterpret_cast<const float2*>(window) + j;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    stage_audio<false>(s.audio, src + r.in_off, r.in_off, t0 * kHop - NFFT / 2, kClStft, r.n, 0, aligned);
    __syncthreads();
#pragma unroll 1
    for (int rd = 0; rd < kF / kFR; ++rd) {
        const int f = rd * kFR + slot;
        float2* xb = s.xch + slot * G::XSTRIDE;
        {
            float2 v[16];
            const float* a = s.audio + f * kHop + 2 * j;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const float2 x = *reinterpret_cast<const float2*>(a + 2 * T * t);
                const float2 w = __ldg(win + T * t);
                v[t] = make_float2(x.x * w.x, x.y * w.y);
            }
            Dft<16>::run(v);
#pragma unroll
            for (int t = 0; t < 16; ++t) xb[xpad(16 * j + t)] = v[t];
        }
        frame_sync<T>();
        fft_tail_passes<kLog2NC, false>(xb, s.tw, nullptr, s.twp, j);
        if (t0 + f < r.frames) {
            float2* o = spec + ((size_t)blockIdx.y * frames_stride + t0 + f) * NB;
#pragma unroll
            for (int r2 = 0; r2 < 8; ++r2) {
                const int k = j + T * r2;
                const float2 A = xb[xpad(k)];
                const float2 B = xb[xpad((NC - k) & (NC - 1))];
                float2 xk, xnk;
                rfft_split(A, B, s.tw2[k], xk, xnk);
                o[k] = make_float2(0.5f * xk.x, 0.5f * xk.y);
                o[NC - k] = make_float2(0.5f * xnk.x, 0.5f * xnk.y);
            }
            if (j == 0) {
                const float2 A = xb[xpad(NC / 2)];
                o[NC / 2] = make_float2(A.x, -A.y);
            }
        }
        frame_sync<T>();
    }
}

// ---- phase vocoder: one thread per bin, sequential over the output steps ------------------------------------------
__global__ void __launch_bounds__(kThreads) phase_vocoder_kernel(const float2* __restrict__ spec, const RowInfo* __restrict__ rows,
                                                                 float2* __restrict__ out, int frames_stride, int steps_stride) {
    const RowInfo r = rows[blockIdx.x];
    const float2* D = spec + (size_t)blockIdx.x * frames_stride * NB;
    float2* O = out + (size_t)blockIdx.x * steps_stride * NB;
    const double val = 1.0 / ((double)NFFT * (1.0 / kTwoPi));                 // np.fft.rfftfreq(n_fft, 1 / (2 pi))
    for (int k = threadIdx.x; k < NB; k += kThreads) {
        const double phi = (double)kHop * ((double)k * val);
        const float2 c0 = D[k];
        float acc = atan2f(c0.y, c0.x);
        for (int t = 0; t < r.n_frames; ++t) {               // (later steps never reach the inverse transform)
            const double step = (double)t * r.rate;
            const int i0 = (int)step;
            const double alpha = step - floor(step);
            const float2 a = i0 < r.frames ? D[(size_t)i0 * NB + k] : make_float2(0.f, 0.f);
            const float2 b = i0 + 1 < r.frames ? D[(size_t)(i0 + 1) * NB + k] : make_float2(0.f, 0.f);
            const double mag = (1.0 - alpha) * (double)hypotf(a.x, a.y) + alpha * (double)hypotf(b.x, b.y);
            float sn, cs;
            sincosf(acc, &sn, &cs);
            O[(size_t)t * NB + k] = make_float2((float)((double)cs * mag), (float)((double)sn * mag));
            double dphase = (double)__fsub_rn(atan2f(b.y, b.x), atan2f(a.y, a.x)) - phi;
            dphase = dphase - kTwoPi * rint(dphase / kTwoPi);
            acc = (float)((double)acc + (phi + dphase));
        }
    }
}

// ---- inverse real FFT of every frame: fr[(row * steps_stride + t) * NFFT + i] ---------------------------------------
__global__ void __launch_bounds__(kThreads, 2) istft_frames_kernel(const float2* __restrict__ spec, const RowInfo* __restrict__ rows,
                                                                   const float2* __restrict__ tw, const float2* __restrict__ tw2,
                                                                   float* __restrict__ fr, int steps_stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const FftSmem s = carve_fft(smem, 0);
    const RowInfo r = rows[blockIdx.y];
    const int t0 = blockIdx.x * kF;
    if (t0 >= r.n_frames) return;
    fill_tw(s, tw, tw2);
    __syncthreads();
    const int j = threadIdx.x % T, slot = threadIdx.x / T;
#pragma unroll 1
    for (int rd = 0; rd < kF / kFR; ++rd) {
        const int t = t0 + rd * kFR + slot;
        const bool live = t < r.n_frames;
        float2* xb = s.xch + slot * G::XSTRIDE;
        const float2* X = spec + ((size_t)blockIdx.y * steps_stride + (live ? t : 0)) * NB;
        // conj(Z[k]) and conj(Z[NC - k]) from X[k], X[NC - k]:  E = (X[k] + conj X[NC-k]) / 2,
        // O = conj(w_k) (X[k] - conj X[NC-k]) / 2, Z[k] = E + i O, Z[NC-k] = conj(E) + i conj(O)
#pragma unroll
        for (int r2 = 0; r2 < 8; ++r2) {
            const int k = j + T * r2;
            float2 a = X[k], b = X[NC - k];
            if (k == 0) { a.y = 0.f; b.y = 0.f; }                            // c2r ignores the imaginary parts of DC / Nyquist
            const float2 w = s.tw2[k];                                        // exp(-i pi k / NC)
            const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
            const float2 Dm = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));
            const float2 O = make_float2(w.x * Dm.x + w.y * Dm.y, w.x * Dm.y - w.y * Dm.x);     // conj(w) * Dm
            // Z[k] = (E.x - O.y, E.y + O.x); Z[NC-k] = (E.x + O.y, -E.y + O.x); store the conjugates
            xb[xpad(k)] = make_float2(E.x - O.y, -(E.y + O.x));
            if (k != 0) xb[xpad(NC - k)] = make_float2(E.x + O.y, E.y - O.x);
        }
        if (j == 0) {
            const float2 a = X[NC / 2];
            xb[xpad(NC / 2)] = make_float2(a.x, a.y);                        // Z[NC/2] = conj(X[NC/2]); its conjugate is X itself
        }
        frame_sync<T>();
        {
            float2 v[16];
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) v[tt] = xb[xpad(j + T * tt)];
            frame_sync<T>();
            Dft<16>::run(v);
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) xb[xpad(16 * j + tt)] = v[tt];
        }
        frame_sync<T>();
        fft_tail_passes<kLog2NC, false>(xb, s.tw, nullptr, s.twp, j);
        if (live) {
            float2* o = reinterpret_cast<float2*>(fr + ((size_t)blockIdx.y * steps_stride + t) * NFFT);
            const float sc = 1.0f / NC;
#pragma unroll
            for (int tt = 0; tt < 16; ++tt) {
                const int n = j + T * tt;
                const float2 F = xb[xpad(n)];
                o[n] = make_float2(F.x * sc, -F.y * sc);                      // x[2n] + i x[2n+1] = conj(F[n]) / NC
            }
        }
        frame_sync<T>();
    }
}

// ---- overlap-add + window sum-of-squares normalisation ----------------------------------------------------------
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ fr, const RowInfo* __restrict__ rows,
                                                          const double* __restrict__ win, float* __restrict__ out,
                                                          const long long* __restrict__ out_off, int steps_stride) {
    const RowInfo r = rows[blockIdx.y];
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m >= r.out_len) return;
    const int p = m + NFFT / 2;
    const int t_hi = min(p / kHop, r.n_frames - 1);
    const int t_lo = max(0, (p - NFFT + kHop) / kHop);
    const float* F = fr + (size_t)blockIdx.y * steps_stride * NFFT;
    float acc = 0.f, ws = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) {
        const int i = p - kHop * t;
        if (i < 0 || i >= NFFT) continue;
        const double w = win[i];
        acc = (float)__dadd_rn((double)acc, __dmul_rn(w, (double)F[(size_t)t * NFFT + i]));
        ws = (float)__dadd_rn((double)ws, __dmul_rn(w, w));
    }
    out[out_off[blockIdx.y] + m] = ws > 1.17549435e-38f ? __fdiv_rn(acc, ws) : acc;
}

// ---- arbitrary-ratio resampler (pitch_shift): out[m] = sum_j y[j] s g1(s (m / ratio - j)), zero past n_target -----------
__global__ void __launch_bounds__(256) resample_arbitrary_kernel(const float* __restrict__ src, const long long* __restrict__ src_off,
                                                                 const int* __restrict__ src_len, const double* __restrict__ ratios,
                                                                 const float* __restrict__ tab, float* __restrict__ out,
                                                                 const long long* __restrict__ out_off, const int* __restrict__ out_len) {
    const int row = blockIdx.y;
    const int m = blockIdx.x * 256 + threadIdx.x;
    const int n_out = out_len[row];
    if (m >= n_out) return;
    const int n = src_len[row];
    const double ratio = ratios[row];
    const int n_target = (int)ceil((double)n * ratio);
    float* o = out + out_off[row];
    if (m >= n_target) { o[m] = 0.f; return; }
    const double s = ratio < 1.0 ? ratio : 1.0;
    const double tau = (double)m / ratio;
    const double reach = kResHalf / s;
    int j0 = (int)ceil(tau - reach), j1 = (int)floor(tau + reach);
    j0 = max(j0, 0); j1 = min(j1, n - 1);
    const float* y = src + src_off[row];
    double acc = 0.0;
    for (int jj = j0; jj <= j1; ++jj) {
        const double u = fabs(s * (tau - (double)jj)) * kResTab;
        const int i0 = (int)u;
        const float fr = (float)(u - (double)i0);
        const float g0 = __ldg(tab + i0), g1 = __ldg(tab + i0 + 1);
        acc += (double)(y[jj] * (g0 + fr * (g1 - g0)));
    }
    o[m] = (float)(acc * s);
}

int fx_fail(int code, const std::string& msg) {
    b2a_internal_set_error(msg.c_str());
    return code;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { cudaFree(p); p = nullptr; return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename U> U* as() const { return (U*)p; }
};

#define FX_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fx_fail(e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA,                  \
                           std::string(#expr) + ": " + cudaGetErrorString(e__));                       \
    } while (0)

std::vector<float> resample_table() {
    // g1(u), u = i / kResTab in [0, 95.5] (+ 2 guard entries), unit area: same closed form as design_resampler
    const double atten = 125.0, pass = 0.913, stop = 1.0;
    const double beta = 0.1102 * (atten - 8.7), fc = 0.5 * (pass + stop) * 0.5, pi = 3.14159265358979323846;
    auto i0f = [](double x) { double s = 1, t = 1; const double q = x * x / 4; for (int k = 1; k < 200; ++k) { t *= q / ((double)k * k); s += t; if (t < 1e-18 * s) break; } return s; };
    const double i0b = i0f(beta);
    auto g = [&](double u) {
        const double r = u / kResHalf;
        if (r > 1.0) return 0.0;
        const double x = 2.0 * fc * u;
        const double sinc = x == 0.0 ? 1.0 : std::sin(pi * x) / (pi * x);
        return 2.0 * fc * sinc * i0f(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    };
    const int n = (int)(kResHalf * kResTab) + 3;
    // area by the trapezoid rule on a 1/1000 grid (the oracle's rule: effects_restated._kernel_area)
    double area = 0.0;
    const int na = 2 * 95500;
    for (int i = 0; i <= na; ++i) {
        const double u = -kResHalf + (2.0 * kResHalf) * i / na;
        area += ((i == 0 || i == na) ? 0.5 : 1.0) * g(std::fabs(u));
    }
    area *= (2.0 * kResHalf) / na;
    std::vector<float> tab(n);
    for (int i = 0; i < n; ++i) tab[i] = (float)(g((double)i / kResTab) / area);
    return tab;
}

struct FxTables {
    DevBuf window, tw, tw2, win64, restab;
    int device = -1;
};

// one set of constant tables per device, built on first use
int get_tables(int device, FxTables** out) {
    static FxTables tabs[64];
    static std::mutex mu;
    if (device < 0 || device >= 64) return fx_fail(B2A_EINVAL, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    FxTables& t = tabs[device];
    if (t.device != device) {
        std::vector<float> window = hann_periodic(NFFT);
        std::vector<float> tw = twiddles(NC, NC), tw2 = twiddles(2 * NC, NC / 2 + 1);
        std::vector<double> w64(NFFT);
        for (int i = 0; i < NFFT; ++i) w64[i] = 0.5 - 0.5 * std::cos(2.0 * 3.14159265358979323846 * i / NFFT);
        std::vector<float> rt = resample_table();
        FX_TRY(t.window.alloc(window.size() * 4)); FX_TRY(cudaMemcpy(t.window.p, window.data(), window.size() * 4, cudaMemcpyHostToDevice));
        FX_TRY(t.tw.alloc(tw.size() * 4)); FX_TRY(cudaMemcpy(t.tw.p, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
        FX_TRY(t.tw2.alloc(tw2.size() * 4)); FX_TRY(cudaMemcpy(t.tw2.p, tw2.data(), tw2.size() * 4, cudaMemcpyHostToDevice));
        FX_TRY(t.win64.alloc(w64.size() * 8)); FX_TRY(cudaMemcpy(t.win64.p, w64.data(), w64.size() * 8, cudaMemcpyHostToDevice));
        FX_TRY(t.restab.alloc(rt.size() * 4)); FX_TRY(cudaMemcpy(t.restab.p, rt.data(), rt.size() * 4, cudaMemcpyHostToDevice));
        t.device = device;
    }
    *out = &t;
    return B2A_OK;
}

// time_stretch of rows [a, b) of a batch whose samples are already on the device; result into d_out (device)
int stretch_chunk(const FxTables& tb, const float* d_src, const std::vector<RowInfo>& rows, size_t a, size_t b,
                  float* d_out, const long long* d_out_off_chunk, cudaStream_t st) {
    const int nr = (int)(b - a);
    int fmax = 1, smax = 1, omax = 1;
    for (size_t i = a; i < b; ++i) {
        fmax = std::max(fmax, rows[i].frames); smax = std::max(smax, rows[i].n_frames); omax = std::max(omax, rows[i].out_len);
    }
    DevBuf d_rows, spec, spec2, fr;
    FX_TRY(d_rows.alloc((size_t)nr * sizeof(RowInfo)));
    FX_TRY(cudaMemcpyAsync(d_rows.p, rows.data() + a, (size_t)nr * sizeof(RowInfo), cudaMemcpyHostToDevice, st));
    FX_TRY(spec.alloc((size_t)nr * fmax * NB * 8));
    FX_TRY(spec2.alloc((size_t)nr * smax * NB * 8));
    FX_TRY(fr.alloc((size_t)nr * smax * NFFT * 4));
    const size_t sm1 = fft_smem_bytes(kClStft), sm3 = fft_smem_bytes(0);
    FX_TRY(cudaFuncSetAttribute(stft2048_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    FX_TRY(cudaFuncSetAttribute(istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
    stft2048_kernel<<<dim3((fmax + kF - 1) / kF, nr), kThreads, sm1, st>>>(d_src, d_rows.as<RowInfo>(), tb.window.as<float>(),
                                                                         tb.tw.as<float2>(), tb.tw2.as<float2>(), spec.as<float2>(), fmax);
    FX_TRY(cudaGetLastError());
    phase_vocoder_kernel<<<nr, kThreads, 0, st>>>(spec.as<float2>(), d_rows.as<RowInfo>(), spec2.as<float2>(), fmax, smax);
    FX_TRY(cudaGetLastError());
    istft_frames_kernel<<<dim3((smax + kF - 1) / kF, nr), kThreads, sm3, st>>>(spec2.as<float2>(), d_rows.as<RowInfo>(), tb.tw.as<float2>(),
                                                                             tb.tw2.as<float2>(), fr.as<float>(), smax);
    FX_TRY(cudaGetLastError());
    overlap_add_kernel<<<dim3((omax + 255) / 256, nr), 256, 0, st>>>(fr.as<float>(), d_rows.as<RowInfo>(), tb.win64.as<double>(), d_out,
                                                                    d_out_off_chunk, smax);
    FX_TRY(cudaGetLastError());
    FX_TRY(cudaStreamSynchronize(st));
    return B2A_OK;
}

}  // namespace
}  // namespace b2a

extern "C" {

int b2a_time_stretch_host(int32_t device, const float* src, int64_t src_elems, const int64_t* src_off, const int32_t* lengths,
                          const double* rates, int64_t n_rows, float* out, int64_t out_elems, const int64_t* out_off,
                          const int32_t* out_len) {
    using namespace b2a;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fx_fail(B2A_ENODEVICE, "no CUDA device visible: time_stretch has no CPU path");
    }
    if (device < 0 || device >= ndev) return fx_fail(B2A_EINVAL, "device index out of range");
    if (n_rows < 0 || src_elems < 0 || out_elems < 0) return fx_fail(B2A_EINVAL, "negative size");
    if (n_rows == 0) return B2A_OK;
    if (!src || !src_off || !lengths || !rates || !out || !out_off || !out_len) return fx_fail(B2A_EINVAL, "NULL buffer");
    std::vector<RowInfo> rows((size_t)n_rows);
    for (int64_t i = 0; i < n_rows; ++i) {
        RowInfo& r = rows[(size_t)i];
        if (!(rates[i] > 0.0)) return fx_fail(B2A_EINVAL, "rate must be a positive number");        // librosa's own check
        if (lengths[i] < 1 || src_off[i] < 0 || src_off[i] + lengths[i] > src_elems) return fx_fail(B2A_EINVAL, "input row out of range");
        if (out_len[i] < 1 || out_off[i] < 0 || out_off[i] + out_len[i] > out_elems) return fx_fail(B2A_EINVAL, "output row out of range");
        r.in_off = src_off[i]; r.n = lengths[i]; r.rate = rates[i];
        r.frames = 1 + r.n / kHop;
        r.steps = (int)std::ceil((double)r.frames / r.rate);                    // len(np.arange(0, frames, rate))
        r.out_len = out_len[i];
        const int padded = r.out_len + NFFT;
        r.n_frames = std::min(r.steps, (padded + kHop - 1) / kHop);
    }
    FX_TRY(cudaSetDevice(device));
    FxTables* tb = nullptr;
    if (int rc = get_tables(device, &tb)) return rc;
    cudaStream_t st = nullptr;
    FX_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    DevBuf d_src, d_out, d_ooff;
    int rc = B2A_OK;
    do {
        cudaError_t e;
        if ((e = d_src.alloc((size_t)src_elems * 4)) != cudaSuccess || (e = d_out.alloc((size_t)out_elems * 4)) != cudaSuccess ||
            (e = d_ooff.alloc((size_t)n_rows * 8)) != cudaSuccess) {
            rc = fx_fail(e == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA, std::string("time_stretch buffers: ") + cudaGetErrorString(e));
            break;
        }
        if ((e = cudaMemcpyAsync(d_src.p, src, (size_t)src_elems * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_ooff.p, out_off, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
            rc = fx_fail(B2A_ECUDA, std::string("time_stretch H2D: ") + cudaGetErrorString(e));
            break;
        }
        // chunks of rows: ~1 GB of spectra + frames at a time
        size_t a = 0;
        while (a < rows.size() && rc == B2A_OK) {
            size_t b = a, bytes = 0;
            while (b < rows.size()) {
                const size_t per = (size_t)rows[b].frames * NB * 8 + (size_t)rows[b].n_frames * (NB * 8 + NFFT * 4);
                if (b > a && (bytes + per > ((size_t)1 << 30) || b - a >= 32768)) break;
                bytes += per; ++b;
            }
            rc = stretch_chunk(*tb, d_src.as<float>(), rows, a, b, d_out.as<float>(), d_ooff.as<long long>() + a, st);
            a = b;
        }
        if (rc != B2A_OK) break;
        if ((e = cudaMemcpyAsync(out, d_out.p, (size_t)out_elems * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
            (e = cudaStreamSynchronize(st)) != cudaSuccess)
            rc = fx_fail(B2A_ECUDA, std::string("time_stretch D2H: ") + cudaGetErrorString(e));
    } while (0);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

int b2a_pitch_shift_host(int32_t device, const float* src, int64_t src_elems, const int64_t* src_off, const int32_t* lengths,
                         const double* rates, const double* ratios, const int32_t* mid_len, int64_t n_rows, float* out,
                         int64_t out_elems, const int64_t* out_off) {
    using namespace b2a;
    if (n_rows < 0) return fx_fail(B2A_EINVAL, "negative size");
    if (n_rows == 0) return B2A_OK;
    if (!ratios || !mid_len || !lengths || !out_off || !out) return fx_fail(B2A_EINVAL, "NULL buffer");
    // 1. time_stretch into a packed host-side intermediate (rows of mid_len[i] samples)
    std::vector<int64_t> moff((size_t)n_rows);
    int64_t mtot = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        if (mid_len[i] < 1) return fx_fail(B2A_EINVAL, "stretched length must be positive");
        if (!(ratios[i] > 0.0)) return fx_fail(B2A_EINVAL, "ratio must be positive");
        if (out_off[i] < 0 || out_off[i] + lengths[i] > out_elems) return fx_fail(B2A_EINVAL, "output row out of range");
        moff[(size_t)i] = mtot; mtot += mid_len[i];
    }
    std::vector<float> mid((size_t)mtot);
    if (int rc = b2a_time_stretch_host(device, src, src_elems, src_off, lengths, rates, n_rows, mid.data(), mtot, moff.data(), mid_len))
        return rc;
    // 2. resample at the exact output instants, 3. crop / zero-pad to the input length
    FX_TRY(cudaSetDevice(device));
    FxTables* tb = nullptr;
    if (int rc = get_tables(device, &tb)) return rc;
    DevBuf d_mid, d_moff, d_mlen, d_rat, d_out, d_ooff, d_olen;
    FX_TRY(d_mid.alloc((size_t)mtot * 4)); FX_TRY(d_moff.alloc((size_t)n_rows * 8)); FX_TRY(d_mlen.alloc((size_t)n_rows * 4));
    FX_TRY(d_rat.alloc((size_t)n_rows * 8)); FX_TRY(d_out.alloc((size_t)out_elems * 4)); FX_TRY(d_ooff.alloc((size_t)n_rows * 8));
    FX_TRY(d_olen.alloc((size_t)n_rows * 4));
    FX_TRY(cudaMemcpy(d_mid.p, mid.data(), (size_t)mtot * 4, cudaMemcpyHostToDevice));
    FX_TRY(cudaMemcpy(d_moff.p, moff.data(), (size_t)n_rows * 8, cudaMemcpyHostToDevice));
    FX_TRY(cudaMemcpy(d_mlen.p, mid_len, (size_t)n_rows * 4, cudaMemcpyHostToDevice));
    FX_TRY(cudaMemcpy(d_rat.p, ratios, (size_t)n_rows * 8, cudaMemcpyHostToDevice));
    FX_TRY(cudaMemcpy(d_ooff.p, out_off, (size_t)n_rows * 8, cudaMemcpyHostToDevice));
    FX_TRY(cudaMemcpy(d_olen.p, lengths, (size_t)n_rows * 4, cudaMemcpyHostToDevice));
    int omax = 1;
    for (int64_t i = 0; i < n_rows; ++i) omax = std::max(omax, lengths[i]);
    for (int64_t a = 0; a < n_rows; a += 32768) {
        const int nr = (int)std::min<int64_t>(32768, n_rows - a);
        resample_arbitrary_kernel<<<dim3((omax + 255) / 256, nr), 256>>>(d_mid.as<float>(), d_moff.as<long long>() + a, d_mlen.as<int>() + a,
                                                                         d_rat.as<double>() + a, tb->restab.as<float>(), d_out.as<float>(),
                                                                         d_ooff.as<long long>() + a, d_olen.as<int>() + a);
        FX_TRY(cudaGetLastError());
    }
    FX_TRY(cudaMemcpy(out, d_out.p, (size_t)out_elems * 4, cudaMemcpyDeviceToHost));
    return B2A_OK;
}

}  // extern "C"
