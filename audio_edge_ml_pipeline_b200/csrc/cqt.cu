// cqt.cu — audio_cqt on the GPU: 2:1 decimation cascade, per-octave rectangular-window STFT
// times the sparsified wavelet basis, amplitude_to_db(ref=max) + min-max.  sm_100a only.
//
// Follows librosa.cqt == vqt(gamma=0) as the reference calls it (deep.py:249-260): octaves are
// processed top-down; each octave's response is  basis_o . STFT(y_o; n_fft_o, hop_o, ones)  and
// y_{o+1} = sqrt(2) * decimate2(y_o).  The decimator is this project's stand-in for soxr_hq
// (tables.h: decimator_taps; DESIGN.md "CQT decimator").
//
// The basis product is a 12 x 129 complex matrix with ~147 non-zeros against [129 x frames]: as a
// dense GEMM it would be 10x the flops of the banded form and needs fp32-grade accuracy 80 dB
// below the peak, so it runs on the CUDA cores as a banded complex dot product, not on tcgen05.
#include "cqt.h"
#include "fft_core.cuh"
#include "gen/decim_taps.inc"   // build-time generated decimator taps (gen_mel.cpp decim)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <type_traits>

namespace b2a {

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// 2:1 decimator:  out[m] = sqrt(2) * sum_j h[j] y[2m + 191 - j],  j = 0..382, zero-extended.
// Polyphase: A[q] = y[2q+191] meets the even taps, B[q] = y[2q+190] the odd taps.  (B[q], A[q]) are
// neighbours in the input, so the staged tile is a plain copy of the signal as float2 pairs and
// ONE packed FFMA2 advances both polyphase sums of an output: acc2 += (B, A) * (h_odd, h_even).
// ------------------------------------------------------------------------------------------
constexpr int kDecThreads = 128;
constexpr int kDecR = 17;                       // consecutive outputs per thread; odd, so the 64-bit window
                                                // loads of a half-warp (pair stride 17) hit 16 distinct bank
                                                // pairs with no padding and constant offsets in the loop
constexpr int kDecTile = kDecThreads * kDecR;   // outputs per CTA
#ifndef B2A_DEC_U
#define B2A_DEC_U 8
#endif
constexpr int kDecU = B2A_DEC_U;                // tap pairs per register-window step (8 or 11)
constexpr int kDecHalf = 192;                   // tap pairs (the odd phase is zero-padded)
#ifndef B2A_AB_DECMID
#define B2A_AB_DECMID 16
#endif
constexpr int kDecMid = B2A_AB_DECMID;          // centre tap pairs accumulated in fp64: pairs [88, 104)
constexpr int kDecOuter = (kDecHalf - kDecMid) / 2;   // 88 pairs on either side
constexpr int kDecLocal = kDecTile + kDecHalf;  // staged sample pairs
static_assert(kDecOuter % kDecU == 0, "outer taps must split into whole window steps");

template <bool I16>
__global__ void __launch_bounds__(kDecThreads, 4) cqt_decimate_kernel(
    const void* __restrict__ in, size_t in_stride, int in_len, float* __restrict__ out,
    size_t out_stride, int out_len, const float* __restrict__ taps) {
    __shared__ __align__(16) float2 s2[kDecLocal];      // s2[l] = (y[nstart + 2l], y[nstart + 2l + 1])
    static_assert(kDecimTapsGen == kDecimTaps, "regenerate gen/decim_taps.inc");
    (void)taps;                                   // the taps are compile-time constants now
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * kDecTile;
    const size_t clip = blockIdx.y;
    // contiguous input range feeding this tile: n = nstart + 2l (+1).  nstart is even, so one 32-bit
    // (int16) / 64-bit (float) load brings a pair; the loads of a batch are all issued before the
    // first conversion (one DRAM round trip per batch instead of one per sample).
    const int nstart = 2 * (m0 - (kDecHalf - 1)) + 190;
    {
        using PairT = typename std::conditional<I16, uint32_t, float2>::type;
        const unsigned char* base = (const unsigned char*)in + (clip * in_stride) * (I16 ? 2 : 4);
        const bool pair_ok = (reinterpret_cast<uintptr_t>(base) & (sizeof(PairT) - 1)) == 0;
        constexpr int kBatch = 10;
        constexpr int kIters = (kDecLocal + kDecThreads - 1) / kDecThreads;      // 19
        // Interior tiles (all but a clip's first and last): no bounds logic, every load of the tile in
        // flight before the first conversion — the general path below spent a third of the kernel's
        // time on index arithmetic and two DRAM round trips.
        const bool interior = pair_ok && nstart >= 0 && nstart + 2 * kDecLocal <= in_len;    // CTA-uniform
        if (interior && (I16 || (reinterpret_cast<uintptr_t>(base + (size_t)nstart * 4) & 15) == 0)) {
            if constexpr (I16) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(base + (size_t)nstart * 2) + tid;
                uint32_t raw[kIters];
#pragma unroll
                for (int u = 0; u < kIters; ++u)
                    if (tid + u * kDecThreads < kDecLocal) raw[u] = __ldg(src + u * kDecThreads);
#pragma unroll
                for (int u = 0; u < kIters; ++u) {
                    const int l = tid + u * kDecThreads;
                    if (l < kDecLocal)
                        s2[l] = make_float2(__int2float_rn((int)(short)(raw[u] & 0xffffu)) * (1.0f / 32768.0f),
                                            __int2float_rn((int)raw[u] >> 16) * (1.0f / 32768.0f));
                }
            } else {
                static_assert(kDecLocal % 2 == 0, "two sample pairs per 128-bit load");
                constexpr int kQuads = kDecLocal / 2, kIt4 = (kQuads + kDecThreads - 1) / kDecThreads;   // 1184, 10
                const float4* src = reinterpret_cast<const float4*>(base + (size_t)nstart * 4) + tid;
                float4 raw[kIt4];
#pragma unroll
                for (int u = 0; u < kIt4; ++u)
                    if (tid + u * kDecThreads < kQuads) raw[u] = __ldg(src + u * kDecThreads);
#pragma unroll
                for (int u = 0; u < kIt4; ++u) {
                    const int q = tid + u * kDecThreads;
                    if (q < kQuads) *reinterpret_cast<float4*>(&s2[2 * q]) = raw[u];
                }
            }
        } else
#pragma unroll 1
        for (int k0 = 0; k0 < kIters; k0 += kBatch) {
            PairT raw[kBatch];
            int state[kBatch];                       // 0: outside the tile, 1: pair load, 2: edge / unaligned
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int l = tid + (k0 + u) * kDecThreads;
                const int nn = nstart + 2 * l;
                state[u] = (k0 + u < kIters && l < kDecLocal) ? ((pair_ok && nn >= 0 && nn + 1 < in_len) ? 1 : 2) : 0;
                if (state[u] == 1) raw[u] = __ldg(reinterpret_cast<const PairT*>(base + (size_t)nn * (I16 ? 2 : 4)));
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (state[u] == 0) continue;
                const int l = tid + (k0 + u) * kDecThreads;
                const int nn = nstart + 2 * l;
                float vb, va;
                if (state[u] == 1) {
                    if constexpr (I16) {
                        const uint32_t w = raw[u];
                        vb = __int2float_rn((int)(short)(w & 0xffffu)) * (1.0f / 32768.0f);
                        va = __int2float_rn((int)w >> 16) * (1.0f / 32768.0f);
                    } else {
                        vb = raw[u].x; va = raw[u].y;
                    }
                } else {
                    auto one = [&](int n1) -> float {
                        if (n1 < 0 || n1 >= in_len) return 0.f;
                        if (I16) return __int2float_rn((int)((const int16_t*)base)[n1]) * (1.0f / 32768.0f);
                        return ((const float*)base)[n1];
                    };
                    vb = one(nn); va = one(nn + 1);
                }
                s2[l] = make_float2(vb, va);
            }
        }
    }
    __syncthreads();

    // Accumulation order and precision matter here: six cascaded stages feed CQT bins that sit
    // 80 dB below the clip's peak.  The 32 taps around the centre carry almost all of the filter's
    // energy and are accumulated in fp64 (B200 issues DFMA at ~0.64 of the FFMA rate,
    // tools/ubench/dfma.cu); the 351 small outer taps run in packed fp32 from the tails inwards so
    // their running sums stay small.  The oracle accumulates everything in float64
    // (oracle/librosa_restated.py: decimate2).
    // Tap pairs come from constant memory, indexed by the loop-uniform step: the shared-memory
    // pipe only carries the samples.
    const int lb = tid * kDecR + (kDecHalf - 1);     // local index of the pair (B[m_0], A[m_0]) of output r = 0
    float2 acc2[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) acc2[r] = make_float2(0.f, 0.f);
    auto step_f32 = [&](int i0) {                    // tap pairs [i0, i0 + kDecU)
        float2 x[kDecR + kDecU - 1];
        const float2* w = s2 + (lb - i0 - (kDecU - 1));
#pragma unroll
        for (int d = 0; d < kDecR + kDecU - 1; ++d) x[d] = w[d];
#pragma unroll
        for (int u = 0; u < kDecU; ++u) {
            const float2 h = kDecTapP[i0 + u];
#pragma unroll
            for (int r = 0; r < kDecR; ++r) acc2[r] = __ffma2_rn(x[r - u + kDecU - 1], h, acc2[r]);
        }
    };
#pragma unroll 1
    for (int c = 0; c < kDecOuter / kDecU; ++c) {
        step_f32(c * kDecU);
        step_f32(kDecHalf - (c + 1) * kDecU);
    }
    float acc[kDecR];
    double accd[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) { acc[r] = acc2[r].x + acc2[r].y; accd[r] = 0.0; }
    // centre pairs [88, 104) in fp64, one polyphase component at a time (a window of doubles for both
    // would not fit the register budget); the 32-bit loads of a phase are 2-way bank conflicted,
    // which this short pass can afford.
    constexpr int kMidU = 8;
    const float* s1 = reinterpret_cast<const float*>(s2);
#pragma unroll 1
    for (int i0 = kDecOuter; i0 < kDecOuter + kDecMid; i0 += kMidU) {
#pragma unroll 1
        for (int ph = 0; ph < 2; ++ph) {             // 0: odd taps on B, 1: even taps on A
            double x[kDecR + kMidU - 1];
            const float* w = s1 + 2 * (lb - i0 - (kMidU - 1)) + ph;
#pragma unroll
            for (int d = 0; d < kDecR + kMidU - 1; ++d) x[d] = (double)w[2 * d];
#pragma unroll
            for (int u = 0; u < kMidU; ++u) {
                const double h = (double)kDecTapF[2 * (i0 + u) + 1 - ph];
#pragma unroll
                for (int r = 0; r < kDecR; ++r) accd[r] = fma(h, x[r - u + kMidU - 1], accd[r]);
            }
        }
    }
    // results leave through shared memory so that the global stores are coalesced
    __syncthreads();
    float* so = reinterpret_cast<float*>(s2);
#pragma unroll
    for (int r = 0; r < kDecR; ++r)
        so[tid * kDecR + r] = (float)((accd[r] + (double)acc[r]) * 1.41421356237309504880);
    __syncthreads();
    float* o = out + clip * out_stride + m0;
    const int n_valid = min(kDecTile, out_len - m0);
    for (int i = tid; i < n_valid; i += kDecThreads) o[i] = so[i];
}

// ------------------------------------------------------------------------------------------
// One octave: rectangular-window STFT (centre zero-pad) -> banded complex basis -> |.|/sqrt(len)
// ------------------------------------------------------------------------------------------
struct OctParams {
    const void* in; size_t in_stride; int in_len;
    int hop, n_frames, n_rows, row0, nnz;
    const float2* tw; const float2* tw2;
    const float2* basis; const int* k0; const int* cnt; const int* off;
    const float* inv_sqrt_len;
    float* out; size_t out_stride;            // out[clip*out_stride + row*n_frames + t]
    unsigned int* clip_max; unsigned int* clip_min;
};

template <int LOG2NC> struct OctCfg {
    using G = FftGeom<LOG2NC>;
    static constexpr int F = (LOG2NC <= 8) ? 32 : 16;
    static constexpr int FR = kThreads / G::T;
    static constexpr int ROUNDS = F / FR;
    static constexpr int SSTRIDE = G::NC + 1;        // float2 per frame in the spectrum tile
};

static size_t oct_smem_bytes(int log2nc, int hop, int n_rows, int nnz) {
    const int NC = 1 << log2nc, n_fft = 2 * NC, T = NC / 16;
    const int F = (log2nc <= 8) ? 32 : 16, FR = kThreads / T;
    const size_t fstride = hop >= n_fft ? (size_t)n_fft : (size_t)hop;
    size_t cl = fstride * (F - 1) + n_fft;
    cl = (cl + 7) & ~(size_t)7;
    size_t b = cl * 4 + (size_t)FR * (NC + NC / 16) * 8 + (size_t)F * (NC + 1) * 8;
    b = (b + 15) & ~(size_t)15;
    const int R1 = log2nc >= 8 ? 16 : (1 << (log2nc - 4));
    b += (size_t)NC * 8 + (size_t)(NC / 2 + 1) * 8 + (size_t)(16 / R1) * (R1 - 1) * T * 8 + (size_t)nnz * 8 + (size_t)n_rows * 12 + 128 * 4;
    return b + 64;
}

template <int LOG2NC, bool I16>
__global__ void __launch_bounds__(kThreads) cqt_octave_kernel(OctParams p) {
    using G = FftGeom<LOG2NC>;
    using C = OctCfg<LOG2NC>;
    constexpr int NC = G::NC, NFFT = G::NFFT, T = G::T, F = C::F, FR = C::FR;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool per_frame = p.hop >= NFFT;               // frames do not overlap: stage frame by frame
    const int fstride = per_frame ? NFFT : p.hop;
    const int cl = (fstride * (F - 1) + NFFT + 7) & ~7;
    float* s_audio = reinterpret_cast<float*>(smem_raw);
    float2* s_xch = reinterpret_cast<float2*>(s_audio + cl);
    float2* s_spec = s_xch + FR * G::XSTRIDE;
    const int off_tw = ((cl * 4 + FR * G::XSTRIDE * 8 + F * C::SSTRIDE * 8) + 15) & ~15;   // no integer round trip
    float2* s_tw = reinterpret_cast<float2*>(smem_raw + off_tw);
    float2* s_tw2 = s_tw + NC;
    float2* s_twp = s_tw2 + NC / 2 + 1;
    float2* s_basis = s_twp + FftTwp<LOG2NC>::SIZE;
    int* s_k0 = reinterpret_cast<int*>(s_basis + p.nnz);
    int* s_cnt = s_k0 + p.n_rows;
    int* s_off = s_cnt + p.n_rows;
    float* s_red = reinterpret_cast<float*>(s_off + p.n_rows);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NC; i += kThreads) s_tw[i] = p.tw[i];
    for (int i = tid; i < NC / 2 + 1; i += kThreads) s_tw2[i] = p.tw2[i];
    FftTwp<LOG2NC>::fill(s_twp, p.tw, tid, kThreads);
    for (int i = tid; i < p.nnz; i += kThreads) s_basis[i] = p.basis[i];
    for (int i = tid; i < p.n_rows; i += kThreads) { s_k0[i] = p.k0[i]; s_cnt[i] = p.cnt[i]; s_off[i] = p.off[i]; }

    const size_t clip = blockIdx.y;
    const int t0 = blockIdx.x * F;
    const int n = p.in_len;
    // ---- stage samples (zero outside [0, n)) ---------------------------------------------------
    // Vector groups (4 floats / 2 int16) with all the loads of a batch issued before the first
    // conversion: one DRAM round trip per batch of eight instead of one per sample (this loop, not
    // the FFT, set the octave kernels' time).  Segment starts are multiples of 8 samples, so a group
    // is either inside [0, n) and aligned, or it takes the scalar edge path.
    {
        using VecT = typename std::conditional<I16, uint32_t, float4>::type;
        constexpr int V = I16 ? 2 : 4;
        constexpr int kBatch = 8;
        const int seg_len = per_frame ? NFFT : cl;                 // samples per contiguous segment
        const int gps = seg_len / V;                               // groups per segment
        const int total = (per_frame ? F : 1) * gps;
        const unsigned char* base = (const unsigned char*)p.in + (clip * p.in_stride) * (I16 ? 2 : 4);
        const bool vec_ok = (reinterpret_cast<uintptr_t>(base) & (sizeof(VecT) - 1)) == 0;
        const int c0 = t0 * p.hop - NFFT / 2;
#pragma unroll 1
        for (int g0 = tid; g0 < total; g0 += kBatch * kThreads) {
            VecT raw[kBatch];
            int src[kBatch], state[kBatch];                        // 0: past the end, 1: vector load, 2: edge / unaligned
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int g = g0 + u * kThreads;
                const int f = per_frame ? g / (NFFT / V) : 0;
                const int o = (per_frame ? g % (NFFT / V) : g) * V;
                src[u] = c0 + f * p.hop + o;
                const bool al = vec_ok && ((src[u] * (I16 ? 2 : 4)) & (int)(sizeof(VecT) - 1)) == 0;   // odd hops
                state[u] = g < total ? ((al && src[u] >= 0 && src[u] + V <= n) ? 1 : 2) : 0;
                if (state[u] == 1) raw[u] = __ldg(reinterpret_cast<const VecT*>(base + (size_t)src[u] * (I16 ? 2 : 4)));
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (state[u] == 0) continue;
                const int g = g0 + u * kThreads;
                float* dst = s_audio + g * V;                      // per_frame: f * NFFT + o == g * V as well
                float v[V];
                if (state[u] == 1) {
                    if constexpr (I16) {
                        const uint32_t w = raw[u];
                        v[0] = __int2float_rn((int)(short)(w & 0xffffu)) * (1.0f / 32768.0f);
                        v[1] = __int2float_rn((int)w >> 16) * (1.0f / 32768.0f);
                    } else {
                        v[0] = raw[u].x; v[1] = raw[u].y; v[2] = raw[u].z; v[3] = raw[u].w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const int s1 = src[u] + e;
                        v[e] = 0.f;
                        if (s1 >= 0 && s1 < n) {
                            if (I16) v[e] = __int2float_rn((int)((const int16_t*)base)[s1]) * (1.0f / 32768.0f);
                            else v[e] = ((const float*)base)[s1];
                        }
                    }
                }
                if constexpr (I16) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
                else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    __syncthreads();

    // ---- packed real FFT per frame -> spectrum tile -------------------------------------------
    const int j = tid % T, slot = tid / T;
    const bool even = (fstride & 1) == 0;
#pragma unroll 1
    for (int r = 0; r < C::ROUNDS; ++r) {
        const int f = r * FR + slot;
        float2* xb = s_xch + slot * G::XSTRIDE;
        {
            float2 v[16];
            const float* a = s_audio + f * fstride + 2 * j;
            if (even) {
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = *reinterpret_cast<const float2*>(a + 2 * T * t);
            } else {
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = make_float2(a[2 * T * t], a[2 * T * t + 1]);
            }
            Dft<16>::run(v);
#pragma unroll
            for (int t = 0; t < 16; ++t) xb[xpad(16 * j + t)] = v[t];
        }
        frame_sync<T>();
        fft_tail_passes<LOG2NC, false>(xb, s_tw, nullptr, s_twp, j);
        {
            float2* sp = s_spec + f * C::SSTRIDE;
#pragma unroll
            for (int r2 = 0; r2 < 8; ++r2) {
                const int k = j + T * r2;
                float2 xk, xnk;
                rfft_split(xb[xpad(k)], xb[xpad((NC - k) & (NC - 1))], s_tw2[k], xk, xnk);
                sp[k] = make_float2(0.5f * xk.x, 0.5f * xk.y);
                sp[NC - k] = make_float2(0.5f * xnk.x, 0.5f * xnk.y);
            }
            if (j == 0) {
                const float2 A = xb[xpad(NC / 2)];
                sp[NC / 2] = make_float2(A.x, -A.y);
            }
        }
        frame_sync<T>();
    }
    __syncthreads();

    // ---- banded complex basis product: item = (row, frame), frame fastest ---------------------
    float vmax = 0.f, vmin = 3.0e38f;
    for (int i = tid; i < p.n_rows * F; i += kThreads) {
        const int b = i / F, f = i % F;
        const float2* x = s_spec + f * C::SSTRIDE + s_k0[b];
        const float2* g = s_basis + s_off[b];
        const int cnt = s_cnt[b];
        float ar = 0.f, ai = 0.f;
        for (int qk = 0; qk < cnt; ++qk) {
            const float2 gg = g[qk], xx = x[qk];
            ar = fmaf(gg.x, xx.x, ar); ar = fmaf(-gg.y, xx.y, ar);
            ai = fmaf(gg.x, xx.y, ai); ai = fmaf(gg.y, xx.x, ai);
        }
        const int t = t0 + f;
        if (t < p.n_frames) {
            const int row = p.row0 + b;
            const float mag = sqrtf(ar * ar + ai * ai) * p.inv_sqrt_len[row];
            p.out[clip * p.out_stride + (size_t)row * p.n_frames + t] = mag;
            vmax = fmaxf(vmax, mag);
            vmin = fminf(vmin, mag);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    }
    if (lane == 0) { s_red[warp] = vmax; s_red[32 + warp] = vmin; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kThreads / 32; ++w) { vmax = fmaxf(vmax, s_red[w]); vmin = fminf(vmin, s_red[32 + w]); }
        atomicMax(p.clip_max + clip, __float_as_uint(vmax));     // magnitudes are >= 0
        atomicMin(p.clip_min + clip, __float_as_uint(vmin));
    }
}

// amplitude_to_db(ref=np.max, amin=1e-5, top_db) then _normalize (deep.py:259-260)
__global__ void __launch_bounds__(kThreads) cqt_finalize_kernel(float* out, size_t out_stride, int total,
                                                               const unsigned int* clip_max,
                                                               const unsigned int* clip_min, float top_db) {
    const size_t clip = blockIdx.x;
    const float mx = __uint_as_float(clip_max[clip]), mn = __uint_as_float(clip_min[clip]);
    auto db = [](float pw) { return 3.01029995663981195f * __log2f(fmaxf(pw, 1e-10f)); };
    const float vref = db(mx * mx);
    const float lo = fmaxf(db(mn * mn) - vref, -top_db);
    const float range = (0.0f - lo) + 1e-8f;
    float* o = out + clip * out_stride;
    for (int i = threadIdx.x; i < total; i += kThreads) {
        const float m = o[i];
        const float L = fmaxf(db(m * m) - vref, -top_db);
        o[i] = __fdiv_rn(L - lo, range);
    }
}

__global__ void cqt_init_minmax(unsigned int* mx, unsigned int* mn, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mx[i] = 0u; mn[i] = 0x7f7fffffu; }
}

template <int LOG2NC>
cudaError_t launch_oct(const OctParams& p, bool i16, dim3 grid, size_t smem, cudaStream_t st) {
    if (i16) {
        auto k = cqt_octave_kernel<LOG2NC, true>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kThreads, smem, st>>>(p);
    } else {
        auto k = cqt_octave_kernel<LOG2NC, false>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kThreads, smem, st>>>(p);
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t up(const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

}  // namespace

#define CQ_TRY(expr)                                                                       \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            *err = std::string(#expr) + ": " + cudaGetErrorString(e__);                    \
            return e__ == cudaErrorMemoryAllocation ? B2A_ENOMEM : B2A_ECUDA;              \
        }                                                                                  \
    } while (0)

void cqt_device_free(CqtDevice* d) {
    cudaFree(d->scratch); cudaFree(d->taps); cudaFree(d->inv_sqrt_len);
    cudaFree(d->clip_max); cudaFree(d->clip_min);
    for (int i = 0; i < 16; ++i) { cudaFree(d->tw[i]); cudaFree(d->tw2[i]); }
    for (auto& o : d->oct) { cudaFree(o.basis); cudaFree(o.k0); cudaFree(o.cnt); cudaFree(o.off); }
    *d = CqtDevice();
}

int cqt_device_init(const CqtPlan& plan, const b2a_config& cfg, int sm_count, size_t smem_optin,
                    CqtDevice* dev, std::string* err) {
    dev->sm_count = sm_count;
    // decimator taps (float32 of the double design)
    {
        std::vector<float> t;
        for (double v : decimator_taps()) t.push_back((float)v);
        CQ_TRY(up(t, &dev->taps));
    }
    {
        std::vector<float> isl;
        for (double l : plan.lengths) isl.push_back((float)(1.0 / std::sqrt(l)));
        CQ_TRY(up(isl, &dev->inv_sqrt_len));
    }
    // scratch layout per clip: early-downsample outputs, then one signal per decimated octave
    size_t off = 0;
    int len = cfg.n_samples;
    for (int e = 0; e < plan.n_early; ++e) {
        len = (len + 1) / 2;
        dev->early_offs.push_back(off);
        dev->early_lens.push_back(len);
        off += ((size_t)len + 3) & ~(size_t)3;
    }
    dev->oct.resize(plan.n_octaves);
    size_t cur_off = plan.n_early ? dev->early_offs.back() : (size_t)-1;   // (size_t)-1: the input itself
    for (int i = 0; i < plan.n_octaves; ++i) {
        const CqtOctave& o = plan.oct[i];
        CqtOctaveDev& od = dev->oct[i];
        od.sig_off = cur_off;
        od.log2nc = ilog2(o.n_fft / 2);
        if (od.log2nc < 7 || od.log2nc > 9) {
            *err = "cqt: per-octave n_fft " + std::to_string(o.n_fft) + " unsupported (256..1024)";
            return B2A_EINVAL;
        }
        // band each row of the sparsified basis
        const int n_bins = o.n_fft / 2 + 1;
        std::vector<float2> bw;
        std::vector<int> k0(o.n_rows), cnt(o.n_rows), offv(o.n_rows);
        for (int r = 0; r < o.n_rows; ++r) {
            int lo = n_bins, hi = -1;
            for (int k = 0; k < n_bins; ++k) {
                const float re = o.basis[((size_t)r * n_bins + k) * 2], im = o.basis[((size_t)r * n_bins + k) * 2 + 1];
                if (re != 0.f || im != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k); }
            }
            offv[r] = (int)bw.size();
            if (hi < 0) { k0[r] = 0; cnt[r] = 0; continue; }
            k0[r] = lo; cnt[r] = hi - lo + 1;
            for (int k = lo; k <= hi; ++k)
                bw.push_back(make_float2(o.basis[((size_t)r * n_bins + k) * 2], o.basis[((size_t)r * n_bins + k) * 2 + 1]));
        }
        od.nnz = (int)bw.size();
        if (bw.empty()) bw.push_back(make_float2(0.f, 0.f));
        CQ_TRY(up(bw, &od.basis));
        CQ_TRY(up(k0, &od.k0));
        CQ_TRY(up(cnt, &od.cnt));
        CQ_TRY(up(offv, &od.off));
        if (oct_smem_bytes(od.log2nc, o.hop, o.n_rows, od.nnz) > smem_optin) {
            *err = "cqt: octave working set exceeds shared memory";
            return B2A_EINVAL;
        }
        if (!dev->tw[od.log2nc]) {
            const int NC = o.n_fft / 2;
            std::vector<float> a = twiddles(NC, NC), b = twiddles(2 * NC, NC / 2 + 1);
            CQ_TRY(up(a, (float**)&dev->tw[od.log2nc]));
            CQ_TRY(up(b, (float**)&dev->tw2[od.log2nc]));
        }
        if (o.decimate_after && i + 1 < plan.n_octaves) {
            cur_off = off;
            off += ((size_t)((o.sig_len + 1) / 2) + 3) & ~(size_t)3;
        }
    }
    dev->scratch_per_clip = off;
    dev->chunk_clips = 1024;
    if (off) CQ_TRY(cudaMalloc((void**)&dev->scratch, dev->chunk_clips * off * sizeof(float)));
    CQ_TRY(cudaMalloc((void**)&dev->clip_max, dev->chunk_clips * sizeof(unsigned int)));
    CQ_TRY(cudaMalloc((void**)&dev->clip_min, dev->chunk_clips * sizeof(unsigned int)));
    return 0;
}

int cqt_run(const CqtPlan& plan, const b2a_config& cfg, CqtDevice* dev, const void* d_clips,
            int64_t n_clips, float* d_out, cudaStream_t st, int64_t* launches, std::string* err) {
    const bool in_i16 = cfg.input_dtype == B2A_IN_I16;
    const size_t in_elem = in_i16 ? 2 : 4;
    const int rows = cfg.n_bins, nfr = plan.n_frames;
    const size_t out_stride = (size_t)rows * nfr;
    for (int64_t c0 = 0; c0 < n_clips; c0 += dev->chunk_clips) {
        const int nb = (int)std::min<int64_t>(dev->chunk_clips, n_clips - c0);
        const void* in = (const unsigned char*)d_clips + (size_t)c0 * cfg.n_samples * in_elem;
        float* out = d_out + (size_t)c0 * out_stride;
        cqt_init_minmax<<<(nb + 255) / 256, 256, 0, st>>>(dev->clip_max, dev->clip_min, nb);
        CQ_TRY(cudaGetLastError());
        ++*launches;
        // current signal descriptor
        const void* sig = in; size_t sig_stride = cfg.n_samples; int sig_len = cfg.n_samples; bool sig_i16 = in_i16;
        auto decimate = [&](size_t dst_off, int dst_len) -> cudaError_t {
            float* dst = dev->scratch + dst_off;
            dim3 grid((dst_len + kDecTile - 1) / kDecTile, nb);
            if (sig_i16) cqt_decimate_kernel<true><<<grid, kDecThreads, 0, st>>>(sig, sig_stride, sig_len, dst, dev->scratch_per_clip, dst_len, dev->taps);
            else cqt_decimate_kernel<false><<<grid, kDecThreads, 0, st>>>(sig, sig_stride, sig_len, dst, dev->scratch_per_clip, dst_len, dev->taps);
            ++*launches;
            sig = dst; sig_stride = dev->scratch_per_clip; sig_len = dst_len; sig_i16 = false;
            return cudaGetLastError();
        };
        for (int e = 0; e < plan.n_early; ++e) CQ_TRY(decimate(dev->early_offs[e], dev->early_lens[e]));
        for (int i = 0; i < plan.n_octaves; ++i) {
            const CqtOctave& o = plan.oct[i];
            const CqtOctaveDev& od = dev->oct[i];
            OctParams p{};
            p.in = sig; p.in_stride = sig_stride; p.in_len = sig_len;
            p.hop = o.hop; p.n_frames = nfr; p.n_rows = o.n_rows; p.row0 = o.row0; p.nnz = od.nnz;
            p.tw = dev->tw[od.log2nc]; p.tw2 = dev->tw2[od.log2nc];
            p.basis = od.basis; p.k0 = od.k0; p.cnt = od.cnt; p.off = od.off;
            p.inv_sqrt_len = dev->inv_sqrt_len;
            p.out = out; p.out_stride = out_stride;
            p.clip_max = dev->clip_max; p.clip_min = dev->clip_min;
            const int F = od.log2nc <= 8 ? 32 : 16;
            dim3 grid((nfr + F - 1) / F, nb);
            const size_t smem = oct_smem_bytes(od.log2nc, o.hop, o.n_rows, od.nnz);
            cudaError_t e = cudaSuccess;
            switch (od.log2nc) {
                case 7: e = launch_oct<7>(p, sig_i16, grid, smem, st); break;
                case 8: e = launch_oct<8>(p, sig_i16, grid, smem, st); break;
                case 9: e = launch_oct<9>(p, sig_i16, grid, smem, st); break;
                default: e = cudaErrorInvalidValue;
            }
            CQ_TRY(e);
            ++*launches;
            if (o.decimate_after && i + 1 < plan.n_octaves)
                CQ_TRY(decimate(dev->oct[i + 1].sig_off, (sig_len + 1) / 2));
        }
        cqt_finalize_kernel<<<nb, kThreads, 0, st>>>(out, out_stride, rows * nfr, dev->clip_max, dev->clip_min, cfg.top_db);
        CQ_TRY(cudaGetLastError());
        ++*launches;
    }
    return 0;
}

}  // namespace b2a
