// cqt.cu — audio_cqt on the GPU: 2:1 decimation cascade, per-octave rectangular-window STFT
// times the sparsified wavelet basis, amplitude_to_db(ref=max) + min-max.  sm_100a only.
//
// Follows librosa.cqt == vqt(gamma=0) as the reference calls it (deep.py:249-260): octaves are
// processed top-down; each octave's response is  basis_o . STFT(y_o; n_fft_o, hop_o, ones)  and
// y_{o+1} = sqrt(2) * decimate2(y_o).  The decimator is this project's stand-in for soxr_hq
// (tables.h: decimator_taps; DESIGN.md "CQT decimator").
//
// The octave response is evaluated in the TIME domain: basis_o . STFT(y_o) == sum_n y_o[t hop + n - N/2] b_o[r][n]
// with b_o[r][n] = sum_k basis_o[r][k] exp(-2 pi i k n / N) built on the host in double precision from the
// same sparsified complex64 basis librosa multiplies with — a dense [frames x N] . [N x 12 complex]
// product per octave (cqt_bank_kernel).  The frequency-domain form needs every rectangular-window STFT bin
// accurate relative to ITSELF (librosa's FFT runs in float64): the wavelet spectrum then cancels a strong
// tone's leakage across ~12 bins, and a float32 FFT, whose error is relative to the frame's PEAK bin, left
// 2e-4 in the normalised features (rows of the un-decimated top octave, profiles/r2_cqt_floor.jsonl).  In
// the time domain the partial sums of an out-of-band tone stay small, so float32 accumulation holds
// 4e-5.  9.3 M real MACs per clip as packed FFMA2: CUDA cores, not tcgen05 (operands
// would need 3-way TF32 splits and an im2col of 32x-overlapping frames in the UMMA layout).
#include "cqt.h"
#include "fft_core.cuh"
#include "gen/decim_taps.inc"   // build-time generated decimator taps (gen_mel.cpp decim)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
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
#ifndef B2A_AB_DECLVL
#define B2A_AB_DECLVL 1
#endif
#if B2A_AB_DECLVL == 0
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
#else
    // Two-level accumulation: every window step (kDecU tap pairs) sums into FRESH packed accumulators whose
    // partial sums stay small, and only the step totals meet the running sum — 22 roundings at the running
    // sum's magnitude instead of 176 (level 1: float32; level 2: float32 (DECLVL 1) or float64 (DECLVL 2)).
#if B2A_AB_DECLVL == 1
    float2 tot2[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) tot2[r] = make_float2(0.f, 0.f);
#else
    double accd[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) accd[r] = 0.0;
#endif
    auto step_f32 = [&](int i0) {                    // tap pairs [i0, i0 + kDecU)
        float2 x[kDecR + kDecU - 1];
        float2 a2[kDecR];
        const float2* w = s2 + (lb - i0 - (kDecU - 1));
#pragma unroll
        for (int d = 0; d < kDecR + kDecU - 1; ++d) x[d] = w[d];
#pragma unroll
        for (int u = 0; u < kDecU; ++u) {
            const float2 h = kDecTapP[i0 + u];
#pragma unroll
            for (int r = 0; r < kDecR; ++r)
                a2[r] = u == 0 ? __fmul2_rn(x[r - u + kDecU - 1], h) : __ffma2_rn(x[r - u + kDecU - 1], h, a2[r]);
        }
#pragma unroll
        for (int r = 0; r < kDecR; ++r) {
#if B2A_AB_DECLVL == 1
            tot2[r] = __fadd2_rn(tot2[r], a2[r]);
#else
            accd[r] += (double)(a2[r].x + a2[r].y);
#endif
        }
    };
#pragma unroll 1
    for (int c = 0; c < kDecOuter / kDecU; ++c) {
        step_f32(c * kDecU);
        step_f32(kDecHalf - (c + 1) * kDecU);
    }
    float acc[kDecR];
#if B2A_AB_DECLVL == 1
    double accd[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) { acc[r] = tot2[r].x + tot2[r].y; accd[r] = 0.0; }
#else
#pragma unroll
    for (int r = 0; r < kDecR; ++r) acc[r] = 0.f;
#endif
#endif
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
// All octaves of a chunk in one launch: time-domain wavelet bank.
//   grid = (frame blocks of 256, octaves * row blocks of 12, clips), 64 threads = 2 warps.
//   Per 32-sample slice of the N-sample kernels the CTA stages an im2col tile [32 n][256 frames] of the
//   octave's signal (16-byte chunks XOR-swizzled by row: the transposing stores spread over the banks,
//   the float4 reads stay conflict-free) and the slice's coefficients [32 n][12 rows] (re, im).  Warp w
//   owns rows 6 w .. 6 w + 5 of the block; a lane owns frames 4 l .. 4 l + 3 and 128 + 4 l .. + 3: per n
//   two 128-bit sample loads, three 128-bit broadcast coefficient loads and 48 packed FFMA2
//   (acc(re, im) += x * (b_re, b_im)) — a broadcast 128-bit load still returns 512 B to the register
//   file (4 cycles of the 128 B/cycle shared-memory pipe), so 8 frames per lane are what keeps the FMA
//   pipe (24 cycles per step) ahead of the load pipe (20).  The loads that fill a tile are issued eight
//   deep per thread; 6 CTAs per SM overlap one CTA's staging with the others' arithmetic.
// ------------------------------------------------------------------------------------------
constexpr int kBankThreads = 64;
constexpr int kBankFrames = 256;             // frames per CTA: 32 lanes x 8
constexpr int kBankRows = 12;                // output rows (bins) per CTA: 2 warps x 6
constexpr int kBankSlice = 32;               // kernel samples per staged slice
constexpr size_t kBankSmem = (size_t)kBankSlice * kBankFrames * 4 + (size_t)kBankSlice * kBankRows * 8 + 64 * 4;

struct BankOct {
    const void* in; long long in_stride; int in_len; int in_i16;     // this octave's signal, per clip
    int hop, n_fft, n_rows, row0;
    const float2* coef;                       // [ceil(n_rows / 12)][n_fft][12] (re, im), zero-padded rows
};
struct BankParams {
    BankOct oct[12];
    int n_oct, n_frames;
    const float* inv_sqrt_len;
    float* out; long long out_stride;         // out[clip * out_stride + row * n_frames + t]
    unsigned int* clip_max; unsigned int* clip_min;
};

__global__ void __launch_bounds__(kBankThreads, 6) cqt_bank_kernel(const __grid_constant__ BankParams p) {
    extern __shared__ __align__(16) unsigned char bank_smem[];
    float* const s_tile = reinterpret_cast<float*>(bank_smem);                                   // [32][256] swizzled
    float2* const s_coef = reinterpret_cast<float2*>(bank_smem + (size_t)kBankSlice * kBankFrames * 4);   // [32][12]
    float* const s_red = reinterpret_cast<float*>(s_coef + kBankSlice * kBankRows);

    // which (octave, row block) this CTA serves
    int oi = 0, rb = (int)blockIdx.y;
    for (;;) {
        const int nb = (p.oct[oi].n_rows + kBankRows - 1) / kBankRows;
        if (rb < nb) break;
        rb -= nb; ++oi;
    }
    const BankOct& o = p.oct[oi];
    const int N = o.n_fft, hop = o.hop, L = o.in_len;
    const bool i16 = o.in_i16 != 0;
    const size_t clip = blockIdx.z;
    const int t0 = blockIdx.x * kBankFrames;
    const int tid = threadIdx.x, lane = tid & 31, rgrp = tid >> 5;
    const unsigned char* const base = (const unsigned char*)o.in + clip * (size_t)o.in_stride * (i16 ? 2 : 4);
    const float2* const coef = o.coef + (size_t)rb * N * kBankRows;

    float2 acc[8][6];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[i][b] = make_float2(0.f, 0.f);

    const int nvalid = min(kBankFrames, p.n_frames - t0);       // frames of this block inside the clip
    const bool vec4 = !i16 && (hop & 3) == 0 && ((N / 2) & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
    const bool vec2 = i16 && (hop & 3) == 0 && ((N / 2) & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 7) == 0;
    const int n_items = nvalid * (kBankSlice / 4);              // item = (frame f, 4 consecutive n), frame-major
    for (int n0 = 0; n0 < N; n0 += kBankSlice) {
        __syncthreads();                                       // the previous slice has been consumed
        for (int i = tid; i < kBankSlice * kBankRows; i += kBankThreads) s_coef[i] = coef[(size_t)n0 * kBankRows + i];
        // ---- im2col tile: eight loads in flight per thread, then the transposing stores --------------------
        // Interior items of a float32 octave whose rows start on 16-byte boundaries (every decimated signal
        // with hop % 4 == 0) take one 128-bit load; clip edges, the int16 top octave and odd geometries take
        // the element-wise path.  The choice per item is two compares; the index arithmetic is shared.
        constexpr int kB = 8;
        const int sbase = t0 * hop + n0 - N / 2;                // sample index of (f = 0, n = n0)
#pragma unroll 1
        for (int it0 = tid; it0 < n_items; it0 += kB * kBankThreads) {
            float4 v[kB];
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int it = it0 + u * kBankThreads;
                const int f = it >> 3, q = it & 7;
                const int s0 = sbase + f * hop + 4 * q;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it >= n_items) continue;
                if (vec4 && s0 >= 0 && s0 + 4 <= L) {
                    v[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + s0));
                } else if (vec2 && s0 >= 0 && s0 + 4 <= L) {
                    const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const int16_t*>(base) + s0));
                    v[u] = make_float4((float)(short)(w.x & 0xffffu) * (1.0f / 32768.0f), (float)((int)w.x >> 16) * (1.0f / 32768.0f),
                                       (float)(short)(w.y & 0xffffu) * (1.0f / 32768.0f), (float)((int)w.y >> 16) * (1.0f / 32768.0f));
                } else {
                    float e4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int s1 = s0 + e;
                        e4[e] = 0.f;
                        if (s1 >= 0 && s1 < L)
                            e4[e] = i16 ? (float)reinterpret_cast<const int16_t*>(base)[s1] * (1.0f / 32768.0f)
                                        : reinterpret_cast<const float*>(base)[s1];
                    }
                    v[u] = make_float4(e4[0], e4[1], e4[2], e4[3]);
                }
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int it = it0 + u * kBankThreads;
                if (it >= n_items) continue;
                const int f = it >> 3, q = it & 7;
                // element (n = 4 q + e, f) lives at row n, 16-byte chunk ((f >> 2) ^ q), word f & 3
                float* d = s_tile + (4 * q) * kBankFrames + (((f >> 2) ^ q) << 2) + (f & 3);
                d[0] = v[u].x; d[kBankFrames] = v[u].y; d[2 * kBankFrames] = v[u].z; d[3 * kBankFrames] = v[u].w;
            }
        }
        __syncthreads();
        // ---- 32 steps: 2 sample loads (8 frames), 3 coefficient loads (6 rows), 48 FFMA2 -------------------
        const float4* const tile4 = reinterpret_cast<const float4*>(s_tile);
        const float4* const c4 = reinterpret_cast<const float4*>(s_coef) + rgrp * 3;
#pragma unroll 2
        for (int r = 0; r < kBankSlice; ++r) {
            const int sw = (r >> 2) & 7;               // (= q of the staging loop)
            const float4 xa = tile4[r * (kBankFrames / 4) + (lane ^ sw)];
            const float4 xb = tile4[r * (kBankFrames / 4) + 32 + (lane ^ sw)];
            const float4 ca = c4[r * (kBankRows / 2)], cb = c4[r * (kBankRows / 2) + 1], cc = c4[r * (kBankRows / 2) + 2];
            const float2 cf[6] = {make_float2(ca.x, ca.y), make_float2(ca.z, ca.w), make_float2(cb.x, cb.y),
                                  make_float2(cb.z, cb.w), make_float2(cc.x, cc.y), make_float2(cc.z, cc.w)};
            const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int b = 0; b < 6; ++b) acc[i][b] = __ffma2_rn(make_float2(xs[i], xs[i]), cf[b], acc[i][b]);
        }
    }

    // ---- |.| / sqrt(length), per-clip extrema --------------------------------------------------------
    float vmax = 0.f, vmin = 3.0e38f;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const int lr = rb * kBankRows + rgrp * 6 + b;             // row inside the octave
        if (lr >= o.n_rows) continue;
        const int row = o.row0 + lr;
        const float sc = p.inv_sqrt_len[row];
        float* const orow = p.out + clip * (size_t)p.out_stride + (size_t)row * p.n_frames;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = t0 + (i >> 2) * 128 + 4 * lane + (i & 3);
            if (t < p.n_frames) {
                const float mag = sqrtf(acc[i][b].x * acc[i][b].x + acc[i][b].y * acc[i][b].y) * sc;
                orow[t] = mag;
                vmax = fmaxf(vmax, mag);
                vmin = fminf(vmin, mag);
            }
        }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, sft));
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, sft));
    }
    if (lane == 0) { s_red[rgrp] = vmax; s_red[32 + rgrp] = vmin; }
    __syncthreads();
    if (tid == 0) {
        vmax = fmaxf(vmax, s_red[1]); vmin = fminf(vmin, s_red[33]);
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

template <typename T>
cudaError_t up(const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}


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
    for (auto& o : d->oct) cudaFree(o.coef);
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
    if (plan.n_octaves > 12) { *err = "cqt: more than 12 octaves"; return B2A_EINVAL; }
    if (kBankSmem > smem_optin) { *err = "cqt: wavelet-bank tile exceeds shared memory"; return B2A_EINVAL; }
    CQ_TRY(cudaFuncSetAttribute(cqt_bank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBankSmem));

    dev->oct.resize(plan.n_octaves);
    size_t cur_off = plan.n_early ? dev->early_offs.back() : (size_t)-1;   // (size_t)-1: the input itself
    for (int i = 0; i < plan.n_octaves; ++i) {
        const CqtOctave& o = plan.oct[i];
        CqtOctaveDev& od = dev->oct[i];
        od.sig_off = cur_off;
        if (o.n_fft % kBankSlice != 0 || o.n_fft < kBankSlice) {
            *err = "cqt: per-octave n_fft " + std::to_string(o.n_fft) + " unsupported (multiple of 64)";
            return B2A_EINVAL;
        }
        // time-domain kernels b[r][n] = sum_k basis[r][k] exp(-2 pi i k n / N), k = 0..N/2 (the bins the
        // reference multiplies with), in double from the float32 basis entries; laid out
        // [row block of 12][n][12] (re, im) so that a slice of 64 n is one contiguous copy
        {
            const int N = o.n_fft, n_bins = N / 2 + 1;
            const int n_blocks = (o.n_rows + kBankRows - 1) / kBankRows;
            std::vector<float2> coef((size_t)n_blocks * N * kBankRows, make_float2(0.f, 0.f));
            std::vector<double> cs(N), sn(N);
            for (int m = 0; m < N; ++m) { cs[m] = std::cos(2.0 * M_PI * m / N); sn[m] = std::sin(2.0 * M_PI * m / N); }
            for (int r = 0; r < o.n_rows; ++r) {
                const float* br = &o.basis[(size_t)(o.filt0 + r) * n_bins * 2];
                for (int n = 0; n < N; ++n) {
                    double re = 0.0, im = 0.0;
                    for (int k = 0; k < n_bins; ++k) {
                        const double gr = br[2 * k], gi = br[2 * k + 1];
                        if (gr == 0.0 && gi == 0.0) continue;
                        const int m = (int)(((long long)k * n) % N);       // exp(-2 pi i k n / N) = cs[m] - i sn[m]
                        re += gr * cs[m] + gi * sn[m];
                        im += gi * cs[m] - gr * sn[m];
                    }
                    coef[((size_t)(r / kBankRows) * N + n) * kBankRows + (r % kBankRows)] = make_float2((float)re, (float)im);
                }
            }
            od.n_blocks = n_blocks;
            CQ_TRY(up(coef, &od.coef));
        }
        if (o.decimate_after && i + 1 < plan.n_octaves) {
            cur_off = off;
            off += ((size_t)((o.sig_len + 1) / 2) + 3) & ~(size_t)3;
        }
    }
    dev->scratch_per_clip = off;
    // Clips per pass of the cascade.  The decimated signals of a pass are written once and read twice (next
    // decimation, wavelet bank): a pass whose scratch fits the 126 MB L2 keeps those reads out of HBM.
    dev->chunk_clips = 1024;
    if (const char* e = std::getenv("B2A_CQT_CHUNK")) { const int v = std::atoi(e); if (v >= 1 && v <= 65535) dev->chunk_clips = v; }
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
        // the decimation cascade first (every octave's signal then sits in the chunk's scratch), ...
        BankParams bp{};
        int n_rb = 0;
        for (int i = 0; i < plan.n_octaves; ++i) {
            const CqtOctave& o = plan.oct[i];
            const CqtOctaveDev& od = dev->oct[i];
            BankOct& bo = bp.oct[i];
            bo.in = sig; bo.in_stride = (long long)sig_stride; bo.in_len = sig_len; bo.in_i16 = sig_i16 ? 1 : 0;
            bo.hop = o.hop; bo.n_fft = o.n_fft; bo.n_rows = o.n_rows; bo.row0 = o.row0; bo.coef = od.coef;
            n_rb += od.n_blocks;
            if (o.decimate_after && i + 1 < plan.n_octaves)
                CQ_TRY(decimate(dev->oct[i + 1].sig_off, (sig_len + 1) / 2));
        }
        // ... then every octave's wavelet bank in one launch
        bp.n_oct = plan.n_octaves; bp.n_frames = nfr;
        bp.inv_sqrt_len = dev->inv_sqrt_len;
        bp.out = out; bp.out_stride = (long long)out_stride;
        bp.clip_max = dev->clip_max; bp.clip_min = dev->clip_min;
        {
            const dim3 grid((nfr + kBankFrames - 1) / kBankFrames, n_rb, nb);
            cqt_bank_kernel<<<grid, kBankThreads, kBankSmem, st>>>(bp);
            CQ_TRY(cudaGetLastError());
            ++*launches;
        }
        cqt_finalize_kernel<<<nb, kThreads, 0, st>>>(out, out_stride, rows * nfr, dev->clip_max, dev->clip_min, cfg.top_db);
        CQ_TRY(cudaGetLastError());
        ++*launches;
    }
    return 0;
}

}  // namespace b2a
