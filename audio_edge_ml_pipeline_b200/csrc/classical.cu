// classical.cu — the `audio_classical` extractor (reference:
// src/preprocessing/feature_extraction/audio/classical.py:272-355) on the shared STFT front end.  sm_100a only.
//
// The reference calls twelve librosa functions per clip, each of which recomputes the STFT; here one
// persistent CTA owns a clip at a time and makes ONE pass over its frames:
//   per tile of F frames: stage the samples (fp32), packed real FFT per frame (fft_core.cuh) -> power tile;
//     mel bands -> dB (for the MFCCs), the power rows -> per-CTA scratch (re-read for the chroma product once the
//     clip's tuning is known), and one warp per frame: magnitude sums (centroid, bandwidth, flatness), the
//     roll-off scan, the seven contrast bands' q smallest / largest magnitudes, the piptrack candidates
//     (parabolic peak interpolation) for the tuning estimate, rms from the staged samples, zero crossings over
//     the frame_length = 2048 window straight from the clip;
//   per clip: top_db clip + DCT-II -> MFCC series, Savitzky-Golay deltas (width 9, edges replicated: the
//     polynomial fit's derivative of matching order is constant over the edge window), the tuning estimate
//     (exact median of the candidate magnitudes by radix select, 100-bin residual histogram, first argmax),
//     chroma with the filterbank of that tuning (100 banks precomputed in double, librosa.filters.chroma),
//     tonnetz, and mean / std of every row -> [6 n_mfcc + 62] floats in the reference's canonical order.
// Frame-level sums that the reference evaluates in float64 (centroid, bandwidth, contrast dB, tonnetz, all
// aggregations) run in double here too: they are a few thousand DFMA per frame next to the FFT.
#include "classical.h"
#include "fft_core.cuh"
#include "front_stage.cuh"

#include <cstdint>

namespace b2a {

namespace {

constexpr int kBands = 7;
constexpr int kCT = 512;                       // threads per CTA: sixteen warps, one resident CTA per SM
constexpr int kWarps = kCT / 32;
constexpr float kTiny = 1.17549435e-38f;        // np.finfo(float32).tiny (librosa.util.normalize threshold)
constexpr int kZcrFrame = 2048;                 // librosa.feature.zero_crossing_rate default frame_length

template <int LOG2NC> struct ClsCfg {
    using G = FftGeom<LOG2NC>;
    static constexpr int F = (LOG2NC <= 8) ? 32 : 16;
    static constexpr int FR = kCT / G::T;
    static constexpr int ROUNDS = (F + FR - 1) / FR;
    static constexpr int CH = G::NC / 32 + 1;   // contiguous bins per lane (odd: conflict-free stride)
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct Scratch {
    float *P, *dbm, *mf, *sc, *pk, *vl, *cp, *cm;
};
__device__ __forceinline__ Scratch carve(float* base, int nfr, int nb, int n_mels, int n_mfcc, int cap) {
    Scratch s;
    s.P = base; base += (size_t)nfr * nb;
    s.dbm = base; base += (size_t)n_mels * nfr;
    s.mf = base; base += (size_t)n_mfcc * nfr;
    s.sc = base; base += (size_t)6 * nfr;
    s.pk = base; base += (size_t)kBands * nfr;
    s.vl = base; base += (size_t)kBands * nfr;
    s.cp = base; base += cap;
    s.cm = base;
    return s;
}

// mean and population std of n values produced by f(t), two passes in double, by one warp
template <typename Fn>
__device__ __forceinline__ void row_stats(int n, int lane, Fn f, float* mean_out, float* std_out) {
    double s = 0.0;
    for (int t = lane; t < n; t += 32) s += (double)f(t);
    const double mean = warp_sum_d(s) / n;
    double q = 0.0;
    for (int t = lane; t < n; t += 32) { const double d = (double)f(t) - mean; q += d * d; }
    const double var = warp_sum_d(q) / n;
    if (lane == 0) { *mean_out = (float)mean; *std_out = (float)sqrt(var); }
}

template <int LOG2NC, bool I16>
__global__ void __launch_bounds__(kCT, 1) classical_kernel(ClassicalParams p) {
    using G = FftGeom<LOG2NC>;
    using C = ClsCfg<LOG2NC>;
    constexpr int NC = G::NC, NFFT = G::NFFT, T = G::T, F = C::F, FR = C::FR, NB = NC + 1, CH = C::CH;
    constexpr int SLOTS = FR;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cl = (p.hop * (F - 1) + NFFT + 7) & ~7;
    float* s_audio = reinterpret_cast<float*>(smem_raw);
    float2* s_xch = reinterpret_cast<float2*>(s_audio + cl);
    float* s_pow = reinterpret_cast<float*>(s_xch + SLOTS * G::XSTRIDE);
    const int off_tw = ((cl * 4 + SLOTS * G::XSTRIDE * 8 + F * G::PSTRIDE * 4) + 15) & ~15;
    float2* s_tw = reinterpret_cast<float2*>(smem_raw + off_tw);
    float2* s_tw2 = s_tw + NC;
    float2* s_twp = s_tw2 + NC / 2 + 1;
    int* s_k0 = reinterpret_cast<int*>(s_twp + FftTwp<LOG2NC>::SIZE);
    int* s_cnt = s_k0 + p.n_mels;
    int* s_off = s_cnt + p.n_mels;
    float* s_w = reinterpret_cast<float*>(s_off + p.n_mels);
    float* s_red = s_w + p.mel_nnz;                          // 64 floats
    int* s_hist = reinterpret_cast<int*>(s_red + 64);        // 256
    int* s_misc = s_hist + 256;                              // [0] candidate count, [1..3] select state, [4] tuning index
    double* s_acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(s_misc + 8) + 7) & ~(uintptr_t)7);   // [kWarps][18][2] chroma / tonnetz partial sums

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NC; i += kCT) s_tw[i] = p.tw[i];
    for (int i = tid; i < NC / 2 + 1; i += kCT) s_tw2[i] = p.tw2[i];
    FftTwp<LOG2NC>::fill(s_twp, p.tw, tid, kCT);
    for (int i = tid; i < p.n_mels; i += kCT) { s_k0[i] = p.mel_k0[i]; s_cnt[i] = p.mel_cnt[i]; s_off[i] = p.mel_off[i]; }
    for (int i = tid; i < p.mel_nnz; i += kCT) s_w[i] = p.mel_w[i];

    const int j = tid % T, slot = tid / T;
    const float2* const win = reinterpret_cast<const float2*>(p.window) + j;   // pairs (w[2n], w[2n+1]), n = j + T t: L1-resident
    __syncthreads();

    int n = p.n_samples, nfr = p.n_frames;                  // per clip when ragged
    const int n_mels = p.n_mels, K = p.n_mfcc, hop = p.hop;
    const bool hop_even = (hop & 1) == 0;
    const size_t esz = I16 ? 2 : 4;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(p.clips) & 15) == 0;
    const Scratch S = carve(p.scratch + (size_t)blockIdx.x * p.scratch_per_cta, p.n_frames, NB, n_mels, K, p.cand_cap);
    const double bin_hz = (double)p.sample_rate / NFFT;
    const int rows_out = 6 * K + 62;

    for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
        if (p.rag_len) { n = p.rag_len[clip]; nfr = 1 + n / hop; }
        const long long clip_elem0 = p.rag_len ? p.rag_in_off[clip] : clip * (long long)n;
        const void* cptr = (const unsigned char*)p.clips + (size_t)clip_elem0 * esz;
        float* const outv = p.rag_len ? p.out + p.rag_out_off[clip] : p.out + (size_t)clip * rows_out;
        float vmax = -3.0e38f;
        if (tid == 0) s_misc[0] = 0;
        __syncthreads();

        for (int t0 = 0; t0 < nfr; t0 += F) {
            stage_audio<I16, kCT>(s_audio, cptr, clip_elem0, t0 * hop - NFFT / 2, cl, n, 0, base_aligned);
            __syncthreads();
#pragma unroll 1
            for (int r = 0; r < C::ROUNDS; ++r) {
                const int f = r * FR + slot;
                float2* xb = s_xch + slot * G::XSTRIDE;
                {
                    float2 v[16];
                    const float* a = s_audio + f * hop + 2 * j;
                    if (hop_even) {
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            const float2 x = *reinterpret_cast<const float2*>(a + 2 * T * t);
                            const float2 w = __ldg(win + T * t);
                            v[t] = make_float2(x.x * w.x, x.y * w.y);
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            const float2 w = __ldg(win + T * t);
                            v[t] = make_float2(a[2 * T * t] * w.x, a[2 * T * t + 1] * w.y);
                        }
                    }
                    Dft<16>::run(v);
#pragma unroll
                    for (int t = 0; t < 16; ++t) xb[xpad(16 * j + t)] = v[t];
                }
                frame_sync<T>();
                fft_tail_passes<LOG2NC, false>(xb, s_tw, nullptr, s_twp, j);
                {
                    float* pw = s_pow + f * G::PSTRIDE;
#pragma unroll
                    for (int r2 = 0; r2 < 8; ++r2) {
                        const int k = j + T * r2;
                        const float2 A = xb[xpad(k)];
                        const float2 B = xb[xpad((NC - k) & (NC - 1))];
                        float2 xk, xnk;
                        rfft_split(A, B, s_tw2[k], xk, xnk);
                        pw[k] = 0.25f * (xk.x * xk.x + xk.y * xk.y);
                        pw[NC - k] = 0.25f * (xnk.x * xnk.x + xnk.y * xnk.y);
                    }
                    if (j == 0) {
                        const float2 A = xb[xpad(NC / 2)];
                        pw[NC / 2] = A.x * A.x + A.y * A.y;
                    }
                }
                frame_sync<T>();
            }
            __syncthreads();

            // ---- mel bands -> dB (librosa.feature.mfcc: power_to_db(melspectrogram)) -------------------
            for (int i = tid; i < n_mels * F; i += kCT) {
                const int m = i / F, f = i % F;
                const float* pf = s_pow + f * G::PSTRIDE + s_k0[m];
                const float* w = s_w + s_off[m];
                const int cnt = s_cnt[m];
                float acc = 0.f;
#pragma unroll 4
                for (int qk = 0; qk < cnt; ++qk) acc = fmaf(w[qk], pf[qk], acc);
                const int t = t0 + f;
                if (t < nfr) {
                    const float v = db10(acc);
                    S.dbm[(size_t)m * nfr + t] = v;
                    vmax = fmaxf(vmax, v);
                }
            }
            // ---- power rows -> scratch (chroma product after the tuning estimate) ----------------------
            for (int i = tid; i < F * NB; i += kCT) {
                const int f = i / NB, k = i - f * NB;
                if (t0 + f < nfr) S.P[(size_t)(t0 + f) * NB + k] = s_pow[f * G::PSTRIDE + k];
            }
            __syncthreads();                                   // the per-frame pass below rewrites its rows in place
            // ---- one warp per frame: every spectral scalar, piptrack, contrast, rms, zcr --------------
            for (int f = warp; f < F; f += kWarps) {
                const int t = t0 + f;
                if (t >= nfr) break;
                const float* pw = s_pow + f * G::PSTRIDE;
                float pv[CH], mg[CH];
                float sm = 0.f, sp = 0.f, sl = 0.f, mx = 0.f;
                double skm = 0.0;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int k = CH * lane + i;
                    const bool in = k < NB;
                    pv[i] = in ? pw[k] : 0.f;
                    mg[i] = sqrtf(pv[i]);
                    if (in) {
                        sm += mg[i];
                        skm += (double)k * (double)mg[i];
                        const float st = fmaxf(pv[i], 1e-10f);                // spectral_flatness: max(amin, S**2)
                        sp += st;
                        sl += logf(st);
                        mx = fmaxf(mx, pv[i]);
                    }
                }
                const float loc = sm;                                          // this lane's share of sum |S|
                float incl = loc;                                              // inclusive scan over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                const float total = __shfl_sync(0xffffffffu, incl, 31);
                const double dtotal = (double)total;
                skm = warp_sum_d(skm);
                sp = warp_sum(sp); sl = warp_sum(sl); mx = warp_max(mx);
                // spectral_centroid / bandwidth: util.normalize(S, norm=1) leaves a column below `tiny` undivided
                const double len = total < kTiny ? 1.0 : dtotal;
                const double cen = skm * bin_hz / len;
                double bw = 0.0;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int k = CH * lane + i;
                    if (k < NB) { const double d = (double)k * bin_hz - cen; bw += (double)mg[i] * d * d; }
                }
                bw = sqrt(warp_sum_d(bw) / len);
                // spectral_rolloff: first bin whose cumulative magnitude reaches 0.85 of the total
                int kro = 0x7fffffff;
                {
                    const float thr = 0.85f * total;
                    float cum = incl - loc;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int k = CH * lane + i;
                        if (k < NB) { cum += mg[i]; if (cum >= thr && kro == 0x7fffffff) kro = k; }
                    }
                    kro = warp_min_i(kro);
                    if (kro == 0x7fffffff) kro = NB - 1;
                }
                // piptrack (estimate_tuning): thresholded local maxima of the POWER spectrum, 150 Hz <= f < 4 kHz
                {
                    const float ref = 0.1f * mx;
#pragma unroll 1
                    for (int i = 0; i < CH; ++i) {
                        const int k = CH * lane + i;
                        if (k < p.pip_k0 || k >= p.pip_k1) continue;
                        const float s0 = pw[k], sl0 = pw[k - 1], sr0 = pw[k + 1];
                        const float x0 = s0 > ref ? s0 : 0.f, xl = sl0 > ref ? sl0 : 0.f, xr = sr0 > ref ? sr0 : 0.f;
                        if (!(x0 > xl && x0 >= xr)) continue;
                        const float a = __fsub_rn(__fadd_rn(sr0, sl0), __fmul_rn(2.f, s0));
                        const float bq = __fmul_rn(__fsub_rn(sr0, sl0), 0.5f);
                        const float shift = fabsf(bq) >= fabsf(a) ? 0.f : __fdiv_rn(-bq, a);
                        const float dskew = __fmul_rn(__fmul_rn(0.5f, bq), shift);        // avg == bq: central difference / 2
                        const float pitch = (float)(((double)k + (double)shift) * (double)p.sample_rate / (double)NFFT);
                        const int pos = atomicAdd(&s_misc[0], 1);
                        if (pos < p.cand_cap) { S.cp[pos] = pitch; S.cm[pos] = __fadd_rn(s0, dskew); }
                    }
                }
                // spectral_contrast: mean of the q smallest / largest magnitudes of every band, in dB.  The frame's row
                // now holds magnitudes, read lane-strided per band (ceil(cnt / 32) values a lane); a selection step is
                // a local scan, one redux.sync on the float bits (magnitudes are >= 0: bit order = value order), a
                // ballot to name the lane that owns the winner, which marks its entry as taken.  A band has more
                // than 2 q bins, so the entries the valley pass took never belong to the peak set.
                {
                    float* const row = s_pow + f * G::PSTRIDE;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int k = CH * lane + i;
                        if (k < NB) row[k] = mg[i];
                    }
                    __syncwarp();
                    constexpr uint32_t kTaken = 0xffffffffu;
                    float vm = 0.f, pm = 0.f;                                  // lane b keeps band b's means
#pragma unroll 1
                    for (int b = 0; b < kBands; ++b) {
                        const int bs = p.band_start[b], cnt = p.band_cnt[b], q = p.band_q[b];
                        uint32_t* const col = reinterpret_cast<uint32_t*>(row) + bs + lane;
                        const int nv = (cnt - lane + 31) >> 5;                 // this lane's entries: bs + lane + 32 i
                        double sv = 0.0, sq = 0.0;
#pragma unroll 1
                        for (int e = 0; e < q; ++e) {
                            uint32_t lm = kTaken;
                            for (int i = 0; i < nv; ++i) lm = min(lm, col[32 * i]);
                            const uint32_t gm = __reduce_min_sync(0xffffffffu, lm);
                            const int owner = __ffs(__ballot_sync(0xffffffffu, lm == gm)) - 1;
                            if (lane == owner)
                                for (int i = 0; i < nv; ++i) if (col[32 * i] == gm) { col[32 * i] = kTaken; break; }
                            sv += (double)__uint_as_float(gm);
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int e = 0; e < q; ++e) {
                            uint32_t lx = 0u;
                            for (int i = 0; i < nv; ++i) { const uint32_t x = col[32 * i]; lx = max(lx, x == kTaken ? 0u : x); }
                            const uint32_t gx = __reduce_max_sync(0xffffffffu, lx);
                            const int owner = __ffs(__ballot_sync(0xffffffffu, lx == gx)) - 1;
                            if (lane == owner)
                                for (int i = 0; i < nv; ++i) if (col[32 * i] == gx) { col[32 * i] = kTaken; break; }
                            sq += (double)__uint_as_float(gx);
                        }
                        __syncwarp();
                        const float vmean = (float)(sv / q), pmean = (float)(sq / q);   // np.mean of float32 values
                        if (lane == b) { vm = vmean; pm = pmean; }
                    }
                    if (lane < kBands) {                                        // power_to_db of float64 arrays (top_db at clip end)
                        S.vl[(size_t)lane * nfr + t] = (float)(10.0 * log10(fmax(1e-10, (double)vm)));
                        S.pk[(size_t)lane * nfr + t] = (float)(10.0 * log10(fmax(1e-10, (double)pm)));
                    }
                }
                // rms over the STFT frame (centre padding with zeros), zero crossings over the 2048 window (edge padding)
                float e2 = 0.f;
                {
                    const float* a = s_audio + f * hop;
                    for (int i = lane; i < NFFT; i += 32) e2 = fmaf(a[i], a[i], e2);
                    e2 = warp_sum(e2);
                }
                int zc = 0;
                {
                    const int start = t * hop - kZcrFrame / 2;
                    for (int i = lane; i < kZcrFrame - 1; i += 32) {
                        const int j1 = start + 1 + i;
                        const int ja = min(max(j1 - 1, 0), n - 1), jb = min(max(j1, 0), n - 1);
                        bool sa, sb;
                        if constexpr (I16) {
                            sa = ((const int16_t*)cptr)[ja] < 0; sb = ((const int16_t*)cptr)[jb] < 0;
                        } else {
                            sa = ((const float*)cptr)[ja] < -1e-10f; sb = ((const float*)cptr)[jb] < -1e-10f;
                        }
                        zc += sa != sb;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) zc += __shfl_xor_sync(0xffffffffu, zc, o);
                }
                if (lane == 0) {
                    S.sc[t] = (float)cen;
                    S.sc[(size_t)nfr + t] = (float)((double)kro * bin_hz);
                    S.sc[(size_t)2 * nfr + t] = (float)bw;
                    S.sc[(size_t)3 * nfr + t] = __fdiv_rn(expf(sl / NB), sp / NB);
                    S.sc[(size_t)4 * nfr + t] = (float)zc / (float)kZcrFrame;
                    S.sc[(size_t)5 * nfr + t] = sqrtf(e2 / NFFT);
                }
            }
            __syncthreads();
        }

        // ================================ per clip ======================================================
        vmax = warp_max(vmax);
        if (lane == 0) s_red[warp] = vmax;
        __syncthreads();
        {
            const float a = lane < kWarps ? s_red[lane] : -3.0e38f;
            vmax = warp_max(a);
        }
        const float floor_db = vmax - p.top_db;
        // ---- MFCC series: DCT-II of the clipped dB, four coefficients per thread ---------------------------
        {
            const int kq = (K + 3) / 4;
            for (int i = tid; i < kq * nfr; i += kCT) {
                const int kb = 4 * (i / nfr), t = i % nfr;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                for (int m = 0; m < n_mels; ++m) {
                    const float v = fmaxf(S.dbm[(size_t)m * nfr + t], floor_db);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = min(kb + e, K - 1);
                        acc[e] = fmaf(__ldg(p.dct + (size_t)k * n_mels + m), v, acc[e]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (kb + e < K) S.mf[(size_t)(kb + e) * nfr + t] = acc[e];
            }
        }
        // contrast: power_to_db's top_db clip is relative to the maximum over the whole (band, frame) array
        float pmax = -3.0e38f, qmax = -3.0e38f;
        for (int i = tid; i < kBands * nfr; i += kCT) { pmax = fmaxf(pmax, S.pk[i]); qmax = fmaxf(qmax, S.vl[i]); }
        pmax = warp_max(pmax); qmax = warp_max(qmax);
        __syncthreads();                                       // (s_red reads above are done; MFCC series visible below)
        if (lane == 0) { s_red[warp] = pmax; s_red[32 + warp] = qmax; }
        __syncthreads();
        {
            const float a = lane < kWarps ? s_red[lane] : -3.0e38f, b = lane < kWarps ? s_red[32 + lane] : -3.0e38f;
            pmax = warp_max(a); qmax = warp_max(b);
        }
        const float pfloor = pmax - 80.0f, qfloor = qmax - 80.0f;

        // ---- row statistics: 3 K MFCC-family rows, 6 scalars, 7 contrast rows ------------------------------
        {
            const int o_sc = 6 * K;                                            // centroid, rolloff, bandwidth
            for (int row = warp; row < 3 * K + 6 + kBands; row += kWarps) {
                if (row < K) {
                    const float* x = S.mf + (size_t)row * nfr;
                    row_stats(nfr, lane, [&](int t) { return x[t]; }, outv + row, outv + K + row);
                } else if (row < 3 * K) {
                    const int ord = row < 2 * K ? 1 : 2, k = row - ord * K;
                    const float* x = S.mf + (size_t)k * nfr;
                    auto d = [&](int t) {
                        const int c = min(max(t, 4), nfr - 5);                 // mode="interp": constant over the edge windows
                        double acc = 0.0;
#pragma unroll
                        for (int i = -4; i <= 4; ++i) {
                            const double w = ord == 1 ? (double)i / 60.0 : (3.0 * i * i - 20.0) / 462.0;
                            acc += w * (double)x[c + i];
                        }
                        return (float)acc;
                    };
                    row_stats(nfr, lane, d, outv + 2 * ord * K + k, outv + 2 * ord * K + K + k);
                } else if (row < 3 * K + 6) {
                    const int s = row - 3 * K;
                    // output order: centroid, rolloff, bandwidth, [contrast], flatness, [chroma], zcr, rms
                    const int off = s < 3 ? o_sc + 2 * s : (s == 3 ? o_sc + 6 + 14 : o_sc + 6 + 14 + 2 + 24 + 2 * (s - 4));
                    const float* x = S.sc + (size_t)s * nfr;
                    row_stats(nfr, lane, [&](int t) { return x[t]; }, outv + off, outv + off + 1);
                } else {
                    const int b = row - 3 * K - 6;
                    const float* pk = S.pk + (size_t)b * nfr;
                    const float* vl = S.vl + (size_t)b * nfr;
                    row_stats(nfr, lane, [&](int t) { return fmaxf(pk[t], pfloor) - fmaxf(vl[t], qfloor); },
                              outv + o_sc + 6 + b, outv + o_sc + 6 + kBands + b);
                }
            }
        }

        // ---- tuning estimate ---------------------------------------------------------------------------------
        const int M = min(s_misc[0], p.cand_cap);
        float med = 0.f;
        if (M > 0) {
            float sel[2];
            const int nsel = (M & 1) ? 1 : 2;
            for (int which = 0; which < nsel; ++which) {
                int rank = (M & 1) ? M / 2 : M / 2 - 1 + which;
                uint32_t prefix = 0, mask = 0;
                for (int pass = 3; pass >= 0; --pass) {
                    __syncthreads();
                    if (tid < 256) s_hist[tid] = 0;
                    __syncthreads();
                    for (int i = tid; i < M; i += kCT) {
                        const uint32_t key = fkey(S.cm[i]);
                        if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> (8 * pass)) & 255], 1);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        int cum = 0, bkt = 0;
                        for (; bkt < 255; ++bkt) { if (cum + s_hist[bkt] > rank) break; cum += s_hist[bkt]; }
                        s_misc[1] = bkt; s_misc[2] = rank - cum;
                    }
                    __syncthreads();
                    prefix |= (uint32_t)s_misc[1] << (8 * pass);
                    mask |= 0xffu << (8 * pass);
                    rank = s_misc[2];
                }
                sel[which] = fkey_inv(prefix);
            }
            med = nsel == 1 ? sel[0] : __fmul_rn(__fadd_rn(sel[0], sel[1]), 0.5f);
        }
        __syncthreads();
        if (tid < 128) s_hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < M; i += kCT) {
            if (S.cm[i] >= med) {
                // pitch_tuning: residual of 12 log2(f / 27.5) in float32, folded to [-0.5, 0.5), 100 bins of 0.01
                float r = fmodf(__fmul_rn(12.0f, log2f(__fdiv_rn(S.cp[i], 27.5f))), 1.0f);
                if (r >= 0.5f) r -= 1.0f;
                const double v = (double)r;
                int b = (int)floor((v + 0.5) * 100.0);
                b = min(max(b, 0), 99);
                while (b > 0 && v < __dadd_rn(__dmul_rn((double)b, 0.01), -0.5)) --b;
                while (b < 99 && v >= __dadd_rn(__dmul_rn((double)(b + 1), 0.01), -0.5)) ++b;
                atomicAdd(&s_hist[b], 1);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int best = 50, bc = -1;                                  // no candidates: tuning 0.0 = bank 50
            if (M > 0) { best = 0; for (int b = 0; b < 100; ++b) if (s_hist[b] > bc) { bc = s_hist[b]; best = b; } }
            s_misc[4] = best;
            if (p.tuning_out) p.tuning_out[clip] = (float)__dadd_rn(__dmul_rn((double)best, 0.01), -0.5);
        }
        __syncthreads();

        // ---- chroma with that tuning's filterbank, tonnetz; one warp per frame ----------------------------------
        {
            // the bank of this clip's tuning, staged over the (now idle) sample / exchange / power tiles
            float* const fb = reinterpret_cast<float*>(smem_raw);
            {
                const float* g = p.chroma + (size_t)s_misc[4] * 12 * NB;
                for (int i = tid; i < 12 * NB; i += kCT) fb[i] = __ldg(g + i);
            }
            __syncthreads();
            double a1 = 0.0, a2 = 0.0;                               // lane c < 12: chroma class c; 12..17: tonnetz dim
            for (int t = warp; t < nfr; t += kWarps) {
                const float* pr = S.P + (size_t)t * NB;
                float acc[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) acc[c] = 0.f;
                for (int k = lane; k < NB; k += 32) {
                    const float pwk = pr[k];
#pragma unroll
                    for (int c = 0; c < 12; ++c) acc[c] = fmaf(fb[c * NB + k], pwk, acc[c]);
                }
                float mxc = 0.f, l1 = 0.f;
#pragma unroll
                for (int c = 0; c < 12; ++c) { acc[c] = warp_sum(acc[c]); mxc = fmaxf(mxc, fabsf(acc[c])); }
                const float den = mxc < kTiny ? 1.f : mxc;
                float ch[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) { ch[c] = __fdiv_rn(acc[c], den); l1 += fabsf(ch[c]); }
                const float den1 = l1 < kTiny ? 1.f : l1;
                double val = 0.0;
                if (lane < 12) {
#pragma unroll
                    for (int c = 0; c < 12; ++c) if (lane == c) val = (double)ch[c];
                } else if (lane < 18) {
#pragma unroll
                    for (int c = 0; c < 12; ++c)
                        val += (double)__ldg(p.tonnetz + (lane - 12) * 12 + c) * (double)__fdiv_rn(ch[c], den1);
                }
                a1 += val; a2 += val * val;
            }
            if (lane < 18) { s_acc[(warp * 18 + lane) * 2] = a1; s_acc[(warp * 18 + lane) * 2 + 1] = a2; }
            __syncthreads();
            if (tid < 18) {
                double s1 = 0.0, s2 = 0.0;
                for (int w = 0; w < kWarps; ++w) { s1 += s_acc[(w * 18 + tid) * 2]; s2 += s_acc[(w * 18 + tid) * 2 + 1]; }
                const double mean = s1 / nfr, var = fmax(s2 / nfr - mean * mean, 0.0);
                const int o_sc = 6 * K;
                if (tid < 12) { outv[o_sc + 22 + tid] = (float)mean; outv[o_sc + 34 + tid] = (float)sqrt(var); }
                else { outv[o_sc + 50 + (tid - 12)] = (float)mean; outv[o_sc + 56 + (tid - 12)] = (float)sqrt(var); }
            }
        }
        __syncthreads();
    }
}

template <int LOG2NC>
size_t smem_bytes_t(int hop, int n_mels, int mel_nnz) {
    using G = FftGeom<LOG2NC>;
    using C = ClsCfg<LOG2NC>;
    const int cl = (hop * (C::F - 1) + G::NFFT + 7) & ~7;
    size_t b = ((size_t)cl * 4 + (size_t)C::FR * G::XSTRIDE * 8 + (size_t)C::F * G::PSTRIDE * 4 + 15) & ~(size_t)15;
    b += (size_t)G::NC * 8 + (size_t)(G::NC / 2 + 1) * 8 + (size_t)FftTwp<LOG2NC>::SIZE * 8;
    b += (size_t)n_mels * 12 + (size_t)mel_nnz * 4;
    b += 64 * 4 + 256 * 4 + 8 * 4 + 8;
    b = (b + 7) & ~(size_t)7;
    b += (size_t)kWarps * 18 * 2 * 8;
    return b + 64;
}

template <int LOG2NC, bool I16>
cudaError_t launch_t(const ClassicalParams& p, int grid, cudaStream_t st) {
    const size_t smem = smem_bytes_t<LOG2NC>(p.hop, p.n_mels, p.mel_nnz);
    auto k = classical_kernel<LOG2NC, I16>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kCT, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

size_t classical_smem_bytes(int log2nc, int hop, int n_mels, int mel_nnz) {
    switch (log2nc) {
        case 8: return smem_bytes_t<8>(hop, n_mels, mel_nnz);
        case 9: return smem_bytes_t<9>(hop, n_mels, mel_nnz);
        case 10: return smem_bytes_t<10>(hop, n_mels, mel_nnz);
        default: return (size_t)-1;
    }
}

size_t classical_scratch_floats(int n_fft, int n_frames, int n_mels, int n_mfcc, int cand_cap) {
    const size_t nb = (size_t)n_fft / 2 + 1;
    size_t f = (size_t)n_frames * nb + (size_t)n_mels * n_frames + (size_t)n_mfcc * n_frames +
               (size_t)(6 + 2 * kBands) * n_frames + 2 * (size_t)cand_cap;
    return (f + 3) & ~(size_t)3;
}

cudaError_t launch_classical(const ClassicalParams& p, int log2nc, bool i16, int grid, cudaStream_t st) {
    switch (log2nc) {
        case 8: return i16 ? launch_t<8, true>(p, grid, st) : launch_t<8, false>(p, grid, st);
        case 9: return i16 ? launch_t<9, true>(p, grid, st) : launch_t<9, false>(p, grid, st);
        case 10: return i16 ? launch_t<10, true>(p, grid, st) : launch_t<10, false>(p, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b2a
