// frontend.cu — fused STFT -> power -> mel -> dB -> normalise kernels (audio_mel_spec,
// audio_mfcc_seq).  sm_100a only.
//
// One persistent CTA owns one clip at a time (clips are independent; the only reductions —
// ref=np.max, min/max, per-coefficient mean/std — are per clip: deep.py:64-67,133,326-328).
// Per tile of F frames:
//   1. stage the samples the tile touches into shared memory ONCE as fp32 (int16 -> x/32768
//      exactly as librosa.load does; frames overlap n_fft/hop times so this amortises the
//      conversion), zeros (or the reflected sample) outside [0, n_samples): librosa.stft
//      centre padding, deep.py:126-132 via melspectrogram;
//   2. per frame: window x packed real FFT (fft_core.cuh) -> |X|^2 into a [frame][bin] tile;
//   3. banded mel dot products (lane = frame, so filter weights are warp-uniform) -> 10 log10;
//      raw dB goes to global memory (L2-resident, re-read by the same CTA a few us later) and
//      the running per-clip max / min stay in registers.
// Then per clip: power_to_db(ref=max, top_db) + min-max (mel) or top_db clip + DCT-II +
// per-row z-score (mfcc), rewriting the L2-hot tile in place.
#include "frontend.h"
#include "fft_core.cuh"
#include "front_stage.cuh"

#include <cstdint>

namespace b2a {

template <int LOG2NC> struct FrontCfg {
    using G = FftGeom<LOG2NC>;
    static constexpr int F = (LOG2NC <= 8) ? 32 : 16;                 // frames per tile
    static constexpr int FR = kThreads / G::T;                        // frames per FFT round
    static constexpr int ROUNDS = (F + FR - 1) / FR;
    static constexpr int MINB = (LOG2NC <= 9) ? 2 : 1;
};

size_t front_smem_bytes(int log2nc, int hop, int n_mels, int mel_nnz) {
    const int NC = 1 << log2nc, n_fft = 2 * NC, T = NC / 16;
    const int F = (log2nc <= 8) ? 32 : 16;
    const int FR = kThreads / T;
    const int slots = FR < F ? FR : F;
    size_t cl = (size_t)hop * (F - 1) + n_fft;
    cl = (cl + 7) & ~(size_t)7;
    size_t b = 0;
    b += cl * 4;                                        // audio
    b += (size_t)slots * (NC + NC / 16) * 8;            // exchange
    b += (size_t)F * (NC + 1) * 4;                      // power tile
    b = (b + 15) & ~(size_t)15;
    b += (size_t)NC * 8;                                // tw
    b += (size_t)(NC / 2 + 1) * 8;                      // tw2
    b += (size_t)(16 / (log2nc >= 8 ? 16 : (1 << (log2nc - 4)))) * ((log2nc >= 8 ? 16 : (1 << (log2nc - 4))) - 1) * T * 8;   // per-thread pass-2 twiddles
    b += (size_t)n_mels * 3 * 4 + (size_t)mel_nnz * 4;  // banded mel
    b += 128 * 4;                                       // reduction scratch
    return b + 64;
}

// DCT-II rows [k0, k0 + KB) of one 32-frame tile of clipped dB, lane = frame.  Register-blocked
// KB coefficients x 4 bands: per step KB warp-uniform 128-bit basis loads (one wavefront each), four
// conflict-free dB loads and 4 KB FMAs — the (coefficient, frame) loop it replaces issued two loads
// per FMA and was 3/4 of the reference-default mfcc's time.  Bands are accumulated in ascending
// order, one chain per coefficient: bit-identical to the scalar loop.  KB = 5: the reference default
// of 40 coefficients is one block per warp (eight instantiations for other counts made the FFT loop spill).
template <int KB>
__device__ __forceinline__ void dct_tile_rows(const float* __restrict__ dct, int n_mels, int n_mfcc, int k0,
                                              const float* __restrict__ s_lf, float* __restrict__ out_t,
                                              int nfr, bool valid) {
    const float4* d4[KB];
    float acc[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        const int k = k0 + i < n_mfcc ? k0 + i : n_mfcc - 1;
        d4[i] = reinterpret_cast<const float4*>(dct + (size_t)k * n_mels);
        acc[i] = 0.f;
    }
#pragma unroll 2
    for (int m4 = 0; m4 < n_mels / 4; ++m4) {
        const float v0 = s_lf[(4 * m4) * 32], v1 = s_lf[(4 * m4 + 1) * 32];
        const float v2 = s_lf[(4 * m4 + 2) * 32], v3 = s_lf[(4 * m4 + 3) * 32];
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            const float4 w = __ldg(d4[i] + m4);
            acc[i] = fmaf(w.x, v0, acc[i]);
            acc[i] = fmaf(w.y, v1, acc[i]);
            acc[i] = fmaf(w.z, v2, acc[i]);
            acc[i] = fmaf(w.w, v3, acc[i]);
        }
    }
    if (valid) {
#pragma unroll
        for (int i = 0; i < KB; ++i)
            if (k0 + i < n_mfcc) out_t[(size_t)(k0 + i) * nfr] = acc[i];
    }
}

template <int LOG2NC, bool I16, int KIND, bool RAG>
__global__ void __launch_bounds__(kThreads, FrontCfg<LOG2NC>::MINB) front_kernel(FrontParams p) {
    using G = FftGeom<LOG2NC>;
    using C = FrontCfg<LOG2NC>;
    constexpr int NC = G::NC, NFFT = G::NFFT, T = G::T, F = C::F, FR = C::FR;
    constexpr int SLOTS = FR;
    static_assert(FR <= F && F % FR == 0, "tile must be a whole number of FFT rounds");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int cl = (p.hop * (F - 1) + NFFT + 7) & ~7;
    float* s_audio = reinterpret_cast<float*>(smem_raw);
    float2* s_xch = reinterpret_cast<float2*>(s_audio + cl);
    float* s_pow = reinterpret_cast<float*>(s_xch + SLOTS * G::XSTRIDE);
    // byte offsets from smem_raw (no integer round trip: keeps the shared address space visible
    // to the compiler, so table reads are LDS, not generic LD)
    const int off_tw = ((cl * 4 + SLOTS * G::XSTRIDE * 8 + F * G::PSTRIDE * 4) + 15) & ~15;
    float2* s_tw = reinterpret_cast<float2*>(smem_raw + off_tw);
    float2* s_tw2 = s_tw + NC;
    float2* s_twp = s_tw2 + NC / 2 + 1;
    int* s_k0 = reinterpret_cast<int*>(s_twp + FftTwp<LOG2NC>::SIZE);
    int* s_cnt = s_k0 + p.n_mels;
    int* s_off = s_cnt + p.n_mels;
    float* s_w = reinterpret_cast<float*>(s_off + p.n_mels);
    float* s_red = s_w + p.mel_nnz;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-CTA tables ---------------------------------------------------------------------
    for (int i = tid; i < NC; i += kThreads) s_tw[i] = p.tw[i];
    for (int i = tid; i < NC / 2 + 1; i += kThreads) s_tw2[i] = p.tw2[i];
    FftTwp<LOG2NC>::fill(s_twp, p.tw, tid, kThreads);
    for (int i = tid; i < p.n_mels; i += kThreads) { s_k0[i] = p.mel_k0[i]; s_cnt[i] = p.mel_cnt[i]; s_off[i] = p.mel_off[i]; }
    for (int i = tid; i < p.mel_nnz; i += kThreads) s_w[i] = p.mel_w[i];

    // ---- per-thread constants (loop invariant over frames and clips) -----------------------
    const int j = tid % T;                 // position inside the frame's thread group
    const int slot = tid / T;              // frame slot inside an FFT round
    float2 win[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = j + T * t;
        win[t] = make_float2(__ldg(p.window + 2 * n), __ldg(p.window + 2 * n + 1));
    }
    constexpr bool HOIST = (LOG2NC == 8);
    float2 tw1[HOIST ? 15 : 1];
    if constexpr (HOIST) {
#pragma unroll
        for (int t = 1; t < 16; ++t) tw1[t - 1] = p.tw[t * j];
    }
    __syncthreads();

    int n = p.n_samples, nfr = p.n_frames;                  // per clip when RAG
    const int n_mels = p.n_mels;
    const bool hop_even = (p.hop & 1) == 0;
    const size_t esz = I16 ? 2 : 4;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(p.clips) & 15) == 0;

    for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
        if constexpr (RAG) { n = p.rag_len[clip]; nfr = 1 + n / p.hop; }
        const long long clip_elem0 = RAG ? p.rag_in_off[clip] : clip * (long long)n;
        const void* cptr = (const unsigned char*)p.clips + (size_t)clip_elem0 * esz;
        float* const outb = RAG ? p.out + p.rag_out_off[clip]
                                : p.out + (size_t)clip * (KIND == 0 ? n_mels : p.n_mfcc) * nfr;
        float* inter = (KIND == 0) ? outb : p.inter + (size_t)blockIdx.x * n_mels * p.n_frames;
        float vmax = -3.0e38f, vmin = 3.0e38f;

        for (int t0 = 0; t0 < nfr; t0 += F) {
            // (1) stage audio for frames [t0, t0+F)
            stage_audio<I16>(s_audio, cptr, clip_elem0, t0 * p.hop - NFFT / 2, cl, n, p.pad_mode, base_aligned);
            __syncthreads();

            // (2) FFT rounds
#pragma unroll 1
            for (int r = 0; r < C::ROUNDS; ++r) {
                const int f = r * FR + slot;               // frame inside the tile
                float2* xb = s_xch + slot * G::XSTRIDE;
                {
                    float2 v[16];
                    const float* a = s_audio + f * p.hop + 2 * j;
                    if (hop_even) {
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            const float2 x = *reinterpret_cast<const float2*>(a + 2 * T * t);
                            v[t] = make_float2(x.x * win[t].x, x.y * win[t].y);
                        }
                    } else {
#pragma unroll
                        for (int t = 0; t < 16; ++t)
                            v[t] = make_float2(a[2 * T * t] * win[t].x, a[2 * T * t + 1] * win[t].y);
                    }
                    Dft<16>::run(v);
#pragma unroll
                    for (int t = 0; t < 16; ++t) xb[xpad(16 * j + t)] = v[t];
                }
                frame_sync<T>();
                fft_tail_passes<LOG2NC, HOIST>(xb, s_tw, tw1, s_twp, j);
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
                        pw[NC / 2] = A.x * A.x + A.y * A.y;      // X[NC/2] = conj(Z[NC/2])
                    }
                }
                frame_sync<T>();
            }
            __syncthreads();

            // (3) mel bands for the tile: item = (band, frame), frame fastest
            for (int i = tid; i < n_mels * F; i += kThreads) {
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
                    inter[(size_t)m * nfr + t] = v;
                    vmax = fmaxf(vmax, v);
                    vmin = fminf(vmin, v);
                }
            }
            // no barrier needed here: the next stage_audio only touches s_audio, and every thread
            // passes the barrier after it only once its own mel items are done.
        }

        // ---- per-clip reductions ---------------------------------------------------------------
        vmax = warp_max(vmax); vmin = warp_min(vmin);
        if (lane == 0) { s_red[warp] = vmax; s_red[32 + warp] = vmin; }
        __syncthreads();      // also makes the tile's global writes visible to the whole CTA
        {
            float a = (lane < kThreads / 32) ? s_red[lane] : -3.0e38f;
            float b = (lane < kThreads / 32) ? s_red[32 + lane] : 3.0e38f;
            vmax = warp_max(a); vmin = warp_min(b);
        }

        if constexpr (KIND == 0) {
            // power_to_db(ref=np.max, top_db) then _normalize  (deep.py:133-134, :64-67)
            const float lo = fmaxf(vmin - vmax, -p.top_db);
            const float range = (0.0f - lo) + 1e-8f;
            const int total = n_mels * nfr;
            if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(inter) & 15) == 0)) {
                float4* o4 = reinterpret_cast<float4*>(inter);
                for (int i = tid; i < total / 4; i += kThreads) {
                    float4 v = o4[i];
                    v.x = __fdiv_rn(fmaxf(v.x - vmax, -p.top_db) - lo, range);
                    v.y = __fdiv_rn(fmaxf(v.y - vmax, -p.top_db) - lo, range);
                    v.z = __fdiv_rn(fmaxf(v.z - vmax, -p.top_db) - lo, range);
                    v.w = __fdiv_rn(fmaxf(v.w - vmax, -p.top_db) - lo, range);
                    o4[i] = v;
                }
            } else {
                for (int i = tid; i < total; i += kThreads)
                    inter[i] = __fdiv_rn(fmaxf(inter[i] - vmax, -p.top_db) - lo, range);
            }
        } else {
            // librosa.feature.mfcc: power_to_db(ref=1.0, top_db) -> DCT-II ortho -> rows [0, n_mfcc)
            // then deep.py:326-328 per-row z-score.
            float* outc = outb;
            const float thr = vmax - p.top_db;
            float* s_l = s_pow;                       // [n_mels][32] tile of clipped dB
            for (int t0 = 0; t0 < nfr; t0 += 32) {
                __syncthreads();
                // raw dB come back from L2: eight loads in flight per thread, not one round trip each
#pragma unroll 1
                for (int i0 = tid; i0 < n_mels * 32; i0 += 8 * kThreads) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kThreads, m = i >> 5, t = t0 + (i & 31);
                        v[u] = (i < n_mels * 32 && t < nfr) ? inter[(size_t)m * nfr + t] : -3.0e38f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kThreads;
                        if (i < n_mels * 32) s_l[i] = (t0 + (i & 31) < nfr) ? fmaxf(v[u], thr) : 0.f;
                    }
                }
                __syncthreads();
                constexpr int KB = 5;                                                  // coefficients per warp and block
                if ((n_mels & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dct) & 15) == 0) {
                    const int t = t0 + lane;
                    for (int k0 = warp * KB; k0 < p.n_mfcc; k0 += (kThreads / 32) * KB)
                        dct_tile_rows<KB>(p.dct, n_mels, p.n_mfcc, k0, s_l + lane, outc + t, nfr, t < nfr);
                } else {
                    for (int i = tid; i < p.n_mfcc * 32; i += kThreads) {
                        const int k = i >> 5, f = i & 31, t = t0 + f;
                        const float* d = p.dct + (size_t)k * n_mels;
                        float acc = 0.f;
#pragma unroll 4
                        for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(d + m), s_l[m * 32 + f], acc);
                        if (t < nfr) outc[(size_t)k * nfr + t] = acc;
                    }
                }
            }
            __syncthreads();
            const float fn = (float)nfr;
            for (int k = warp; k < p.n_mfcc; k += kThreads / 32) {
                float* row = outc + (size_t)k * nfr;
                // mean about the first sample: exact for a constant row (silence -> z == 0, as the
                // reference gives) and better conditioned otherwise
                const float x0 = row[0];
                float s = 0.f;
                for (int t = lane; t < nfr; t += 32) s += row[t] - x0;
                const float mean = x0 + __fdiv_rn(warp_sum(s), fn);
                float ss = 0.f;
                for (int t = lane; t < nfr; t += 32) { const float d = row[t] - mean; ss = fmaf(d, d, ss); }
                const float sd = sqrtf(__fdiv_rn(warp_sum(ss), fn)) + 1e-8f;
                for (int t = lane; t < nfr; t += 32) row[t] = __fdiv_rn(row[t] - mean, sd);
            }
        }
        __syncthreads();     // s_red / s_pow reuse by the next clip
    }
}

template <int LOG2NC, bool I16, int KIND, bool RAG>
static cudaError_t launch_one(const FrontParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = front_kernel<LOG2NC, I16, KIND, RAG>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // persistent CTAs: never launch more than fit at once (shared memory may allow fewer than MINB)
    int occ = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kThreads, smem) == cudaSuccess && occ > 0 &&
        cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        grid = grid < occ * sms ? grid : occ * sms;
    k<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int LOG2NC, bool RAG>
static cudaError_t launch_r(const FrontParams& p, bool i16, int kind, int grid, size_t smem, cudaStream_t st) {
    if (kind == 0) return i16 ? launch_one<LOG2NC, true, 0, RAG>(p, grid, smem, st) : launch_one<LOG2NC, false, 0, RAG>(p, grid, smem, st);
    return i16 ? launch_one<LOG2NC, true, 1, RAG>(p, grid, smem, st) : launch_one<LOG2NC, false, 1, RAG>(p, grid, smem, st);
}

template <int LOG2NC>
static cudaError_t launch_l(const FrontParams& p, bool i16, int kind, int grid, size_t smem, cudaStream_t st) {
    return p.rag_len ? launch_r<LOG2NC, true>(p, i16, kind, grid, smem, st)
                     : launch_r<LOG2NC, false>(p, i16, kind, grid, smem, st);
}

int front_ctas_per_sm(int log2nc) { return log2nc <= 9 ? 2 : 1; }

cudaError_t launch_front(const FrontParams& p, int log2nc, bool i16, int kind, int grid, cudaStream_t st) {
    const size_t smem = front_smem_bytes(log2nc, p.hop, p.n_mels, p.mel_nnz);
    switch (log2nc) {
        case 7: return launch_l<7>(p, i16, kind, grid, smem, st);
        case 8: return launch_l<8>(p, i16, kind, grid, smem, st);
        case 9: return launch_l<9>(p, i16, kind, grid, smem, st);
        case 10: return launch_l<10>(p, i16, kind, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace b2a
