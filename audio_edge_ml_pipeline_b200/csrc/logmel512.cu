// logmel512.cu — the headline kernel: n_fft = 512 fused log-mel / MFCC front end, sm_100a only.
//
// One persistent, warp-specialised CTA per SM (640 threads):
//   * 16 FFT warps (setmaxnreg 104).  One half-warp per frame, 16 complex points per lane in
//     registers for both radix-16 passes of the 256-point packed FFT, all complex arithmetic in
//     packed FP32 (FADD2 / FMUL2 / FFMA2, fft_core.cuh).  Pass 1 reads the packed int16 pairs
//     straight from the raw PCM tile (one 32-bit word = one complex point) and widens them in
//     registers, the exact 1/32768 riding on the window; pass-2 results never leave registers —
//     only the 8 rows the mirror lane needs go through shared memory (X[k] needs Z[k] and
//     Z[256-k], which lives in lane 16-j).  4|X|^2 goes to a [bin pair][frame] power tile.
//     Raw words, mirror operands and split twiddles are issued as explicit batches of loads.
//     These warps never touch global memory and never meet a CTA-wide barrier.
//   * 4 mel warps (setmaxnreg 64).  Warp 0 is also the TMA producer: one cp.async.bulk per
//     32-frame tile of raw PCM into a ring (4 slots for int16 clips), issued a ring ahead (also
//     across clip boundaries).  All four consume power tiles (2-deep ring): lane = frame, band
//     warp-uniform; for the headline configuration the band sweep is generated at build time
//     (gen_mel.cpp) with the 490 filter weights as FFMA immediates and every power pair loaded
//     once.  Raw dB goes to the output buffer (L2-resident), per-clip max/min stay in registers,
//     and a clip is normalised in place during the NEXT clip's tiles, a prefetched slice per tile
//     (mfcc, headline shape: the DCT-II of the whole clip runs once at clip end from the raw-dB
//     scratch in L2, one frame per thread through cp.async-staged columns, basis folded on its
//     symmetry and generated as FFMA immediates, float64 z-score sums on the way; other mfcc
//     shapes: in-tile table-driven DCT, recomputed from the scratch only when the top_db clip
//     engages.  Either way the rows are z-scored by the same deferred pass).
//   * Hand-off by mbarriers only: raw_full (TMA tx bytes) -> FFT; pow_full (16 warp arrivals)
//     -> mel; pow_empty (4 warp arrivals) -> FFT.  pow_full of tile i also tells the producer
//     that the raw slot of tile i is free again.
// The FFT warps are bound by the FMA and shared-memory pipes; mel, dB, normalisation, staging and
// every global access overlap them instead of alternating with them (the phase-alternating
// predecessor spent 58 % of its time in the FFT rounds).
//
// Reference arithmetic: deep.py:126-134 (mel), :318-328 (mfcc) via librosa 0.11.0.
#include "frontend.h"
#include "fft_core.cuh"
#include "ws_common.cuh"
#include "gen/mel_special.inc"   // build-time generated straight-line mel code (gen_mel.cpp)

// arithmetic of the generated per-clip DCT (KIND 2): float64 accumulation of the 20 folded terms
#ifndef B2A_AB_DCT64
#define B2A_AB_DCT64 0
#endif
#ifndef B2A_AB_LOG2F
#define B2A_AB_LOG2F 0
#endif
#if B2A_AB_DCT64
#define B2A_DCT_T double
#define B2A_DCT_CVT(x) ((double)(x))
#define B2A_DCT_FMA(cf, cd, f, a) fma((cd), (f), (a))
#else
#define B2A_DCT_T float
#define B2A_DCT_CVT(x) (x)
#define B2A_DCT_FMA(cf, cd, f, a) fmaf((cf), (f), (a))
#endif

#include <cstdint>
#include <type_traits>

namespace b2a {

#ifdef B2A_TRACE   // debugging aid (tools/trace_ws.py): clock64 stamps of CTA 0's hand-offs, never in a product build
__device__ long long g_trace[8 * 512];
#define B2A_STAMP(SLOT, K) do { if (blockIdx.x == 0 && lane == 0 && (K) < 512u) g_trace[(SLOT) * 512 + (K)] = clock64(); } while (0)
#else
#define B2A_STAMP(SLOT, K) do { } while (0)
#endif

namespace {

using namespace ws;

constexpr int kFftWarps = 16, kMelWarps = 4;
constexpr int kThreads = 32 * (kFftWarps + kMelWarps);
constexpr int kMelThreads = 32 * kMelWarps;
constexpr int kFftRegs = 104, kMelRegs = 64;    // 512*112 + 128*32 = 640*96 (the launch allocation)
constexpr int NC = 256, NFFT = 512, F = 32;     // complex points, frame length, frames per tile
constexpr int ROUNDS = F / (2 * kFftWarps);     // rounds of 2 frames per FFT warp per tile
// ring depth of raw PCM tiles: 4 for int16; float32 input doubles a slot, so 3 (mel) / 2 (mfcc, which
// also carries the dB tile, the DCT basis and the partial sums) keep the CTA inside 227 KB
__host__ __device__ constexpr int nraw(bool i16, bool mfcc) { return i16 ? 4 : (mfcc ? 2 : 3); }
constexpr int kMaxRaw = 4;
// ring depth of power tiles: 3 absorbs the mel warps' per-clip normalisation pause; float32 input
// doubles the raw ring, leaving room for 2
__host__ __device__ constexpr int npow(bool i16) { return i16 ? 2 : 2; }
constexpr int XS = 17;                          // exchange row stride (float2): 64-bit pass-1 stores (row j) and
constexpr int XSLOT = 16 * XS + 2;              // 64-bit pass-2 loads (column j) are both conflict-free
constexpr int PROW = 68;                        // power tile: row = 2 adjacent bins x (32 frames + 2 pad)
constexpr int PROWS = 132;                      // bin pairs (0,1)..(256,257) + 3 zero rows for 8-bin padding
constexpr int kZFast = 16;                      // mfcc: up to this many coefficients take the one-sweep z-score
static_assert(ROUNDS >= 1 && ROUNDS * 2 * kFftWarps == F, "tile must be a whole number of rounds");
static_assert(B2A_MELSPEC_WARPS == kMelWarps, "regenerate gen/mel_special.inc for this warp count");
static_assert(B2A_DCTSPEC_NMFCC <= kZFast && kMelWarps == 4, "generated DCT: one-sweep z-score, four mel warps");
// KIND 2 overlays [n_mels][128] float columns + [4][n_mfcc][2] doubles on the dB tile, partial-sum and basis regions
static_assert(B2A_MELSPEC_NMELS * 32 * 4 + (2 * 4 * B2A_DCTSPEC_NMFCC * 32 * 4 + 512) +
                      B2A_MELSPEC_NMELS * 4 * (4 * ((((B2A_DCTSPEC_NMFCC + 3) / 4) + 3) / 4)) * 4 >=
                  B2A_MELSPEC_NMELS * 128 * 4 + 4 * B2A_DCTSPEC_NMFCC * 16,
              "generated DCT: the per-thread dB columns do not fit the regions they overlay");

// barrier among the mel warps only (named barrier 1); the FFT warps never stop for it
__device__ __forceinline__ void mel_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kMelThreads) : "memory");
}

struct Layout {
    int chunk;          // samples staged per tile, multiple of 8
    int raw_bytes;      // bytes of one raw slot (multiple of 16)
    int gp;             // mfcc: DCT coefficients per mel warp, padded to a multiple of 4
    int off_raw, off_xch, off_pow, off_tw2, off_melw, off_melk, off_red, off_bar, off_db, off_dct, off_part, off_zq, total;
};

__host__ __device__ inline Layout make_layout(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc) {
    const bool mfcc = n_mfcc > 0;
    Layout L;
    L.chunk = (hop * (F - 1) + NFFT + 7) & ~7;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.raw_bytes = (L.chunk * (i16 ? 2 : 4) + 15) & ~15;
    L.off_raw = take(nraw(i16, mfcc) * L.raw_bytes);  // ring of raw PCM tiles, read directly by pass 1
    L.off_xch = take(2 * kFftWarps * XSLOT * 8);      // one exchange slot per half-warp
    L.off_pow = take(npow(i16) * PROWS * PROW * 4);   // ring of power tiles
    L.off_tw2 = take(8 * 16 * 8);
    L.off_melw = take(mel_wpad * 4);
    L.off_melk = take(n_mels * 16);
    L.off_red = take((64 + 2 * kZFast) * 4);          // per-warp max/min, then (mean, sd) per coefficient
    L.off_bar = take((kMaxRaw + 2 * 3) * 8);
    L.off_db = take(mfcc ? n_mels * 32 * 4 : 0);      // mfcc: [n_mels][32] dB tile feeding the in-tile DCT
    L.off_part = take(mfcc && n_mfcc <= kZFast ? 2 * kMelWarps * n_mfcc * 32 * 4 + 512 : 0);   // reserve: with off_db and off_dct it holds
                                                                                             // KIND 2's [n_mels][128] columns + per-warp sums
    L.gp = 4 * ((((n_mfcc + kMelWarps - 1) / kMelWarps) + 3) / 4);
    L.off_dct = take(mfcc ? n_mels * kMelWarps * L.gp * 4 : 0);   // mfcc: DCT-II basis as [mel][mel warp][gp]
    // headline mfcc shape: [2 * odd coefficients][128 threads] float64 running sums of the per-clip DCT phase
    L.off_zq = take(mfcc && n_mfcc == B2A_DCTSPEC_NMFCC && n_mels == B2A_MELSPEC_NMELS
                        ? 2 * B2A_DCT_GROUP1_NK * kMelThreads * 8 : 0);
    L.total = o;
    return L;
}

template <bool I16, int KIND, bool SPEC, bool RAG>
__global__ void __launch_bounds__(kThreads, 1) logmel512_kernel(FrontParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool MFCC = KIND >= 1;          // KIND 2: mfcc with the generated DCT (headline bands, 13 coefficients)
    const Layout L = make_layout(p.hop, p.n_mels, p.mel_wpad, I16, MFCC ? p.n_mfcc : 0);
    float* const s_db = reinterpret_cast<float*>(smem + L.off_db);
    float2* const s_xch = reinterpret_cast<float2*>(smem + L.off_xch);
    float* const s_pow = reinterpret_cast<float*>(smem + L.off_pow);
    float2* const s_tw2 = reinterpret_cast<float2*>(smem + L.off_tw2);
    float* const s_melw = reinterpret_cast<float*>(smem + L.off_melw);
    int4* const s_desc = reinterpret_cast<int4*>(smem + L.off_melk);
    float* const s_red = reinterpret_cast<float*>(smem + L.off_red);
    float* const s_zs = s_red + 64;
    float* const s_dct = reinterpret_cast<float*>(smem + L.off_dct);
    float* const s_part = reinterpret_cast<float*>(smem + L.off_part);
    uint64_t* const bar_raw_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* const bar_pow_full = bar_raw_full + kMaxRaw;
    uint64_t* const bar_pow_empty = bar_pow_full + 3;

    constexpr int NPOW = npow(I16);
    constexpr int NRAW = nraw(I16, KIND >= 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hop = p.hop, n_mels = p.n_mels, chunk = L.chunk;
    using E = typename std::conditional<I16, int16_t, float>::type;

    // ---- per-CTA tables --------------------------------------------------------------------
    for (int i = tid; i < 128; i += kThreads) {          // s_tw2[r/2][j][r&1] = exp(-i pi (j+16r)/256): the split
        const int r = i >> 4, jj = i & 15;               // step reads two twiddles per conflict-free 128-bit load
        s_tw2[(r >> 1) * 32 + jj * 2 + (r & 1)] = p.tw2[jj + 16 * r];
    }
    for (int i = tid; i < p.mel_wpad; i += kThreads) s_melw[i] = p.mel_wq[i];
    for (int i = tid; i < n_mels; i += kThreads) {
        const int m = p.mel_order[i];                        // position i is served by mel warp i % kMelWarps
        s_desc[i] = make_int4((p.mel_k0e[m] >> 1) * (PROW / 2), p.mel_cnt4[m], p.mel_off4[m], m);
    }
    if constexpr (MFCC) {
        // DCT-II basis re-laid out so that mel warp w finds its coefficients k = w + 4 g, g < gp, as
        // consecutive floats per mel band (128-bit broadcast loads in the in-tile DCT)
        for (int i = tid; i < n_mels * kMelWarps * L.gp; i += kThreads) {
            const int g = i % L.gp, w = (i / L.gp) % kMelWarps, m = i / (L.gp * kMelWarps);
            const int k = w + kMelWarps * g;
            s_dct[i] = k < p.n_mfcc ? p.dct[(size_t)k * n_mels + m] : 0.f;
        }
    }
    for (int b = 0; b < NPOW; ++b)                           // bins 256..263 of every power tile
        for (int i = tid; i < 4 * PROW; i += kThreads) s_pow[b * PROWS * PROW + 128 * PROW + i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) mbar_init(bar_raw_full + i, 1);
        for (int i = 0; i < NPOW; ++i) { mbar_init(bar_pow_full + i, kFftWarps); mbar_init(bar_pow_empty + i, kMelWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (blockIdx.x >= p.n_clips) return;

    if (warp < kFftWarps) {
        // =========================== FFT warps ===================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFftRegs));
        const int j = lane & 15, h = lane >> 4;
        float2 win[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int q = j + 16 * t;
            // int16 path: the exact power-of-two 1/32768 of librosa.load rides on the window
            const float sc = I16 ? (1.0f / 32768.0f) : 1.0f;
            win[t] = make_float2(__ldg(p.window + 2 * q) * sc, __ldg(p.window + 2 * q + 1) * sc);
        }
        float2 tw1[15];
#pragma unroll
        for (int t = 1; t < 16; ++t) tw1[t - 1] = p.tw[t * j];
        float2* const xs = s_xch + (2 * warp + h) * XSLOT;       // this frame's exchange slot
        float2* const x1 = xs + XS * j;                          // pass-1 store base (row j)
        float2* const x2 = xs + j;                               // pass-2 load base (column j, stride XS)
        float2* const mst = xs + j;                              // mirror store base: M[(row-8)*16 + j]
        const float2* const mld = xs + (j ? 16 - j : 16);        // mirror load base:  M[(7-r)*16 + ...]
        const float4* const t2 = reinterpret_cast<const float4*>(s_tw2) + j;

        uint32_t it = 0;                                         // tiles this CTA has processed
        for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
            const int nfr = RAG ? 1 + p.rag_len[clip] / hop : p.n_frames;
            const int tiles = (nfr + F - 1) / F;
            for (int tile = 0; tile < tiles; ++tile, ++it) {
                const int t0 = tile * F;
                const uint32_t rb = it % NRAW, pb = it % NPOW;
                if (warp == 0) B2A_STAMP(1, it);
                if (warp == kFftWarps - 1) B2A_STAMP(3, it);
                mbar_wait(bar_raw_full + rb, (it / NRAW) & 1);           // this tile's samples have landed
                if (warp == 0) B2A_STAMP(2, it);
                if (warp == kFftWarps - 1) B2A_STAMP(4, it);
                bool pow_free = false;
                const E* const cur = reinterpret_cast<const E*>(smem + L.off_raw + rb * L.raw_bytes);
                float* const pw = s_pow + pb * (PROWS * PROW);
#pragma unroll 1
                for (int r = 0; r < ROUNDS; ++r) {
                    const int f0 = 2 * kFftWarps * r + 2 * warp;
                    if (t0 + f0 >= nfr) break;                          // both frames past the clip's end
                    const int f = f0 + h;
                    float2 v[16];
                    if constexpr (I16) {
                        // one 32-bit word = one packed complex point (two int16 samples)
                        const uint32_t* a = reinterpret_cast<const uint32_t*>(cur) + ((f * hop) >> 1) + j;
                        // all sixteen words in flight before the first conversion
                        uint32_t rw[16];
                        {
                            const uint32_t ra = smem_u32(a);
#define B2A_LDR(T) asm volatile("ld.shared.b32 %0, [%1+%2];" : "=r"(rw[T]) : "r"(ra), "n"((T) * 64))
                            B2A_LDR(0); B2A_LDR(1); B2A_LDR(2); B2A_LDR(3); B2A_LDR(4); B2A_LDR(5); B2A_LDR(6); B2A_LDR(7);
                            B2A_LDR(8); B2A_LDR(9); B2A_LDR(10); B2A_LDR(11); B2A_LDR(12); B2A_LDR(13); B2A_LDR(14); B2A_LDR(15);
#undef B2A_LDR
                        }
#pragma unroll
                        for (int t = 0; t < 16; ++t) v[t] = __fmul2_rn(cvt_pcm2(rw[t]), win[t]);
                    } else {
                        const float* a = cur + f * hop + 2 * j;
#pragma unroll
                        for (int t = 0; t < 16; ++t) v[t] = __fmul2_rn(*reinterpret_cast<const float2*>(a + 32 * t), win[t]);
                    }
                    Dft<16>::run(v);
#pragma unroll
                    for (int t = 0; t < 16; ++t) x1[t] = v[t];         // (128-bit stores would need 4 MOVs each to
                                                                       //  line the register pairs up)
                    __syncwarp();
                    v[0] = x2[0];
#pragma unroll
                    for (int t = 1; t < 16; ++t) v[t] = cmul(x2[XS * t], tw1[t - 1]);
                    Dft<16>::run(v);                                   // v[t] = Z[j + 16 t]
                    __syncwarp();
#pragma unroll
                    for (int t = 8; t < 16; ++t) mst[(t - 8) * 16] = v[t];
                    mst[8 * 16] = v[0];                                // "row 16": Z[256] == Z[0] for lane 0
                    __syncwarp();
                    // all eight mirror operands in flight at once (v[9..15] are dead by now): left to
                    // itself ptxas loads them one per split step and every step eats the latency
                    float2 Bm[8];
                    {
                        const uint32_t ma = smem_u32(mld);
#define B2A_LDM(R2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" \
                                 : "=f"(Bm[R2].x), "=f"(Bm[R2].y) : "r"(ma), "n"((7 - (R2)) * 16 * 8))
                        B2A_LDM(0); B2A_LDM(1); B2A_LDM(2); B2A_LDM(3); B2A_LDM(4); B2A_LDM(5); B2A_LDM(6); B2A_LDM(7);
#undef B2A_LDM
                    }
                    if (!pow_free) {                                   // the mel warps are done with this slot
                        mbar_wait(bar_pow_empty + pb, ((it / NPOW) & 1) ^ 1);
                        pow_free = true;
                    }
                    // element (bin k, frame f) lives at word (k>>1)*PROW + 2f + (k&1)
                    float* pk = pw + (j >> 1) * PROW + 2 * f + (j & 1);                 // bin j + 16 r2
                    float* pn = pw + ((NC - j) >> 1) * PROW + 2 * f + (j & 1);          // bin 256 - j - 16 r2
                    // split twiddles, two per 128-bit load, fetched one pair of steps ahead
                    const uint32_t ta = smem_u32(t2);
                    float4 w4, w4n;
#define B2A_LDT(DST, P) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" \
                                     : "=f"(DST.x), "=f"(DST.y), "=f"(DST.z), "=f"(DST.w) : "r"(ta), "n"((P) * 16 * 16))
                    B2A_LDT(w4n, 0);
#pragma unroll
                    for (int r2 = 0; r2 < 8; ++r2) {
                        const float2 B = Bm[r2];
                        if ((r2 & 1) == 0) {
                            w4 = w4n;
                            if (r2 == 0) B2A_LDT(w4n, 1);
                            if (r2 == 2) B2A_LDT(w4n, 2);
                            if (r2 == 4) B2A_LDT(w4n, 3);
                        }
                        const float2 w = (r2 & 1) ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
                        float2 xk, xnk;
                        rfft_split(v[r2], B, w, xk, xnk);              // 2 X[k], 2 X[256-k]
                        pk[8 * r2 * PROW] = xk.x * xk.x + xk.y * xk.y;         // 4|X|^2: the 1/4 lives
                        pn[-8 * r2 * PROW] = xnk.x * xnk.x + xnk.y * xnk.y;    // in the mel weights
                    }
#undef B2A_LDT
                    if (j == 0) pw[(NC / 4) * PROW + 2 * f] = 4.0f * (v[8].x * v[8].x + v[8].y * v[8].y);
                    __syncwarp();
                }
                // a warp with no frame in this tile still takes its turn on both barriers: every
                // pow_full phase must collect exactly one arrival per warp
                if (!pow_free) mbar_wait(bar_pow_empty + pb, ((it / NPOW) & 1) ^ 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pow_full + pb);
            }
        }
    } else {
        // =========================== mel warps (warp 0 of them also stages the raw tiles) =========
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMelRegs));   // (this .dec is what funds the FFT warps' .inc)
        const int mw = warp - kFftWarps, mtid = tid - 32 * kFftWarps;
        constexpr int V = 16 / (int)sizeof(E);                   // samples per 16 bytes
        const bool base_aligned = (reinterpret_cast<uintptr_t>(p.clips) & 15) == 0;

        // Stage tile (clip, t0) into raw slot `slot`.  Aligned zero-padded clips go by one TMA bulk
        // copy with the clip's head/tail zero-filled by plain stores; reflect padding or unaligned
        // clips use plain loads.  Exactly one arrival on the slot's barrier either way.
        auto stage = [&](long long clip, int t0, uint32_t slot) {
            E* const dst = reinterpret_cast<E*>(smem + L.off_raw + slot * L.raw_bytes);
            uint64_t* const bar = bar_raw_full + slot;
            const int n = RAG ? p.rag_len[clip] : p.n_samples;
            const long long e0 = RAG ? p.rag_in_off[clip] : clip * (long long)n;
            const E* cptr = reinterpret_cast<const E*>(p.clips) + e0;
            const int c0 = t0 * hop - NFFT / 2;
            const int lo = c0 < 0 ? 0 : c0;
            const int hi = (c0 + chunk < n) ? c0 + chunk : n;
            // samples the tile's valid frames read (a clip's last tile is mostly past its end; filling
            // the whole chunk cost this one warp ~5000 cycles per clip and stalled the pipeline)
            const int nfr = 1 + n / hop;
            const int vf = nfr - t0 < F ? nfr - t0 : F;
            const int need = ((vf - 1) * hop + NFFT + V - 1) & ~(V - 1);      // <= chunk
            const bool ok = p.pad_mode == 0 && base_aligned && hi > lo && (((e0 + lo) & (V - 1)) == 0) &&
                            (((lo - c0) & (V - 1)) == 0);
            const int nb = ok ? ((hi - lo) / V) * V : 0;         // bulk part, whole 16-byte units
            if (nb == 0) {
                for (int i = lane; i < need; i += 32) dst[i] = raw_sample<E>(cptr, c0 + i, n, p.pad_mode);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar);
                return;
            }
            // [0, head) and [head+nb, need) are not written by the bulk copy: zeros (centre padding),
            // except for the clip's last < V samples
            const int head = lo - c0, tb = head + nb;
            const int4 z4 = make_int4(0, 0, 0, 0);
            for (int i = lane * V; i < head; i += 32 * V) *reinterpret_cast<int4*>(dst + i) = z4;
            if (tb < need) {
                if (lane < V) dst[tb + lane] = raw_sample<E>(cptr, c0 + tb + lane, n, 0);
                for (int i = tb + V + lane * V; i < need; i += 32 * V) *reinterpret_cast<int4*>(dst + i) = z4;
            }
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses -> async write
                const int cb = (nb < need - head ? nb : need - head) * (int)sizeof(E);    // never more than needed
                mbar_expect_tx(bar, (uint32_t)cb);
                bulk_g2s(dst + head, cptr + lo, (uint32_t)cb, bar);
            }
        };
        // producer cursor: runs NRAW tiles ahead of the consumers
        long long pclip = blockIdx.x;
        int ptile = 0;
        uint32_t pit = 0;
        auto stage_next = [&]() {
            if (pclip >= p.n_clips) return;
            B2A_STAMP(0, pit);
            stage(pclip, ptile * F, pit % NRAW);
            ++pit;
            const int pnfr = RAG ? 1 + p.rag_len[pclip] / hop : p.n_frames;
            if (++ptile * F >= pnfr) { ptile = 0; pclip += gridDim.x; }
        };
        if (mw == 0)
            for (int i = 0; i < NRAW; ++i) stage_next();

        // Deferred normalisation (mel): the previous clip is rewritten in place a slice per tile of
        // the current one, its raw dB (L2) prefetched into registers BEFORE the wait for the power
        // tile, so the L2 round trip hides behind that wait and the band sweep.
        constexpr int NPF = 3;                                   // float4 per thread per tile
        float4* nq = nullptr;                                    // previous clip's features, as float4
        int nq_n4 = 0, nq_done = 0;                              // float4 count, float4 already rewritten
        float nq_vmax = 0.f, nq_lo = 0.f, nq_range = 1.f, nq_inv = 1.f;
        auto nrm = [&](float x) {
            // x / range by one Newton step on x * (1/range): correctly rounded for these operand
            // ranges (so the clip's peak is exactly 1.0, as with numpy's true division)
            const float num = fmaxf(x - nq_vmax, -p.top_db) - nq_lo;
            const float q = num * nq_inv;
            return fmaf(fmaf(-q, nq_range, num), nq_inv, q);
        };
        auto nrm4 = [&](float4 x) { return make_float4(nrm(x.x), nrm(x.y), nrm(x.z), nrm(x.w)); };
        // mfcc: the deferred pass is the per-row z-score (deep.py:326-328), element-wise because rows of
        // n_frames floats do not keep 16-byte alignment; (mean, sd) per coefficient wait in s_zs.
        constexpr int NPZ = 4;                                   // floats per thread per tile
        float* nz = nullptr;
        int nz_nfr = 1;
        float nz_inv = 1.f;
        const bool zfast = MFCC && p.n_mfcc <= kZFast;
        auto zs1 = [&](float x, int i) {
            const int k = __float2int_rd(((float)i + 0.5f) * nz_inv);    // row of element i (never near an integer)
            return (x - s_zs[2 * k]) * s_zs[2 * k + 1];                   // s_zs = (mean, 1 / (sd + 1e-8))
        };
        auto nq_finish = [&]() {                                 // whatever is left of the previous clip
            if constexpr (!MFCC) {
                for (int i = nq_done + mtid; i < nq_n4; i += kMelThreads) nq[i] = nrm4(nq[i]);
            } else {
                for (int i = nq_done + mtid; i < nq_n4; i += kMelThreads) nz[i] = zs1(nz[i], i);
            }
            nq_done = nq_n4;
        };
        float zS[4] = {0.f, 0.f, 0.f, 0.f}, zQ[4] = {0.f, 0.f, 0.f, 0.f}, zx0[4] = {0.f, 0.f, 0.f, 0.f};

        uint32_t it = 0;
        for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
            const int nfr = RAG ? 1 + p.rag_len[clip] / hop : p.n_frames;
            const int tiles = (nfr + F - 1) / F;
            float* const outb = RAG ? p.out + p.rag_out_off[clip]
                                    : p.out + (size_t)clip * (MFCC ? p.n_mfcc : n_mels) * nfr;
            float* const inter = MFCC ? p.inter + (size_t)blockIdx.x * n_mels * p.n_frames : outb;
            float vmax = -3.0e38f, vmin = 3.0e38f;

            for (int tile = 0; tile < tiles; ++tile, ++it) {
                const int t0 = tile * F;
                const uint32_t pb = it % NPOW;
                float4 nx[MFCC ? 1 : NPF];
                float zx[MFCC ? NPZ : 1];
                const bool nq_live = nq_done < nq_n4;
                if (nq_live) {
                    if constexpr (!MFCC) {
#pragma unroll
                        for (int k = 0; k < NPF; ++k) {
                            const int i = nq_done + mtid + k * kMelThreads;
                            if (i < nq_n4) nx[k] = nq[i];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < NPZ; ++k) {
                            const int i = nq_done + mtid + k * kMelThreads;
                            if (i < nq_n4) zx[k] = nz[i];
                        }
                    }
                }
                if (mw == 0) B2A_STAMP(5, it);
                mbar_wait(bar_pow_full + pb, (it / NPOW) & 1);     // power tile complete ...
                if (mw == 0) B2A_STAMP(6, it);
                if (mw == 0) stage_next();                         // ... and raw slot it % NRAW is free again
                // mel bands: lane = frame, warp-uniform band
                {
                    const int t = t0 + lane;
                    const bool valid = t < nfr;
                    float* const outp = inter + t;
                    const float2* pl = reinterpret_cast<const float2*>(s_pow + pb * (PROWS * PROW)) + lane;
                    if constexpr (SPEC) {
                        const uint32_t pla = smem_u32(pl);
#define B2A_LDS2(DST, ADDR, OFF) \
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(DST.x), "=f"(DST.y) : "r"(ADDR), "n"(OFF))
                        // headline configuration: every band unrolled, weights are FFMA immediates
#define B2A_EMIT(M, VAL)                                                            \
    {                                                                               \
        const float vv = db10(VAL);                                                 \
        if (KIND == 1) s_db[(M) * 32 + lane] = vv;                                  \
        if (valid) outp[(M) * nfr] = vv;      /* predicated store: the band sweep stays one basic block */ \
        vmax = fmaxf(vmax, valid ? vv : vmax);                                      \
        if (KIND != 2) vmin = fminf(vmin, valid ? vv : vmin);   /* KIND 2 clamps at clip end and needs only the peak */ \
    }
                        switch (mw) {
                            case 0: B2A_MEL_WARP0(pla, B2A_EMIT) break;
                            case 1: B2A_MEL_WARP1(pla, B2A_EMIT) break;
                            case 2: B2A_MEL_WARP2(pla, B2A_EMIT) break;
                            default: B2A_MEL_WARP3(pla, B2A_EMIT) break;
                        }
#undef B2A_EMIT
#undef B2A_LDS2
                    } else {
                        // per 4-bin step one 128-bit broadcast weight load and two 64-bit power
                        // loads (bands padded with zero weights)
                        for (int i = mw; i < n_mels; i += kMelWarps) {
                            const int4 d = s_desc[i];      // {pair-row offset (float2), n 4-bin steps, weight offset, m}
                            const float2* pr = pl + d.x;
                            const float4* wq = reinterpret_cast<const float4*>(s_melw + d.z);
                            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
                            for (int q4 = 0; q4 < d.y; ++q4) {
                                const float4 w = *wq++;
                                const float2 p0 = pr[0], p1 = pr[PROW / 2];
                                a0 = fmaf(w.x, p0.x, a0);
                                a1 = fmaf(w.y, p0.y, a1);
                                a2 = fmaf(w.z, p1.x, a2);
                                a3 = fmaf(w.w, p1.y, a3);
                                pr += PROW;
                            }
                            const float vv = db10((a0 + a1) + (a2 + a3));
                            if (MFCC) s_db[d.w * 32 + lane] = vv;
                            if (valid) {
                                outp[d.w * nfr] = vv;
                                vmax = fmaxf(vmax, vv);
                                vmin = fminf(vmin, vv);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pow_empty + pb);    // the FFT warps may refill this slot
                if (nq_live) {
                    if constexpr (!MFCC) {
#pragma unroll
                        for (int k = 0; k < NPF; ++k) {
                            const int i = nq_done + mtid + k * kMelThreads;
                            if (i < nq_n4) nq[i] = nrm4(nx[k]);
                        }
                        nq_done += NPF * kMelThreads;
                        // more bands than the prefetched slices cover per tile (n_mels > 48): the rest of
                        // this tile's share goes without prefetch, so the backlog never reaches the
                        // end-of-clip flush
                        if constexpr (!SPEC) {       // (the generated configuration has 40 bands: nothing left)
                            for (int more = (n_mels * (F / 4) + kMelThreads - 1) / kMelThreads - NPF; more > 0; --more) {
                                const int i = nq_done + mtid;
                                if (i < nq_n4) nq[i] = nrm4(nq[i]);
                                nq_done += kMelThreads;
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < NPZ; ++k) {
                            const int i = nq_done + mtid + k * kMelThreads;
                            if (i < nq_n4) nz[i] = zs1(zx[k], i);
                        }
                        nq_done += NPZ * kMelThreads;
                        if constexpr (KIND != 2) {   // (13 coefficients x 32 frames = 4 slices: nothing left)
                            for (int more = (p.n_mfcc * F + kMelThreads - 1) / kMelThreads - NPZ; more > 0; --more) {
                                const int i = nq_done + mtid;
                                if (i < nq_n4) nz[i] = zs1(nz[i], i);
                                nq_done += kMelThreads;
                            }
                        }
                    }
                    if (tile + 1 == tiles) nq_finish();            // short clip after a long one
                }
                if constexpr (KIND == 1) {
                    // DCT-II of this tile straight from shared memory, assuming the top_db clip
                    // (known only after the clip's last tile) will not engage; checked below.
                    // lane = frame; this warp owns coefficients k = mw + 4 g.  One conflict-free dB
                    // load and one 128-bit broadcast basis load feed four FMAs.
                    mel_sync();
                    const int t = t0 + lane;
                    const bool valid = t < nfr;
                    const float* const dbl = s_db + lane;
                    for (int c = 0; c < L.gp / 4; ++c) {
                        const float4* w4 = reinterpret_cast<const float4*>(s_dct + mw * L.gp) + c;
                        float a[4] = {0.f, 0.f, 0.f, 0.f};
                        // software pipeline: the loads of the next four bands are in flight while the
                        // current four are multiplied (this warp has its scheduler's LSU latency to itself)
                        const int ws = kMelWarps * L.gp / 4;
                        float dq[4];
                        float4 wq[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = u < n_mels ? u : n_mels - 1;
                            dq[u] = dbl[m * 32];
                            wq[u] = w4[m * ws];
                        }
                        for (int m0 = 0; m0 < n_mels; m0 += 4) {
                            float dc[4];
                            float4 wc[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { dc[u] = dq[u]; wc[u] = wq[u]; }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int m = m0 + 4 + u < n_mels ? m0 + 4 + u : n_mels - 1;
                                dq[u] = dbl[m * 32];
                                wq[u] = w4[m * ws];
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (m0 + u < n_mels) {
                                    a[0] = fmaf(wc[u].x, dc[u], a[0]);
                                    a[1] = fmaf(wc[u].y, dc[u], a[1]);
                                    a[2] = fmaf(wc[u].z, dc[u], a[2]);
                                    a[3] = fmaf(wc[u].w, dc[u], a[3]);
                                }
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int k = mw + kMelWarps * (4 * c + g);
                            if (k < p.n_mfcc && valid) outb[(size_t)k * nfr + t] = a[g];
                            if (c == 0) {
                                // running sums about the row's first sample for the one-sweep z-score
                                if (tile == 0) zx0[g] = __shfl_sync(0xffffffffu, a[g], 0);
                                const float dd = valid ? a[g] - zx0[g] : 0.f;
                                zS[g] += dd;
                                zQ[g] = fmaf(dd, dd, zQ[g]);
                            }
                        }
                    }
                    mel_sync();                                    // s_db is rewritten by the next tile
                }
            }

            // ---- per-clip reductions (mel warps only) ----------------------------------------------
            vmax = warp_max(vmax); vmin = warp_min(vmin);
            if (lane == 0) { s_red[mw] = vmax; s_red[32 + mw] = vmin; }
            mel_sync();                                            // also: every warp's raw dB stores are visible
            {
                const float a = (lane < kMelWarps) ? s_red[lane] : -3.0e38f;
                const float b = (lane < kMelWarps) ? s_red[32 + lane] : 3.0e38f;
                vmax = warp_max(a); vmin = warp_min(b);
            }
            if constexpr (!MFCC) {
                nq_finish();                                          // (only if this clip had no tile)
                nq_vmax = vmax;
                nq_lo = fmaxf(vmin - vmax, -p.top_db);
                nq_range = (0.0f - nq_lo) + 1e-8f;
                nq_inv = __frcp_rn(nq_range);
                const int total = n_mels * nfr;
                if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(inter) & 15) == 0)) {
                    nq = reinterpret_cast<float4*>(inter);            // rewritten during the next clip's tiles
                    nq_n4 = total / 4;
                    nq_done = 0;
                } else {
                    for (int i = mtid; i < total; i += kMelThreads) inter[i] = nrm(inter[i]);
                    nq_n4 = nq_done = 0;
                }
            } else if constexpr (KIND == 2) {
                // Headline mfcc: the DCT of the whole clip runs HERE, once, from the raw-dB scratch (L2
                // resident), one frame per thread: the top_db threshold is known by now (no recompute
                // path), the 520 immediate-weight FFMAs are fetched once per clip instead of once per
                // tile (in the tile loop they pushed the hot code past the 32 KB instruction cache and
                // the mel warps, not the FFT warps, set the pace: 2.6 M clips/s), and the per-row
                // statistics of the z-score (deep.py:326-328) are float64 sums taken on the way.
                // The FFT warps run ahead into the next clip meanwhile (two power tiles of slack).
                float* const outc = outb;
                const float thr = vmax - p.top_db;
                nq_finish();                                          // the previous clip's z-score, if any is left
                mel_sync();                                           // ... by every warp, before s_zs changes
                // A frame's 40 raw dB come from L2 by cp.async, all in flight at once and without holding
                // registers (plain loads in batches of eight made this phase a chain of L2 round trips:
                // 31 k cycles per clip), into a column of shared memory private to the thread — no
                // barrier, the thread only reads what it copied itself.  KIND 2 has no other use for the
                // dB tile / partial-sum / basis regions, which the columns overlay.
                float* const s_col = reinterpret_cast<float*>(smem + L.off_db) + mtid;           // [n_mels][128]
                double* const s_zd = reinterpret_cast<double*>(smem + L.off_db + B2A_MELSPEC_NMELS * kMelThreads * 4);   // [mel warp][coefficient][S, Q]
                const uint32_t s_col_a = smem_u32(s_col);
#define B2A_DCT_LD(M) fmaxf(s_col[(M) * kMelThreads], thr)
                // One staged column serves both coefficient groups: the even group's float64 sums live in
                // registers across the frames, the odd group's in shared memory next to the thread's
                // column (both sets in registers would not fit the 64-register role; a second pass over
                // the frames cost a second L2 round trip per frame).
                constexpr int NK0 = B2A_DCT_GROUP0_NK, NK1 = B2A_DCT_GROUP1_NK;
                double* const zq = reinterpret_cast<double*>(smem + L.off_zq) + mtid;          // [2 NK1][128]
                double S0[NK0], Q0[NK0];
#pragma unroll
                for (int k = 0; k < NK0; ++k) { S0[k] = 0.0; Q0[k] = 0.0; }
#pragma unroll
                for (int k = 0; k < 2 * NK1; ++k) zq[k * kMelThreads] = 0.0;
#pragma unroll 1
                for (int t = mtid; t < nfr; t += kMelThreads) {
                    {
                        const float* src = inter + t;
#pragma unroll
                        for (int m = 0; m < B2A_MELSPEC_NMELS; ++m, src += nfr)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_col_a + m * kMelThreads * 4), "l"(src) : "memory");
                        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
                    }
                    {
                        B2A_DCT_T a[NK0];
#pragma unroll
                        for (int k = 0; k < NK0; ++k) a[k] = 0;
                        B2A_DCT_GROUP0(B2A_DCT_LD, a)
#pragma unroll
                        for (int k = 0; k < NK0; ++k) {
                            const float af = (float)a[k];
                            outc[(size_t)B2A_DCT_GROUP0_KOF(k) * nfr + t] = af;
                            const double ad = (double)af;      // statistics of the stored float32 row (deep.py:326-328)
                            S0[k] += ad;
                            Q0[k] = fma(ad, ad, Q0[k]);
                        }
                    }
                    {
                        B2A_DCT_T a[NK1];
#pragma unroll
                        for (int k = 0; k < NK1; ++k) a[k] = 0;
                        B2A_DCT_GROUP1(B2A_DCT_LD, a)
#pragma unroll
                        for (int k = 0; k < NK1; ++k) {
                            const float af = (float)a[k];
                            outc[(size_t)B2A_DCT_GROUP1_KOF(k) * nfr + t] = af;
                            const double ad = (double)af;
                            zq[(2 * k) * kMelThreads] += ad;
                            zq[(2 * k + 1) * kMelThreads] = fma(ad, ad, zq[(2 * k + 1) * kMelThreads]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NK0; ++k) {
                    const double Sw = warp_sum_d(S0[k]), Qw = warp_sum_d(Q0[k]);
                    if (lane == 0) { s_zd[(mw * B2A_DCTSPEC_NMFCC + B2A_DCT_GROUP0_KOF(k)) * 2] = Sw; s_zd[(mw * B2A_DCTSPEC_NMFCC + B2A_DCT_GROUP0_KOF(k)) * 2 + 1] = Qw; }
                }
#pragma unroll
                for (int k = 0; k < NK1; ++k) {
                    const double Sw = warp_sum_d(zq[(2 * k) * kMelThreads]), Qw = warp_sum_d(zq[(2 * k + 1) * kMelThreads]);
                    if (lane == 0) { s_zd[(mw * B2A_DCTSPEC_NMFCC + B2A_DCT_GROUP1_KOF(k)) * 2] = Sw; s_zd[(mw * B2A_DCTSPEC_NMFCC + B2A_DCT_GROUP1_KOF(k)) * 2 + 1] = Qw; }
                }
#undef B2A_DCT_LD
                mel_sync();
                if (mtid < B2A_DCTSPEC_NMFCC) {
                    double S = 0.0, Q = 0.0;
#pragma unroll
                    for (int w = 0; w < kMelWarps; ++w) {
                        S += s_zd[(w * B2A_DCTSPEC_NMFCC + mtid) * 2];
                        Q += s_zd[(w * B2A_DCTSPEC_NMFCC + mtid) * 2 + 1];
                    }
                    // constant rows (silence): S = n c exactly, so the mean is c and every z is exactly 0
                    const double mean = S / (double)nfr;
                    const double var = fmax(Q / (double)nfr - mean * mean, 0.0);
                    s_zs[2 * mtid] = (float)mean;
                    s_zs[2 * mtid + 1] = __fdiv_rn(1.0f, sqrtf((float)var) + 1e-8f);
                }
                nz = outc;
                nz_nfr = nfr;
                nz_inv = __fdiv_rn(1.0f, (float)nfr);
                nq_n4 = B2A_DCTSPEC_NMFCC * nfr;
                nq_done = 0;
                mel_sync();                                           // s_zs and the DCT rows are visible to every mel warp
                continue;
            } else {
                float* outc = outb;
                const float thr = vmax - p.top_db;
                if (vmin < thr) {
                    // rare: some band fell more than top_db below the clip's peak, so the clipped dB differ
                    // from what the in-tile DCT saw -> recompute from the raw-dB scratch (L2 resident)
                    float* s_l = s_db;                                // [n_mels][32] clipped dB tile
                    for (int t0 = 0; t0 < nfr; t0 += 32) {
                        mel_sync();
                        for (int i = mtid; i < n_mels * 32; i += kMelThreads) {
                            const int m = i >> 5, f = i & 31, t = t0 + f;
                            s_l[i] = (t < nfr) ? fmaxf(inter[(size_t)m * nfr + t], thr) : 0.f;
                        }
                        mel_sync();
                        for (int i = mtid; i < p.n_mfcc * 32; i += kMelThreads) {
                            const int k = i >> 5, f = i & 31, t = t0 + f;
                            const float* d = p.dct + (size_t)k * n_mels;
                            float acc = 0.f;
#pragma unroll 4
                            for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(d + m), s_l[m * 32 + f], acc);
                            if (t < nfr) outc[(size_t)k * nfr + t] = acc;
                        }
                    }
                }
                nq_finish();                                          // the previous clip's z-score, if any is left
                mel_sync();                                           // ... by every warp, before s_zs changes
                const float fn = (float)nfr;
                if (zfast && !(vmin < thr)) {
                    // common case: mean and variance from the running sums; the rows are rewritten
                    // during the next clip's tiles (deferred, prefetched)
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float S = warp_sum(zS[g]), Q = warp_sum(zQ[g]);
                        const int k = mw + kMelWarps * g;
                        if (lane == 0 && k < p.n_mfcc) {
                            const float md = __fdiv_rn(S, fn);
                            s_zs[2 * k] = zx0[g] + md;
                            s_zs[2 * k + 1] = __fdiv_rn(1.0f, sqrtf(fmaxf(fmaf(-md, md, __fdiv_rn(Q, fn)), 0.f)) + 1e-8f);
                        }
                        zS[g] = 0.f; zQ[g] = 0.f;
                    }
                    nz = outc;
                    nz_nfr = nfr;
                    nz_inv = __fdiv_rn(1.0f, fn);
                    nq_n4 = p.n_mfcc * nfr;
                    nq_done = 0;
                    mel_sync();
                    continue;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) { zS[g] = 0.f; zQ[g] = 0.f; }
                nq_n4 = nq_done = 0;
                mel_sync();                                           // DCT stores visible to every mel warp
                // deep.py:326-328: per-row z-score, three sweeps over an L2-resident row
                for (int k = mw; k < p.n_mfcc; k += kMelWarps) {
                    float* row = outc + (size_t)k * nfr;
                    const float x0 = row[0];
                    float s_ = 0.f;
                    for (int t = lane; t < nfr; t += 32) s_ += row[t] - x0;
                    const float mean = x0 + __fdiv_rn(warp_sum(s_), fn);
                    float ss = 0.f;
                    for (int t = lane; t < nfr; t += 32) { const float d = row[t] - mean; ss = fmaf(d, d, ss); }
                    const float sd = sqrtf(__fdiv_rn(warp_sum(ss), fn)) + 1e-8f;
                    for (int t = lane; t < nfr; t += 32) row[t] = __fdiv_rn(row[t] - mean, sd);
                }
            }
            mel_sync();                                               // s_red / s_db are reused by the next clip
        }
        nq_finish();                                                  // the CTA's last clip
    }
}

}  // namespace

size_t logmel512_smem_bytes(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc) {
    return (size_t)make_layout(hop, n_mels, mel_wpad, i16, n_mfcc).total + 128;
}

int logmel512_ctas_per_sm() { return 1; }
int logmel512_mel_warps() { return kMelWarps; }

bool logmel512_has_special(int sample_rate, int n_mels) {
    return sample_rate == B2A_MELSPEC_SR && n_mels == B2A_MELSPEC_NMELS && B2A_MELSPEC_NFFT == NFFT;
}

template <bool I16, int KIND, bool SPEC, bool RAG>
static cudaError_t launch_k(const FrontParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = logmel512_kernel<I16, KIND, SPEC, RAG>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool SPEC, bool RAG>
static cudaError_t launch_s(const FrontParams& p, bool i16, int kind, int grid, size_t smem, cudaStream_t st) {
    if (kind == 0) return i16 ? launch_k<true, 0, SPEC, RAG>(p, grid, smem, st) : launch_k<false, 0, SPEC, RAG>(p, grid, smem, st);
    if constexpr (SPEC) {
        if (p.n_mfcc == B2A_DCTSPEC_NMFCC)      // headline mfcc: generated DCT
            return i16 ? launch_k<true, 2, SPEC, RAG>(p, grid, smem, st) : launch_k<false, 2, SPEC, RAG>(p, grid, smem, st);
    }
    return i16 ? launch_k<true, 1, SPEC, RAG>(p, grid, smem, st) : launch_k<false, 1, SPEC, RAG>(p, grid, smem, st);
}

cudaError_t launch_logmel512(const FrontParams& p, bool i16, int kind, int grid, cudaStream_t st) {
    const size_t smem = logmel512_smem_bytes(p.hop, p.n_mels, p.mel_wpad, i16, kind == 1 ? p.n_mfcc : 0);
    if (p.rag_len) return p.mel_special ? launch_s<true, true>(p, i16, kind, grid, smem, st)
                                        : launch_s<false, true>(p, i16, kind, grid, smem, st);
    return p.mel_special ? launch_s<true, false>(p, i16, kind, grid, smem, st)
                         : launch_s<false, false>(p, i16, kind, grid, smem, st);
}

}  // namespace b2a

#ifdef B2A_TRACE
extern "C" int b2a_debug_trace_read(long long* out, int n) {
    if (n > 8 * 512) n = 8 * 512;
    return (int)cudaMemcpyFromSymbol(out, b2a::g_trace, sizeof(long long) * n);
}
#endif
