// logmel1024.cu — n_fft = 1024 fused log-mel / MFCC front end, sm_100a only.
//
// n_fft 1024 / hop 512 is the reference default of audio_mfcc_seq (deep.py:290-297) and of every
// non-Nicla experiment (experiments/birdeep_feature_extraction.yaml:31-47).  A 1024-sample frame is the
// radix-2 combination of two 512-sample real FFTs, one over its even and one over its odd samples:
//     X[k] = E0[k] + W^k E1[k],   X[512 - k] = conj(E0[k] - W^k E1[k]),   W = exp(-2 pi i / 1024), k = 0..256
// so the transform is the half-warp machinery of logmel512.cu twice: ONE WARP PER FRAME, half-warp h
// transforms samples 4q + h and 4q + 2 + h as the packed complex sequence of its 512-sample real FFT
// (both radix-16 passes in registers, packed FP32), and after the split step the two halves swap one
// operand per bin through warp shuffles and finish with a twiddle butterfly each.
//
// One persistent CTA per SM, 640 threads, 96 registers each (no setmaxnreg):
//   * 16 FRAME warps.  A tile is 16 frames, one per warp: FFT -> 4|X|^2 into the tile's [bin pair][16 frames] power
//     tile (half 1 windows its samples with (-1)^m, which turns its sub-FFT into E1[k + 256]: both halves then hand
//     each other the same-named register in the combination step and run identical code) -> named barrier among the
//     512 threads -> ALL sixteen warps sweep the tile's mel bands: lane & 15 = frame, the two halves of a warp take
//     neighbouring bands of the width-sorted order (padded to the same number of 4-bin steps on the host), broadcast
//     weights, conflict-free powers, packed FMAs -> 10 log10 into the [row][16] output tile -> named barrier ->
//     (mfcc) the tile's DCT-II as a register-blocked 16 frames x n_mfcc x n_mels product (4 x 4 block per lane, the
//     eight band lanes folded by a transposing shuffle reduction).  Raw PCM comes from the TMA-staged ring.
//     (History — DESIGN.md 3.2, profiles/r2_ncu_1024_summary.txt: four dedicated mel warps as in logmel512.cu ran
//     their dependent load -> FMA chains at ~0.1 instructions per cycle and stalled the FFT warps 39 % of the time;
//     a per-frame sweep with lane = band put 676 shared-memory wavefronts per frame on the load pipe; a per-frame
//     DCT with lane = coefficient was half of all shared-memory traffic.)
//   * 4 EPILOGUE warps, warp 0 also the TMA producer: move finished tiles to global memory with
//     coalesced stores, track the clip's max / min, and rewrite the PREVIOUS clip in place during the
//     current clip's tiles — power_to_db(ref=max, top_db) + min-max for mel, the per-row z-score
//     (deep.py:326-328) from per-thread running sums for mfcc.
//   * mbarriers only: raw_full (TMA bytes) -> frame warps; tile_full (16 arrivals) -> epilogue;
//     tile_empty (4 arrivals) -> frame warps.
//
// Reference arithmetic: deep.py:126-134 (mel), :318-328 (mfcc) via librosa 0.11.0.
#include "frontend.h"
#include "fft_core.cuh"
#include "ws_common.cuh"

#include <cstdint>
#include <type_traits>

namespace b2a {

namespace {

using namespace ws;

constexpr int kFftWarps = 16, kMelWarps = 4;
constexpr int kThreads = 32 * (kFftWarps + kMelWarps);
constexpr int kMelThreads = 32 * kMelWarps;
constexpr int NFFT = 1024, F = 16;              // frame length (two 512-sample sub-FFTs of 256 complex points), frames per tile
constexpr int kMaxRaw = 3;
__host__ __device__ constexpr int nraw(bool i16, bool mfcc) { return i16 ? 3 : 2; }
constexpr int NTILE = 2;                        // ring of finished [row][16] tiles
constexpr int XS = 17;                          // exchange row stride (float2), as in logmel512.cu
constexpr int XSLOT = 16 * XS + 2;
constexpr int PROW = 2 * F + 4;                 // power tile: row = 2 adjacent bins x (16 frames + 2 pad)
constexpr int PROWS = 260;                      // bin pairs (0,1)..(512,513) + 3 zero rows for 8-bin padding
constexpr int kMaxCoefPerLane = 2;              // n_mfcc <= 64

__device__ __forceinline__ void mel_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kMelThreads) : "memory");
}
// one int16 of a 32-bit word, sign-extended: sel = 0x9910 (low) / 0xBB32 (high)
__device__ __forceinline__ int pcm_pick(uint32_t w, uint32_t sel) {
    int r;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(w), "r"(sel));
    return r;
}
// (e + w conj(o), e - w conj(o)); the second as 2e - first like bfly_w
__device__ __forceinline__ void bfly_wc(float2 e, float2 o, float2 w, float2& a, float2& b) {
    a = __ffma2_rn(make_float2(o.y, o.x), bc2(w.y), __ffma2_rn(o, make_float2(w.x, -w.x), e));
    b = __ffma2_rn(e, bc2(2.0f), make_float2(-a.x, -a.y));
}

// barrier among the 16 frame warps only (named barrier 2)
__device__ __forceinline__ void frame_warps_sync() {
    asm volatile("bar.sync 2, %0;" ::"n"(32 * kFftWarps) : "memory");
}

// floats between the dB blocks of two frame quads: + 4 keeps a half-warp's sixteen frames on sixteen banks
__host__ __device__ constexpr int dbc_quad(int n_mels) { return 4 * n_mels + 4; }

struct Layout {
    int chunk, raw_bytes, rows;                  // rows of a tile: n_mels (+ n_mfcc)
    int off_raw, off_xch, off_pow, off_tile, off_win, off_tw2, off_tw1, off_twc, off_melw, off_melk, off_red, off_bar, off_dbc, off_dct, total;
};

__host__ __device__ inline Layout make_layout(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc) {
    const bool mfcc = n_mfcc > 0;
    Layout L;
    L.chunk = (hop * (F - 1) + NFFT + 7) & ~7;
    L.rows = n_mels + (mfcc ? n_mfcc : 0);
    int o = 0;
    auto take = [&](int bytes) { int r = o; o += (bytes + 15) & ~15; return r; };
    L.raw_bytes = ((L.chunk + 8) * (i16 ? 2 : 4) + 15) & ~15;      // + 8: a tile may sit up to 7 elements into its slot
    L.off_raw = take(nraw(i16, mfcc) * L.raw_bytes);
    L.off_xch = take(2 * kFftWarps * XSLOT * 8);
    L.off_pow = take(PROWS * PROW * 4);               // power tile of the current 16 frames: [bin pair][frame]
    L.off_tile = take(NTILE * L.rows * F * 4);        // finished tiles: [row][16 frames]
    L.off_win = take(16 * 32 * 8);                    // window pairs (w[4q + h], w[4q + 2 + h]) as [t][lane], q = j + 16 t
    L.off_tw2 = take(8 * 16 * 8);                     // split twiddles of the 512-sample sub-FFTs
    L.off_tw1 = take(15 * 16 * 8);                    // pass-2 twiddles exp(-2 pi i t j / 256) as [t - 1][j]
    L.off_twc = take(257 * 8);                        // combine twiddles W_1024^k, k = 0..256
    L.off_melw = take(mel_wpad * 4);                  // banded weights in 4-bin steps, x 0.25 (the tile holds 4|X|^2)
    L.off_melk = take(n_mels * 16);                   // per position: {pair-row offset, 4-bin steps, weight offset, band}
    L.off_red = take((64 + 2 * (mfcc ? n_mfcc : 1)) * 4);
    L.off_bar = take((kMaxRaw + 2 * NTILE) * 8 + kMaxRaw * 4);
    L.off_dbc = take(mfcc ? 4 * dbc_quad(n_mels) * 4 : 0);   // mfcc: dB of the tile as [frame quad][band][4 frames]
    L.off_dct = take(mfcc ? n_mels * ((n_mfcc + 3) & ~3) * 4 : 0);   // DCT-II basis as [coefficient / 4][band][4], zero rows past n_mfcc
    L.total = o;
    return L;
}

template <bool I16, int KIND, bool RAG>
__global__ void __launch_bounds__(kThreads, 1) logmel1024_kernel(FrontParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool MFCC = KIND == 1;
    const Layout L = make_layout(p.hop, p.n_mels, p.mel_wpad, I16, MFCC ? p.n_mfcc : 0);
    float2* const s_xch = reinterpret_cast<float2*>(smem + L.off_xch);
    float* const s_pow = reinterpret_cast<float*>(smem + L.off_pow);
    float* const s_tile = reinterpret_cast<float*>(smem + L.off_tile);
    float2* const s_win = reinterpret_cast<float2*>(smem + L.off_win);
    float2* const s_tw2 = reinterpret_cast<float2*>(smem + L.off_tw2);
    float2* const s_tw1 = reinterpret_cast<float2*>(smem + L.off_tw1);
    float2* const s_twc = reinterpret_cast<float2*>(smem + L.off_twc);
    float* const s_melw = reinterpret_cast<float*>(smem + L.off_melw);
    int4* const s_desc = reinterpret_cast<int4*>(smem + L.off_melk);
    float* const s_red = reinterpret_cast<float*>(smem + L.off_red);
    float* const s_zs = s_red + 64;
    float* const s_dbc = reinterpret_cast<float*>(smem + L.off_dbc);
    float* const s_dct = reinterpret_cast<float*>(smem + L.off_dct);
    uint64_t* const bar_raw_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* const bar_tile_full = bar_raw_full + kMaxRaw;
    uint64_t* const bar_tile_empty = bar_tile_full + NTILE;
    int* const s_sft = reinterpret_cast<int*>(bar_tile_empty + NTILE);    // tile sample c0 + i sits at slot element i + s_sft[slot]

    constexpr int NRAW = nraw(I16, MFCC);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hop = p.hop, n_mels = p.n_mels, chunk = L.chunk, rows = L.rows;
    using E = typename std::conditional<I16, int16_t, float>::type;

    // ---- per-CTA tables (p.tw = exp(-2 pi i k / 512), p.tw2 = exp(-2 pi i k / 1024)) ---------------------
    for (int i = tid; i < 128; i += kThreads) {              // split twiddles exp(-i pi (j + 16 r) / 256), two per
        const int r = i >> 4, jj = i & 15;                   // conflict-free 128-bit load (logmel512.cu layout)
        s_tw2[(r >> 1) * 32 + jj * 2 + (r & 1)] = p.tw[jj + 16 * r];
    }
    for (int i = tid; i < 16 * 32; i += kThreads) {          // librosa.load's exact 1/32768 rides on the window (int16 input)
        const int t = i >> 5, l = i & 31, q = (l & 15) + 16 * t, hh = l >> 4;
        const float sc = I16 ? (1.0f / 32768.0f) : 1.0f;
        // half 1 (odd samples) is modulated by (-1)^m, m its sub-sequence index: its sub-FFT then holds
        // Y[k] = E1[k + 256], so that its "v" is conj(E1[256 - k]) and its mirror operand conj(E1[k]) — the
        // two halves hand each other the same-named register in the combination step
        s_win[i] = make_float2(p.window[4 * q + hh] * sc, p.window[4 * q + 2 + hh] * (hh ? -sc : sc));
    }
    for (int i = tid; i < 15 * 16; i += kThreads) s_tw1[i] = p.tw[2 * ((i >> 4) + 1) * (i & 15)];
    for (int i = tid; i < 257; i += kThreads) s_twc[i] = p.tw2[i];
    for (int i = tid; i < p.mel_wpad; i += kThreads) s_melw[i] = p.mel_wq[i];
    for (int i = tid; i < n_mels; i += kThreads) {
        const int m = p.mel_order[i];                        // positions in descending band width: neighbours pair up
        s_desc[i] = make_int4((p.mel_k0e[m] >> 1) * (PROW / 2), p.mel_cnt4[m], p.mel_off4[m], m);
    }
    for (int i = tid; i < 4 * PROW; i += kThreads) s_pow[256 * PROW + i] = 0.f;   // bins 512..519 (512 is rewritten per tile)
    if constexpr (MFCC) {
        for (int i = tid; i < n_mels * ((p.n_mfcc + 3) & ~3); i += kThreads) {   // [k / 4][m][k & 3]
            const int e = i & 3, m = (i >> 2) % n_mels, k = 4 * ((i >> 2) / n_mels) + e;
            s_dct[i] = k < p.n_mfcc ? p.dct[(size_t)k * n_mels + m] : 0.f;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < NRAW; ++i) mbar_init(bar_raw_full + i, 1);
        for (int i = 0; i < NTILE; ++i) { mbar_init(bar_tile_full + i, kFftWarps); mbar_init(bar_tile_empty + i, kMelWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (blockIdx.x >= p.n_clips) return;

    if (warp < kFftWarps) {
        // =========================== frame warps: one frame per warp per tile ==============================
        // (no setmaxnreg: every role fits the launch allocation of 96 registers — an .inc here could only be
        //  funded by a .dec of the other warps, and with this much shared memory there is no L1 to spill into)
        const int j = lane & 15, h = lane >> 4;
        // (window pairs and pass-2 twiddles come from shared memory: sixteen + fifteen register pairs next to
        //  v[] and the band / DCT loops do not fit 96 registers, and spills have no L1 to land in)
        const float2* const win = s_win + lane;
        const float2* const tw1 = s_tw1 + j;
        float2* const xs = s_xch + (2 * warp + h) * XSLOT;
        float2* const x1 = xs + XS * j;
        float2* const x2 = xs + j;
        float2* const mst = xs + j;
        const float2* const mld = xs + (j ? 16 - j : 16);
        const float4* const t2 = reinterpret_cast<const float4*>(s_tw2) + j;
        const uint32_t psel = h ? 0xBB32u : 0x9910u;             // int16: prmt selector of the half's sample of each 32-bit
                                                                 // word, sign-extended (nibble msb = replicate the sign)

        uint32_t it = 0;
        for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
            const int nfr = RAG ? 1 + p.rag_len[clip] / hop : p.n_frames;
            const int tiles = (nfr + F - 1) / F;
            for (int tile = 0; tile < tiles; ++tile, ++it) {
                const int t0 = tile * F;
                const uint32_t rb = it % NRAW, pb = it % NTILE;
                mbar_wait(bar_raw_full + rb, (it / NRAW) & 1);
                const E* const cur = reinterpret_cast<const E*>(smem + L.off_raw + rb * L.raw_bytes);
                const int f = warp;
                if (t0 + f < nfr) {
                    float2 v[16];
                    const int sft = s_sft[rb];                          // (visible: written before the slot's barrier completed)
                    if constexpr (I16) {
                        // 64 bits = samples 4q .. 4q+3; this half takes (4q + h, 4q + 2 + h)
                        const uint32_t ra = smem_u32(reinterpret_cast<const int16_t*>(cur) + f * hop + 4 * j + sft);
                        if ((sft & 3) == 0) {
                            // two batches of eight loads: sixteen 64-bit words would hold 32 registers at once
#define B2A_LDR(T) asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(rw[(T) & 7].x), "=r"(rw[(T) & 7].y) : "r"(ra), "n"((T) * 128))
#pragma unroll
                            for (int hb = 0; hb < 2; ++hb) {
                                uint2 rw[8];
                                if (hb == 0) { B2A_LDR(0); B2A_LDR(1); B2A_LDR(2); B2A_LDR(3); B2A_LDR(4); B2A_LDR(5); B2A_LDR(6); B2A_LDR(7); }
                                else { B2A_LDR(8); B2A_LDR(9); B2A_LDR(10); B2A_LDR(11); B2A_LDR(12); B2A_LDR(13); B2A_LDR(14); B2A_LDR(15); }
#pragma unroll
                                for (int t = 0; t < 8; ++t) {
                                    const float a = __int2float_rn(pcm_pick(rw[t].x, psel));
                                    const float b = __int2float_rn(pcm_pick(rw[t].y, psel));
                                    v[8 * hb + t] = __fmul2_rn(make_float2(a, b), win[32 * (8 * hb + t)]);
                                }
                            }
#undef B2A_LDR
                        } else {
                            // clip whose start is 4 (not 8) bytes off the 16-byte grid: the same words as two 32-bit loads
#pragma unroll
                            for (int t = 0; t < 16; ++t) {
                                uint32_t w0, w1;
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(ra + 128 * t));
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(ra + 128 * t + 4));
                                const float a = __int2float_rn(pcm_pick(w0, psel));
                                const float b = __int2float_rn(pcm_pick(w1, psel));
                                v[t] = __fmul2_rn(make_float2(a, b), win[32 * t]);
                            }
                        }
                    } else {
                        const float* a = reinterpret_cast<const float*>(cur) + f * hop + 4 * j + sft;
                        if (sft == 0) {
#pragma unroll
                            for (int t = 0; t < 16; ++t) {
                                const float4 q4 = *reinterpret_cast<const float4*>(a + 64 * t);
                                v[t] = __fmul2_rn(h ? make_float2(q4.y, q4.w) : make_float2(q4.x, q4.z), win[32 * t]);
                            }
                        } else {
#pragma unroll
                            for (int t = 0; t < 16; ++t)
                                v[t] = __fmul2_rn(make_float2(a[64 * t + h], a[64 * t + 2 + h]), win[32 * t]);
                        }
                    }
                    Dft<16>::run(v);
#pragma unroll
                    for (int t = 0; t < 16; ++t) x1[t] = v[t];
                    __syncwarp();
                    v[0] = x2[0];
#pragma unroll
                    for (int t = 1; t < 16; ++t) v[t] = cmul(x2[XS * t], tw1[16 * (t - 1)]);
                    Dft<16>::run(v);                                   // v[t] = Z_h[j + 16 t]
                    __syncwarp();
#pragma unroll
                    for (int t = 8; t < 16; ++t) mst[(t - 8) * 16] = v[t];
                    mst[8 * 16] = v[0];                                // "row 16": Z[256] == Z[0] for lane 0
                    __syncwarp();
                    float2 Bm[8];
                    {
                        const uint32_t ma = smem_u32(mld);
#define B2A_LDM(R2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" \
                                 : "=f"(Bm[R2].x), "=f"(Bm[R2].y) : "r"(ma), "n"((7 - (R2)) * 16 * 8))
                        B2A_LDM(0); B2A_LDM(1); B2A_LDM(2); B2A_LDM(3); B2A_LDM(4); B2A_LDM(5); B2A_LDM(6); B2A_LDM(7);
#undef B2A_LDM
                    }
                    // split step of the sub-FFT, in place: v[r2] = 2 E_h[j + 16 r2], Bm[r2] = 2 E_h[256 - j - 16 r2]
                    {
                        const uint32_t ta = smem_u32(t2);
                        float4 w4[4];
#define B2A_LDT(P) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" \
                                : "=f"(w4[P].x), "=f"(w4[P].y), "=f"(w4[P].z), "=f"(w4[P].w) : "r"(ta), "n"((P) * 16 * 16))
                        B2A_LDT(0); B2A_LDT(1); B2A_LDT(2); B2A_LDT(3);
#undef B2A_LDT
#pragma unroll
                        for (int r2 = 0; r2 < 8; ++r2) {
                            const float4 q4 = w4[r2 >> 1];
                            const float2 w = (r2 & 1) ? make_float2(q4.z, q4.w) : make_float2(q4.x, q4.y);
                            float2 xk, xnk;
                            rfft_split(v[r2], Bm[r2], w, xk, xnk);
                            v[r2] = xk; Bm[r2] = xnk;
                        }
                    }
                    // E_h[128] = conj(Z_h[128]) lives in lane j == 0 only (v[8] is untouched by the split)
                    const float2 e128 = make_float2(2.0f * v[8].x, -2.0f * v[8].y);
                    // ---- radix-2 combination of the two halves -------------------------------------------
                    // half 0 finishes bins kk = j + 16 r2 (and 512 - kk) from v = 2 E0[kk] and the partner's mirror
                    // operand conj(2 E1[kk]); half 1 finishes kk = 256 - j - 16 r2 (and 512 - kk) from v = conj(2 E1[kk])
                    // and the partner's mirror operand 2 E0[kk].  |E0 + W E1| = |v + W conj(recv)| in both (for half 1:
                    // conjugate, then multiply by the unit W), so every lane runs the same code on the same registers.
                    {
                        const int kk0 = h ? 256 - j : j;
                        const float2* wc = s_twc + kk0;
                        const int wstep = h ? -16 : 16, pstep = h ? -8 * PROW : 8 * PROW;
                        float* p1 = s_pow + 2 * f + (kk0 >> 1) * PROW + (kk0 & 1);              // bin kk: word (kk >> 1) PROW + 2 frame + (kk & 1)
                        float* p2 = s_pow + 2 * f + ((512 - kk0) >> 1) * PROW + (kk0 & 1);      // bin 512 - kk (same parity)
#pragma unroll
                        for (int r2 = 0; r2 < 8; ++r2) {
                            float2 recv;
                            recv.x = __shfl_xor_sync(0xffffffffu, Bm[r2].x, 16);
                            recv.y = __shfl_xor_sync(0xffffffffu, Bm[r2].y, 16);
                            float2 x1c, x2c;
                            bfly_wc(v[r2], recv, *wc, x1c, x2c);
                            *p1 = x1c.x * x1c.x + x1c.y * x1c.y;          // 4|X|^2: the 1/4 lives in the mel weights
                            *p2 = x2c.x * x2c.x + x2c.y * x2c.y;
                            wc += wstep; p1 += pstep; p2 -= pstep;
                        }
                    }
                    {   // bins 128 and 384 (lane j == 0 of half 0 has both operands after one more exchange;
                        // half 1's is conj(2 E1[128]) under its modulation)
                        float2 r128;
                        r128.x = __shfl_xor_sync(0xffffffffu, e128.x, 16);
                        r128.y = -__shfl_xor_sync(0xffffffffu, e128.y, 16);
                        if (lane == 0) {
                            float2 x1c, x2c;
                            bfly_p(e128, r128, x1c, x2c);               // W^128 = s (1 - i)
                            float* const pcol = s_pow + 2 * f;
                            pcol[64 * PROW] = x1c.x * x1c.x + x1c.y * x1c.y;
                            pcol[192 * PROW] = x2c.x * x2c.x + x2c.y * x2c.y;
                        }
                    }
                }
                // ---- the tile's mel bands, by all sixteen warps: lane & 15 = frame, the two halves of a warp take
                // neighbouring bands (pairs pp = warp, warp + 16, ...); broadcast weights, conflict-free powers ------
                mbar_wait(bar_tile_empty + pb, ((it / NTILE) & 1) ^ 1);       // the epilogue warps are done with this tile slot
                frame_warps_sync();                                           // every frame's powers are in the tile
                {
                    const int l16 = lane & 15, par = lane >> 4;
                    float* const trow = s_tile + pb * (rows * F) + l16;        // element (row, frame l16)
                    float* const dbt = s_dbc + (l16 >> 2) * dbc_quad(n_mels) + (l16 & 3);   // mfcc: (quad, band 0, frame)
                    const float2* pl = reinterpret_cast<const float2*>(s_pow) + l16;
                    for (int pp = warp; 2 * pp < n_mels; pp += kFftWarps) {
                        const int i = 2 * pp + par;
                        const int4 d = s_desc[min(i, n_mels - 1)];             // (an odd last band pairs with itself)
                        const float2* pr = pl + d.x;
                        const float4* wq = reinterpret_cast<const float4*>(s_melw + d.z);
                        float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll 4
                        for (int q4 = 0; q4 < d.y; ++q4) {                      // d.y is the pair's: api.cu pads the narrower band
                            const float4 w = wq[q4];
                            const float2 p0 = pr[0], p1 = pr[PROW / 2];
                            a01 = __ffma2_rn(make_float2(w.x, w.y), p0, a01);
                            a23 = __ffma2_rn(make_float2(w.z, w.w), p1, a23);
                            pr += PROW;
                        }
                        if (i < n_mels) {
                            const float vv = db10((a01.x + a01.y) + (a23.x + a23.y));
                            trow[d.w * F] = vv;
                            if (MFCC) dbt[4 * d.w] = vv;
                        }
                    }
                }
                frame_warps_sync();                                           // all dB written; the power tile may be refilled
                if constexpr (MFCC) {
                    // ---- DCT-II of the tile as a 16 frames x n_mfcc x n_mels product, assuming the top_db clip (known
                    // only after the clip's last tile) will not engage; checked at clip end.  A warp takes four
                    // coefficients at a time (kc = warp, warp + 16, ...); lane >> 3 = frame quad, lane & 7 strides the
                    // bands, so a lane accumulates a 4 frames x 4 coefficients block from two conflict-free 128-bit
                    // loads per band (16 FMA per 32 bytes of shared memory), and the eight band lanes are folded
                    // with a transposing shuffle reduction (14 shuffles for 16 sums).
                    const int fq = lane >> 3, ml = lane & 7;
                    const float4* const d4 = reinterpret_cast<const float4*>(s_dbc + fq * dbc_quad(n_mels));
                    float* const tq = s_tile + pb * (rows * F) + n_mels * F + 4 * fq + (ml >> 1);
                    for (int kc = warp; 4 * kc < p.n_mfcc; kc += kFftWarps) {
                        const float4* const b4 = reinterpret_cast<const float4*>(s_dct) + kc * n_mels;
                        float2 acc[8];                                       // [frame of the quad][coefficient pair]
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll 4
                        for (int m = ml; m < n_mels; m += 8) {
                            const float4 dv = d4[m], bv = b4[m];
                            const float2 b01 = make_float2(bv.x, bv.y), b23 = make_float2(bv.z, bv.w);
                            acc[0] = __ffma2_rn(bc2(dv.x), b01, acc[0]); acc[1] = __ffma2_rn(bc2(dv.x), b23, acc[1]);
                            acc[2] = __ffma2_rn(bc2(dv.y), b01, acc[2]); acc[3] = __ffma2_rn(bc2(dv.y), b23, acc[3]);
                            acc[4] = __ffma2_rn(bc2(dv.z), b01, acc[4]); acc[5] = __ffma2_rn(bc2(dv.z), b23, acc[5]);
                            acc[6] = __ffma2_rn(bc2(dv.w), b01, acc[6]); acc[7] = __ffma2_rn(bc2(dv.w), b23, acc[7]);
                        }
                        float a[16];                                         // a[4 f + k]
#pragma unroll
                        for (int i = 0; i < 8; ++i) { a[2 * i] = acc[i].x; a[2 * i + 1] = acc[i].y; }
                        const bool h4 = ml & 4, h2 = ml & 2, h1 = ml & 1;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float keep = h4 ? a[j + 8] : a[j], send = h4 ? a[j] : a[j + 8];
                            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float keep = h2 ? a[j + 4] : a[j], send = h2 ? a[j] : a[j + 4];
                            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                        }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float keep = h1 ? a[j + 2] : a[j], send = h1 ? a[j] : a[j + 2];
                            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                        // a[j] is now the full sum of block element j + 2 ml: frame ml >> 1, coefficient 2 (ml & 1) + j
                        const int k = 4 * kc + 2 * (ml & 1);
                        if (k < p.n_mfcc) tq[k * F] = a[0];
                        if (k + 1 < p.n_mfcc) tq[(k + 1) * F] = a[1];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tile_full + pb);
            }
        }
    } else {
        // =========================== epilogue warps (warp 0 of them also stages the raw tiles) =========
        const int mw = warp - kFftWarps, mtid = tid - 32 * kFftWarps;
        const int ef = mtid & 15, er0 = mtid >> 4;               // this thread's frame lane and first row of a tile (rows er0 + 8 i)
        constexpr int V = 16 / (int)sizeof(E);
        const bool base_aligned = (reinterpret_cast<uintptr_t>(p.clips) & 15) == 0;

        auto stage = [&](long long clip, int t0, uint32_t slot) {
            E* const dst = reinterpret_cast<E*>(smem + L.off_raw + slot * L.raw_bytes);
            uint64_t* const bar = bar_raw_full + slot;
            const int n = RAG ? p.rag_len[clip] : p.n_samples;
            const long long e0 = RAG ? p.rag_in_off[clip] : clip * (long long)n;
            const E* cptr = reinterpret_cast<const E*>(p.clips) + e0;
            const int c0 = t0 * hop - NFFT / 2;
            const int lo = c0 < 0 ? 0 : c0;
            const int hi = (c0 + chunk < n) ? c0 + chunk : n;
            const int nfr = 1 + n / hop;
            const int vf = nfr - t0 < F ? nfr - t0 : F;
            const int need = ((vf - 1) * hop + NFFT + V - 1) & ~(V - 1);      // <= chunk
            // The clip's own 16-byte grid decides where the tile sits in the slot: tile sample c0 + i goes to
            // dst[i + sft], sft < V chosen so that the first sample on a 16-byte boundary of GLOBAL memory lands on a
            // 16-byte boundary of the slot (110 250-sample int16 clips start 0 / 4 / 8 / 12 bytes off the grid).
            // The aligned interior is one TMA bulk copy; up to V - 1 samples on either side, the centre padding
            // and whatever the valid frames read past the clip go by plain stores.
            const int mis = (int)((e0 + lo) & (V - 1));
            const int lead = (V - mis) & (V - 1);
            const int lo_a = lo + lead;                                       // first sample of the bulk copy
            const bool ok = p.pad_mode == 0 && base_aligned && hi > lo && (!I16 || (mis & 1) == 0);
            if (!ok) {
                for (int i = lane; i < need; i += 32) dst[i] = raw_sample<E>(cptr, c0 + i, n, p.pad_mode);
                __syncwarp();
                if (lane == 0) { s_sft[slot] = 0; mbar_arrive(bar); }
                return;
            }
            const int sft = (V - ((lo_a - c0) & (V - 1))) & (V - 1);
            const int nb = hi > lo_a ? ((hi - lo_a) / V) * V : 0;            // bulk part, whole 16-byte units
            const int d_lo = lo - c0 + sft;                                   // slot index of sample lo
            const int d_a = lo_a - c0 + sft;                                  // ... of the bulk copy (multiple of V)
            const int d_tb = d_a + nb;                                        // ... of the first sample behind it
            const int d_end = (need + sft + V - 1) & ~(V - 1);                // <= chunk + V
            const int4 z4 = make_int4(0, 0, 0, 0);
            // zeros: [0, d_a) (centre padding + the 16 bytes the lead samples share) and [d_tb, d_end)
            for (int i = lane * V; i < d_a; i += 32 * V) *reinterpret_cast<int4*>(dst + i) = z4;
            for (int i = d_tb + lane * V; i < d_end; i += 32 * V) *reinterpret_cast<int4*>(dst + i) = z4;
            __syncwarp();
            if (lane < lead && lo + lane < hi) dst[d_lo + lane] = cptr[lo + lane];
            if (lane < V && lo_a + nb + lane < hi && d_tb + lane < d_end) dst[d_tb + lane] = cptr[lo_a + nb + lane];
            __syncwarp();
            if (lane == 0) {
                s_sft[slot] = sft;
                if (nb > 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic accesses -> async write
                    const int room = d_end - d_a;
                    const int cb = (nb < room ? nb : room) * (int)sizeof(E);          // never more than the frames read
                    mbar_expect_tx(bar, (uint32_t)cb);
                    bulk_g2s(dst + d_a, cptr + lo_a, (uint32_t)cb, bar);
                } else {
                    mbar_arrive(bar);
                }
            }
        };
        long long pclip = blockIdx.x;
        int ptile = 0;
        uint32_t pit = 0;
        auto stage_next = [&]() {
            if (pclip >= p.n_clips) return;
            stage(pclip, ptile * F, pit % NRAW);
            ++pit;
            const int pnfr = RAG ? 1 + p.rag_len[pclip] / hop : p.n_frames;
            if (++ptile * F >= pnfr) { ptile = 0; pclip += gridDim.x; }
        };
        if (mw == 0)
            for (int i = 0; i < NRAW; ++i) stage_next();

        // Deferred rewrite of the PREVIOUS clip during the current clip's tiles (prefetched before the wait for
        // the tile): mel = power_to_db(ref=max, top_db) + min-max on float4 slices; mfcc = per-row z-score.
        constexpr int NPF = 4;                                   // float4 (mel) / floats (mfcc) per thread per tile
        float4* nq = nullptr;
        float* nz = nullptr;
        int nq_n = 0, nq_done = 0;
        float nq_vmax = 0.f, nq_lo = 0.f, nq_range = 1.f, nq_inv = 1.f, nz_inv = 1.f;
        auto nrm = [&](float x) {
            const float num = fmaxf(x - nq_vmax, -p.top_db) - nq_lo;
            const float q = num * nq_inv;
            return fmaf(fmaf(-q, nq_range, num), nq_inv, q);     // correctly rounded x / range (peak exactly 1.0)
        };
        auto nrm4 = [&](float4 x) { return make_float4(nrm(x.x), nrm(x.y), nrm(x.z), nrm(x.w)); };
        auto zs1 = [&](float x, int i) {
            const int k = __float2int_rd(((float)i + 0.5f) * nz_inv);
            return (x - s_zs[2 * k]) * s_zs[2 * k + 1];
        };
        auto nq_finish = [&]() {
            if constexpr (!MFCC) {
                for (int i = nq_done + mtid; i < nq_n; i += kMelThreads) nq[i] = nrm4(nq[i]);
            } else {
                for (int i = nq_done + mtid; i < nq_n; i += kMelThreads) nz[i] = zs1(nz[i], i);
            }
            nq_done = nq_n;
        };
        // mfcc: this thread owns rows er0 + 8 i of every tile at frame lane ef: running sums about the row's first value
        constexpr int kZR = 8;                                   // n_mfcc <= 64
        float zS[kZR], zQ[kZR], zx0[kZR];

        uint32_t it = 0;
        for (long long clip = blockIdx.x; clip < p.n_clips; clip += gridDim.x) {
            const int nfr = RAG ? 1 + p.rag_len[clip] / hop : p.n_frames;
            const int tiles = (nfr + F - 1) / F;
            float* const outb = RAG ? p.out + p.rag_out_off[clip]
                                    : p.out + (size_t)clip * (MFCC ? p.n_mfcc : n_mels) * nfr;
            float* const inter = MFCC ? p.inter + (size_t)blockIdx.x * n_mels * p.n_frames : outb;
            float vmax = -3.0e38f, vmin = 3.0e38f;
#pragma unroll
            for (int i = 0; i < kZR; ++i) { zS[i] = 0.f; zQ[i] = 0.f; zx0[i] = 0.f; }

            for (int tile = 0; tile < tiles; ++tile, ++it) {
                const int t0 = tile * F;
                const uint32_t pb = it % NTILE;
                float4 nx[MFCC ? 1 : NPF];
                float zx[MFCC ? NPF : 1];
                const bool nq_live = nq_done < nq_n;
                if (nq_live) {
#pragma unroll
                    for (int k = 0; k < NPF; ++k) {
                        const int i = nq_done + mtid + k * kMelThreads;
                        if (i < nq_n) {
                            if constexpr (!MFCC) nx[k] = nq[i]; else zx[k] = nz[i];
                        }
                    }
                }
                mbar_wait(bar_tile_full + pb, (it / NTILE) & 1);
                if (mw == 0) stage_next();                         // raw slot it % NRAW is free again
                // ---- tile -> global: 16 consecutive frames of a row per half-warp ---------------------------
                {
                    const int t = t0 + ef;
                    const bool valid = t < nfr;
                    const float* tl = s_tile + pb * (rows * F) + ef;
                    for (int r = er0; r < n_mels; r += 8) {            // raw dB rows
                        const float vv = tl[r * F];
                        if (valid) {
                            inter[(size_t)r * nfr + t] = vv;
                            vmax = fmaxf(vmax, vv);
                            vmin = fminf(vmin, vv);
                        }
                    }
                    if constexpr (MFCC) {
#pragma unroll
                        for (int i = 0; i < kZR; ++i) {
                            const int k = er0 + 8 * i;
                            const float a = k < p.n_mfcc ? tl[(n_mels + k) * F] : 0.f;
                            const float first = __shfl_sync(0xffffffffu, a, lane & 16);           // (outside the per-half condition)
                            if (k < p.n_mfcc) {
                                if (valid) outb[(size_t)k * nfr + t] = a;
                                if (tile == 0) zx0[i] = first;                                    // frame 0 of the clip
                                const float dd = valid ? a - zx0[i] : 0.f;
                                zS[i] += dd;
                                zQ[i] = fmaf(dd, dd, zQ[i]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tile_empty + pb);
                if (nq_live) {
#pragma unroll
                    for (int k = 0; k < NPF; ++k) {
                        const int i = nq_done + mtid + k * kMelThreads;
                        if (i < nq_n) {
                            if constexpr (!MFCC) nq[i] = nrm4(nx[k]); else nz[i] = zs1(zx[k], i);
                        }
                    }
                    nq_done += NPF * kMelThreads;
                    // whatever a tile's share exceeds the prefetched slices goes without prefetch
                    const int per_tile = MFCC ? p.n_mfcc * F : n_mels * (F / 4);
                    for (int more = (per_tile + kMelThreads - 1) / kMelThreads - NPF; more > 0; --more) {
                        const int i = nq_done + mtid;
                        if (i < nq_n) {
                            if constexpr (!MFCC) nq[i] = nrm4(nq[i]); else nz[i] = zs1(nz[i], i);
                        }
                        nq_done += kMelThreads;
                    }
                    if (tile + 1 == tiles) nq_finish();
                }
            }

            // ---- per-clip reductions (epilogue warps only) --------------------------------------------
            vmax = warp_max(vmax); vmin = warp_min(vmin);
            if (lane == 0) { s_red[mw] = vmax; s_red[32 + mw] = vmin; }
            mel_sync();                                            // also: every warp's stores of this clip are visible
            {
                const float a = (lane < kMelWarps) ? s_red[lane] : -3.0e38f;
                const float b = (lane < kMelWarps) ? s_red[32 + lane] : 3.0e38f;
                vmax = warp_max(a); vmin = warp_min(b);
            }
            if constexpr (!MFCC) {
                nq_finish();
                nq_vmax = vmax;
                nq_lo = fmaxf(vmin - vmax, -p.top_db);
                nq_range = (0.0f - nq_lo) + 1e-8f;
                nq_inv = __frcp_rn(nq_range);
                const int total = n_mels * nfr;
                if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(inter) & 15) == 0)) {
                    nq = reinterpret_cast<float4*>(inter);
                    nq_n = total / 4;
                    nq_done = 0;
                } else {
                    for (int i = mtid; i < total; i += kMelThreads) inter[i] = nrm(inter[i]);
                    nq_n = nq_done = 0;
                }
            } else {
                float* outc = outb;
                const float thr = vmax - p.top_db;
                const bool clipped = vmin < thr;
                nq_finish();                                          // the previous clip's z-score, if any is left
                mel_sync();                                           // ... by every warp, before s_zs changes
                const float fn = (float)nfr;
                if (!clipped) {
                    // mean and variance from the running sums of the 16 frame lanes (a half-warp holds one row set)
#pragma unroll
                    for (int i = 0; i < kZR; ++i) {
                        float S = zS[i], Q = zQ[i];
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) {
                            S += __shfl_xor_sync(0xffffffffu, S, o);
                            Q += __shfl_xor_sync(0xffffffffu, Q, o);
                        }
                        const int k = er0 + 8 * i;
                        if (ef == 0 && k < p.n_mfcc) {
                            const float md = __fdiv_rn(S, fn);
                            s_zs[2 * k] = zx0[i] + md;
                            s_zs[2 * k + 1] = __fdiv_rn(1.0f, sqrtf(fmaxf(fmaf(-md, md, __fdiv_rn(Q, fn)), 0.f)) + 1e-8f);
                        }
                    }
                    nz = outc;
                    nz_inv = __fdiv_rn(1.0f, fn);
                    nq_n = p.n_mfcc * nfr;
                    nq_done = 0;
                    mel_sync();
                    continue;
                }
                // rare: a band fell more than top_db below the clip's peak, so the clipped dB differ from what the
                // frame warps' DCT saw -> recompute from the raw-dB scratch (L2 resident), then the three-sweep z-score
                nq_n = nq_done = 0;
                for (int i = mtid; i < p.n_mfcc * nfr; i += kMelThreads) {
                    const int k = i / nfr, t = i - k * nfr;
                    const float* bk = p.dct + (size_t)k * n_mels;
                    float acc = 0.f;
                    for (int m = 0; m < n_mels; ++m) acc = fmaf(__ldg(bk + m), fmaxf(inter[(size_t)m * nfr + t], thr), acc);
                    outc[i] = acc;
                }
                mel_sync();
                for (int k = mw; k < p.n_mfcc; k += kMelWarps) {      // deep.py:326-328, three sweeps over an L2-resident row
                    float* row = outc + (size_t)k * nfr;
                    const float x0 = row[0];
                    float s_ = 0.f;
                    for (int t = lane; t < nfr; t += 32) s_ += row[t] - x0;
                    const float mean = x0 + __fdiv_rn(warp_sum(s_), fn);
                    float ss = 0.f;
                    for (int t = lane; t < nfr; t += 32) { const float dlt = row[t] - mean; ss = fmaf(dlt, dlt, ss); }
                    const float sd = sqrtf(__fdiv_rn(warp_sum(ss), fn)) + 1e-8f;
                    for (int t = lane; t < nfr; t += 32) row[t] = __fdiv_rn(row[t] - mean, sd);
                }
            }
            mel_sync();
        }
        nq_finish();
    }
}

}  // namespace

size_t logmel1024_smem_bytes(int hop, int n_mels, int mel_wpad, bool i16, int n_mfcc) {
    return (size_t)make_layout(hop, n_mels, mel_wpad, i16, n_mfcc).total + 128;
}

int logmel1024_pow_rows() { return PROWS; }

bool logmel1024_supports(int hop, int n_mels, int n_mfcc) {
    return hop > 0 && (hop % 4) == 0 && n_mels <= 512 && n_mfcc <= 32 * kMaxCoefPerLane &&
           (n_mfcc == 0 || (n_mels % 4) == 0);
}

template <bool I16, int KIND, bool RAG>
static cudaError_t launch_k(const FrontParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = logmel1024_kernel<I16, KIND, RAG>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_logmel1024(const FrontParams& p, bool i16, int kind, int grid, cudaStream_t st) {
    const size_t smem = logmel1024_smem_bytes(p.hop, p.n_mels, p.mel_wpad, i16, kind == 1 ? p.n_mfcc : 0);
    if (p.rag_len) {
        if (kind == 0) return i16 ? launch_k<true, 0, true>(p, grid, smem, st) : launch_k<false, 0, true>(p, grid, smem, st);
        return i16 ? launch_k<true, 1, true>(p, grid, smem, st) : launch_k<false, 1, true>(p, grid, smem, st);
    }
    if (kind == 0) return i16 ? launch_k<true, 0, false>(p, grid, smem, st) : launch_k<false, 0, false>(p, grid, smem, st);
    return i16 ? launch_k<true, 1, false>(p, grid, smem, st) : launch_k<false, 1, false>(p, grid, smem, st);
}

}  // namespace b2a
