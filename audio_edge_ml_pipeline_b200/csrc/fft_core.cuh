// fft_core.cuh — register-resident radix-2/4/8/16 butterflies and the per-frame real FFT.
//
// A real frame of n_fft = 2*NC samples is packed as NC complex points z[n] = x[2n] + i x[2n+1],
// transformed by a Stockham autosort FFT whose passes keep 16 complex points per thread in
// registers (T = NC/16 threads per frame, passes exchange through padded shared memory), and
// un-packed by the usual split step X[k] = E[k] + W_{2NC}^k O[k].
//
//   NC = 128 : radix 16, 8          (n_fft  256, CQT)
//   NC = 256 : radix 16, 16         (n_fft  512, the headline log-mel configuration)
//   NC = 512 : radix 16, 16, 2      (n_fft 1024)
//   NC = 1024: radix 16, 16, 4      (n_fft 2048)
//
// Replaces scipy.fft.rfft inside librosa.stft (reference call site deep.py:126-132).
#pragma once
#include <cuda_runtime.h>

#ifndef B2A_AB_EXACT_BFLY
#define B2A_AB_EXACT_BFLY 0
#endif

namespace b2a {

// ---- packed FP32 (sm_100a FADD2 / FMUL2 / FFMA2) -----------------------------------------------
// A complex point is a float2 in an aligned register pair.  Blackwell's packed-FP32 instructions
// work on such pairs with free per-half negation, half swap (.LO_HI) and scalar broadcast operand
// modifiers, so a complex add is ONE instruction and a complex multiply TWO (FMUL2 + FFMA2):
// the same FMA-pipe cycles as the scalar form but half the issue slots, which is what bounds
// these kernels (tools/ubench/fp32_issue.cu: FFMA2 issues every 2 cycles per scheduler, scalar
// FFMA every cycle).  ptxas folds the make_float2(...) permutations below into operand modifiers.
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (a.x b.x - a.y b.y, a.y b.x + a.x b.y) = a * (b.x, b.x) + swap(a) * (-b.y, b.y)
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), __fmul2_rn(a, bc2(b.x)));
}
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

#define B2A_SQRT1_2 0.70710678118654752440f
#define B2A_COS_PI_8 0.92387953251128675613f
#define B2A_SIN_PI_8 0.38268343236508977173f

// Fused twiddle butterflies: (e + w*o, e - w*o) in 3 FFMA2.  The second output is formed as
// 2e - first (one extra rounding, ~1 ulp).
__device__ __forceinline__ void bfly_w(float2 e, float2 o, float wr, float wi, float2& a, float2& b) {
    a = __ffma2_rn(make_float2(o.y, o.x), make_float2(-wi, wi), __ffma2_rn(o, bc2(wr), e));
#if B2A_AB_EXACT_BFLY
    b = __ffma2_rn(make_float2(o.y, o.x), make_float2(wi, -wi), __ffma2_rn(o, bc2(-wr), e));
#else
    b = __ffma2_rn(e, bc2(2.0f), make_float2(-a.x, -a.y));
#endif
}
// w = s(1 - i):  w*o = s((o.x + o.y), (o.y - o.x))
__device__ __forceinline__ void bfly_p(float2 e, float2 o, float2& a, float2& b) {
    const float2 t = __fadd2_rn(o, make_float2(o.y, -o.x));
    a = __ffma2_rn(t, bc2(B2A_SQRT1_2), e);
    b = __ffma2_rn(t, bc2(-B2A_SQRT1_2), e);
}
// w = s(-1 - i): w*o = s((o.y - o.x), -(o.x + o.y))
__device__ __forceinline__ void bfly_m(float2 e, float2 o, float2& a, float2& b) {
    const float2 t = __fadd2_rn(make_float2(o.y, -o.x), make_float2(-o.x, -o.y));
    a = __ffma2_rn(t, bc2(B2A_SQRT1_2), e);
    b = __ffma2_rn(t, bc2(-B2A_SQRT1_2), e);
}

// In-place forward DFT (e^{-2 pi i nk/N}), natural order in and out.
template <int N> struct Dft;

template <> struct Dft<1> {
    static __device__ __forceinline__ void run(float2*) {}
};
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(float2* v) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    static __device__ __forceinline__ void run(float2* v) {
        const float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        const float2 t2 = cadd(v[1], v[3]), t3 = cmul_mi(csub(v[1], v[3]));
        v[0] = cadd(t0, t2); v[2] = csub(t0, t2);
        v[1] = cadd(t1, t3); v[3] = csub(t1, t3);
    }
};
template <> struct Dft<8> {
    static __device__ __forceinline__ void run(float2* v) {
        float2 e[4] = {v[0], v[2], v[4], v[6]};
        float2 o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e); Dft<4>::run(o);
        // W8^1 = s(1 - i), W8^2 = -i, W8^3 = s(-1 - i)
        const float2 o2 = cmul_mi(o[2]);
        v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
        bfly_p(e[1], o[1], v[1], v[5]);
        v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
        bfly_m(e[3], o[3], v[3], v[7]);
    }
};
template <> struct Dft<16> {
    static __device__ __forceinline__ void run(float2* v) {
        float2 e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<8>::run(e); Dft<8>::run(o);
        const float c1 = B2A_COS_PI_8, s1 = B2A_SIN_PI_8;
        // W16^k = (cos(k pi/8), -sin(k pi/8))
        const float2 t4 = cmul_mi(o[4]);
        v[0] = cadd(e[0], o[0]); v[8]  = csub(e[0], o[0]);
        bfly_w(e[1], o[1], c1, -s1, v[1], v[9]);
        bfly_p(e[2], o[2], v[2], v[10]);
        bfly_w(e[3], o[3], s1, -c1, v[3], v[11]);
        v[4] = cadd(e[4], t4);   v[12] = csub(e[4], t4);
        bfly_w(e[5], o[5], -s1, -c1, v[5], v[13]);
        bfly_m(e[6], o[6], v[6], v[14]);
        bfly_w(e[7], o[7], -c1, -s1, v[7], v[15]);
    }
};

// Exchange-buffer padding: one float2 of padding after every 16 keeps the stride-16 (pass 1
// store) and unit-stride (pass 2 load) patterns conflict-free for 64-bit accesses.
__device__ __forceinline__ int xpad(int i) { return i + (i >> 4); }

template <int LOG2NC> struct FftGeom {
    static constexpr int NC = 1 << LOG2NC;          // complex points
    static constexpr int NFFT = 2 * NC;             // real frame length
    static constexpr int T = NC / 16;               // threads per frame
    static constexpr int R1 = (LOG2NC >= 8) ? 16 : (1 << (LOG2NC - 4));
    static constexpr int R2 = (LOG2NC > 8) ? (1 << (LOG2NC - 8)) : 1;
    static constexpr int XSTRIDE = NC + NC / 16;    // float2 per frame slot in the exchange buffer
    static constexpr int PSTRIDE = NC + 1;          // floats per frame in the power tile (odd)
    static_assert(LOG2NC >= 7 && LOG2NC <= 10, "n_fft must be 256..2048");
};

// Synchronise the T threads that share one frame.
template <int T> __device__ __forceinline__ void frame_sync() {
    if constexpr (T <= 32) __syncwarp(); else __syncthreads();
}

// Pass-2 twiddles re-ordered per thread: entry [(u*(R1-1) + (t-1))*T + j] = exp(-2 pi i t k / (16 R1))
// with k = (j + T u) & 15.  Lanes read consecutive entries (a lookup into the natural-order table is
// a stride-t access across lanes: up to 16-way bank conflicts in the ncu source view).
template <int LOG2NC> struct FftTwp {
    using G = FftGeom<LOG2NC>;
    static constexpr int NB = 16 / G::R1;
    static constexpr int SIZE = NB * (G::R1 - 1) * G::T;       // float2 entries
    static __device__ __forceinline__ void fill(float2* __restrict__ s_twp, const float2* __restrict__ tw_global,
                                                int tid, int nthreads) {
        constexpr int TWS = G::NC / (16 * G::R1);
        for (int i = tid; i < SIZE; i += nthreads) {
            const int j = i % G::T, r = i / G::T;
            const int u = r / (G::R1 - 1), t = r % (G::R1 - 1) + 1;
            const int k = (j + G::T * u) & 15;
            s_twp[i] = tw_global[(t * k) * TWS];
        }
    }
};

// Passes 2.. of the packed FFT on one frame.  `xb` is this frame's exchange slot holding the
// output of pass 1 (index 16*j + t, padded); on return it holds Z[0..NC) in natural order.
// `tw` = exp(-2 pi i k / NC), k in [0, NC), in shared memory.  `tw1` (optional) = this thread's
// pass-2 twiddles hoisted to registers by the caller when R1 == 16 && 16/R1 == 1.
template <int LOG2NC, bool HOISTED>
__device__ __forceinline__ void fft_tail_passes(float2* __restrict__ xb, const float2* __restrict__ tw,
                                                const float2* tw1, const float2* __restrict__ twp, int j) {
    using G = FftGeom<LOG2NC>;
    constexpr int NC = G::NC, T = G::T, R1 = G::R1, R2 = G::R2;
    float2 v[16];
    // ---- pass 2: radix R1, Ns = 16 --------------------------------------------------------
    {
        constexpr int NB = 16 / R1;                 // butterflies per thread
        constexpr int STR = NC / R1;                // input stride
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int jb = j + T * u;
            const int k = jb & 15;
#pragma unroll
            for (int t = 0; t < R1; ++t) {
                float2 x = xb[xpad(jb + t * STR)];
                if (t > 0) {
                    const float2 w = HOISTED ? tw1[t - 1] : twp[(u * (R1 - 1) + (t - 1)) * T + j];
                    x = cmul(x, w);
                }
                v[u * R1 + t] = x;
            }
            Dft<R1>::run(&v[u * R1]);
        }
        frame_sync<T>();
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int jb = j + T * u;
            const int k = jb & 15;
            const int base = (jb - k) * R1 + k;
#pragma unroll
            for (int t = 0; t < R1; ++t) xb[xpad(base + t * 16)] = v[u * R1 + t];
        }
        frame_sync<T>();
    }
    // ---- pass 3: radix R2, Ns = 256 -------------------------------------------------------
    if constexpr (R2 > 1) {
        constexpr int NB = 16 / R2;
        constexpr int STR = NC / R2;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int jb = j + T * u;               // in [0, NC/R2) = [0, 256)
            const int k = jb & 255;
#pragma unroll
            for (int t = 0; t < R2; ++t) {
                float2 x = xb[xpad(jb + t * STR)];
                if (t > 0) x = cmul(x, tw[t * k]);  // NC/(256*R2) == 1
                v[u * R2 + t] = x;
            }
            Dft<R2>::run(&v[u * R2]);
        }
        frame_sync<T>();
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int jb = j + T * u;
            const int k = jb & 255;
            const int base = (jb - k) * R2 + k;
#pragma unroll
            for (int t = 0; t < R2; ++t) xb[xpad(base + t * 256)] = v[u * R2 + t];
        }
        frame_sync<T>();
    }
}

// Split step for one (k, NC-k) pair of the packed transform.
//   A = Z[k], Bz = Z[NC-k] (Z[NC] == Z[0]), w = exp(-i pi k / NC)
//   returns 2*X[k] in xk and 2*X[NC-k] in xnk  (the factor 2 is removed by the caller)
__device__ __forceinline__ void rfft_split(float2 A, float2 Bz, float2 w, float2& xk, float2& xnk) {
    const float2 E = __fadd2_rn(A, make_float2(Bz.x, -Bz.y));
    const float2 O = __fadd2_rn(make_float2(A.y, -A.x), make_float2(Bz.y, Bz.x));
    xk = __ffma2_rn(make_float2(O.y, O.x), make_float2(-w.y, w.y), __ffma2_rn(O, bc2(w.x), E));   // E + w*O
#if B2A_AB_EXACT_BFLY
    {   // conj(E - w*O), each half from its own FMA chain
        const float2 d = __ffma2_rn(make_float2(O.y, O.x), make_float2(w.y, -w.y), __ffma2_rn(O, bc2(-w.x), E));
        xnk = make_float2(d.x, -d.y);
    }
#else
    xnk = __ffma2_rn(make_float2(E.x, -E.y), bc2(2.0f), make_float2(-xk.x, xk.y));   // conj(E - w*O) = conj(2E - xk)
#endif
}

}  // namespace b2a
