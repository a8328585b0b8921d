// ws_common.cuh — device helpers shared by the warp-specialised front-end kernels (logmel512.cu,
// logmel1024.cu): dB, PCM widening, warp reductions, mbarrier / TMA bulk-copy wrappers, padded sample reads.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#ifndef B2A_AB_LOG2F
#define B2A_AB_LOG2F 0
#endif
#ifndef B2A_MBAR_HINT
#define B2A_MBAR_HINT 100000
#endif

namespace b2a {
namespace ws {

__device__ __forceinline__ float db10(float s) {
    // 10*log10(max(amin, s)); the argument is >= 1e-10, never denormal -> lg2.approx.ftz
#if B2A_AB_LOG2F
    return 3.01029995663981195f * log2f(fmaxf(s, 1e-10f));
#else
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fmaxf(s, 1e-10f)));
    return 3.01029995663981195f * l;
#endif
}
// two packed int16 PCM samples -> two integer-valued floats (the 1/32768 rides on the window).
// Sign-extend on the ALU (PRMT / SHF) then I2FP.F32.S32: cvt.f32.s16 (I2F.S16) runs on the
// quarter-rate XU pipe and cost 4.5x an FADD per instruction in the ncu source view.
__device__ __forceinline__ float2 cvt_pcm2(uint32_t u) {
    int lo;                                                  // bytes {b0, b1, sign(b1), sign(b1)}: selector
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(lo) : "r"(u));   // msb = replicate sign (PTX prmt; the
                                                             // __byte_perm intrinsic ignores that bit)
    const int hi = (int)u >> 16;
    return make_float2(__int2float_rn(lo), __int2float_rn(hi));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if B2A_MBAR_HINT
    // with a suspend-time hint (ns): the waiting warp stays parked until the phase completes or the hint
    // expires instead of re-issuing the try_wait / branch pair every few cycles next to the working warps
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)B2A_MBAR_HINT) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// sample s of a clip in its storage type; zero (or the reflected sample) outside [0, n)
template <typename E>
__device__ __forceinline__ E raw_sample(const E* clip, int s, int n, int pad_mode) {
    if (s < 0 || s >= n) {
        if (pad_mode == 0) return (E)0;
        s = (s < 0) ? -s : 2 * (n - 1) - s;          // np.pad(mode="reflect")
        if (s < 0 || s >= n) return (E)0;
    }
    return clip[s];
}

}  // namespace ws
}  // namespace b2a
