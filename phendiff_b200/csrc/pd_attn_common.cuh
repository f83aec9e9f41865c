// phendiff_b200 — exponent evaluation shared by the attention kernels (pd_attn_mma.cu, pd_attn_tc.cu).
#pragma once
#include "pd_kernels.h"
#include <type_traits>

namespace pd {

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial,
// scaled by 2^n through the exponent field
__device__ __forceinline__ float ex2_poly(float x) {
    x = fminf(fmaxf(x, -100.0f), 128.0f);   // 128 -> exponent field 255 = +inf (overflow of a stale row max stays visible)
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = 0.05508868396282196f;
    p = fmaf(p, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t h2op_add(uint32_t a, uint32_t b) { uint32_t d; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2op_sub(uint32_t a, uint32_t b) { uint32_t d; asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2op_mul(uint32_t a, uint32_t b) { uint32_t d; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2op_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t h2op_min(uint32_t a, uint32_t b) { uint32_t d; asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t h2op_max(uint32_t a, uint32_t b) { uint32_t d; asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

// (2^xa, 2^xb) as packed fp16 (low half = xa), on the FMA / ALU pipes.  x is clamped to [-15, 16]: n = round(x) rides in
// the low mantissa bits of w = x + 1039 (ulp 1 in [1024, 2048)) as 15 + n = the fp16 exponent field of 2^n, so
// x <= -14.5 gives exactly 0 and x >= 15.5 gives +inf (which is what flags the overflow of a stale row max).
__device__ __forceinline__ uint32_t ex2_pair_h2(float xa, float xb) {
    uint32_t x = pack_f16x2(xa, xb);
    x = h2op_min(h2op_max(x, 0xCB80CB80u /* -15 */), 0x4C004C00u /* 16 */);
    const uint32_t w = h2op_add(x, 0x640F640Fu /* 1039 */);
    const uint32_t f = h2op_sub(x, h2op_sub(w, 0x640F640Fu));          // x - n in [-0.5, 0.5], exact
    uint32_t p = h2op_fma(0x2B0D2B0Du /* 0.05508868 */, f, 0x33C333C3u /* 0.24260405 */);
    p = h2op_fma(p, f, 0x398C398Cu /* 0.69327623 */);
    p = h2op_fma(p, f, 0x3C003C00u /* 0.99992895 -> 1 */);
    // exponent fields (15 + n) << 10 of both halves in one IMAD: (w - 0x64006400) * 1024, the subtraction folded into the addend
    return h2op_mul(p, w * 1024u + (0u - 0x64006400u * 1024u));
}

template <typename T> __device__ __forceinline__ uint32_t ex2_pair_poly(float xa, float xb);
template <> __device__ __forceinline__ uint32_t ex2_pair_poly<f16>(float xa, float xb) { return ex2_pair_h2(xa, xb); }
template <> __device__ __forceinline__ uint32_t ex2_pair_poly<bf16>(float xa, float xb) { return pack_bf16x2(ex2_poly(xa), ex2_poly(xb)); }


}  // namespace pd
