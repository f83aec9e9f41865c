// phendiff_b200 — fused flash-style self-attention for head_dim 8 (diffusers Attention + AttnProcessor2_0 core:
// softmax(q k^T / sqrt(8)) v), 16-bit in / fp32 softmax statistics / 16-bit out, on packed qkv (N, S, 3C).
//
// Shape note: d = 8 is below the tcgen05 16-bit K step (16) and P.V has N = 8; the matmuls are 4.5 % of the forward
// FLOPs while the S^2 exponentials dominate: per image-forward 6 x 64 heads x 1024^2 = 403 M exp against 16 MUFU
// results / clk / SM (measured, profiles/r1b_halo_probe.md) = 0.089 ms at 148 SMs.  The kernel is therefore organised
// around the exp budget, not the tensor pipe:
//   * warp-level tensor-core shapes fit d = 8 exactly (QK^T: m16n8k8, P.V: m16n8k16 with the S accumulator fragment
//     re-used as the A operand), so S and P never leave registers;
//   * the softmax denominator is one more m16n8k16 against a ones operand instead of 32 FADDs per lane and step;
//   * K and V^T of the head sit in shared memory pre-arranged in FRAGMENT order: a lane fetches its eight B fragments of a
//     64-key step with two 128-bit loads each (no bank conflicts, 4 instead of 16 load instructions per step);
//   * one exponential in four is evaluated on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial,
//     7.7e-5 relative error, below the 16-bit rounding of P) so the MUFU pipe and the issue slots run out together.
#include "pd_kernels.h"
#include <type_traits>

namespace pd {

constexpr int AT_WARPS = 8;        // 16 queries per warp, 128 per CTA
constexpr int AT_MAXS = 1024;      // keys resident in shared memory per pass

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial,
// scaled by 2^n through the exponent field
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -100.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = 0.05508868396282196f;
    p = fmaf(p, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <typename T>
__device__ __forceinline__ void mma_16x8x8(float c[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    if (std::is_same<T, bf16>::value)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(b0));
    else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(b0));
}
template <typename T>
__device__ __forceinline__ void mma_16x8x16(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                            uint32_t b1) {
    if (std::is_same<T, bf16>::value)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Shared-memory fragment layouts for one 64-key step `blk` (each 1 KB):
//   Kf[blk][g][t][kb]   word = K[key = 64 blk + 8 kb + g][dims 2t, 2t+1]            (B of m16n8k8, n-block kb)
//   Vf[blk][g][t][2j+h] word = V[keys 64 blk + 16 j + 8 h + 2t, +1][dim g]          (b0 (h=0) / b1 (h=1) of m16n8k16, k-block j)
// lane (g = lane >> 2, t = lane & 3) reads its 8 words of either with two 16-byte loads at word offset lane * 8.
template <typename T, int kPolyEvery>
__global__ void __launch_bounds__(AT_WARPS * 32) attention_mma_kernel(const T* __restrict__ qkv, int S, int C,
                                                                      T* __restrict__ out) {
    __shared__ __align__(16) uint32_t Kf[AT_MAXS * 4];
    __shared__ __align__(16) uint32_t Vf[AT_MAXS * 4];
    const int n = blockIdx.z, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const size_t rowp = (size_t)3 * C;
    const T* base = qkv + (size_t)n * S * rowp + head * 8;
    const int q0 = blockIdx.x * (AT_WARPS * 16) + warp * 16;
    const bool active = q0 < S;   // S % 16 == 0 is required, so a warp is fully in or fully out

    // Q fragment (A of m16n8k8): rows g, g+8; k = 2t, 2t+1
    uint32_t qa0 = 0, qa1 = 0;
    if (active) {
        qa0 = *reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g) * rowp + 2 * t);
        qa1 = *reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g + 8) * rowp + 2 * t);
    }
    const float sl = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)
    const uint32_t ones = std::is_same<T, bf16>::value ? 0x3F803F80u : 0x3C003C00u;
    float m0 = -INFINITY, m1 = -INFINITY;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    float l[4] = {0.f, 0.f, 0.f, 0.f};   // row sums of P through the tensor core: l[0] = row g, l[2] = row g+8

    for (int k0 = 0; k0 < S; k0 += AT_MAXS) {
        const int tk = min(AT_MAXS, S - k0);
        __syncthreads();
        for (int j = threadIdx.x; j < tk; j += blockDim.x) {
            const T* kp = base + (size_t)(k0 + j) * rowp + C;
            const uint4 kv = *reinterpret_cast<const uint4*>(kp);
            const uint4 vv = *reinterpret_cast<const uint4*>(kp + C);
            const int blk = j >> 6, r = j & 63;
            {   // K: key (kb = r >> 3, g = r & 7), word tt = dims 2tt, 2tt+1
                uint32_t* dst = Kf + blk * 256 + (r & 7) * 32 + (r >> 3);
                dst[0] = kv.x; dst[8] = kv.y; dst[16] = kv.z; dst[24] = kv.w;
            }
            {   // V: key r = 16 jj + 8 h + 2 tt + e -> half e of word [dim][tt][2 jj + h]
                const int jj = r >> 4, h = (r >> 3) & 1, tt = (r >> 1) & 3, e = r & 1;
                T* dst = reinterpret_cast<T*>(Vf + blk * 256 + tt * 8 + 2 * jj + h) + e;
                const T* ve = reinterpret_cast<const T*>(&vv);
#pragma unroll
                for (int d = 0; d < 8; ++d) dst[d * 64] = ve[d];   // 32 words per dim row
            }
        }
        __syncthreads();
        if (!active) continue;
        for (int c0 = 0; c0 < tk; c0 += 64) {
            const uint32_t* kf = Kf + (c0 >> 6) * 256 + lane * 8;
            const uint4 kA = *reinterpret_cast<const uint4*>(kf), kB = *reinterpret_cast<const uint4*>(kf + 4);
            const uint32_t kb_[8] = {kA.x, kA.y, kA.z, kA.w, kB.x, kB.y, kB.z, kB.w};
            float s[8][4];
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
                mma_16x8x8<T>(s[kb], qa0, qa1, kb_[kb]);
            }
            float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                cm0 = fmaxf(cm0, fmaxf(s[kb][0], s[kb][1]));
                cm1 = fmaxf(cm1, fmaxf(s[kb][2], s[kb][3]));
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);
            const float corr0 = ex2((m0 - nm0) * sl), corr1 = ex2((m1 - nm1) * sl);
            m0 = nm0; m1 = nm1;
            o[0] *= corr0; o[1] *= corr0; o[2] *= corr1; o[3] *= corr1;
            l[0] *= corr0; l[2] *= corr1;
            const float ms0 = nm0 * sl, ms1 = nm1 * sl;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                // one value in kPolyEvery goes to the FMA/ALU pipes (rotating over the 4 accumulator slots)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = s[kb][i] * sl - (i < 2 ? ms0 : ms1);
                    const bool poly = kPolyEvery > 0 && ((kb * 4 + i) % kPolyEvery) == (kPolyEvery - 1);
                    s[kb][i] = poly ? ex2_poly(x) : ex2(x);
                }
            }
            const uint32_t* vf = Vf + (c0 >> 6) * 256 + lane * 8;
            const uint4 vA = *reinterpret_cast<const uint4*>(vf), vB = *reinterpret_cast<const uint4*>(vf + 4);
            const uint32_t vb_[8] = {vA.x, vA.y, vA.z, vA.w, vB.x, vB.y, vB.z, vB.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a0 = pack2<T>(s[2 * j][0], s[2 * j][1]);
                const uint32_t a1 = pack2<T>(s[2 * j][2], s[2 * j][3]);
                const uint32_t a2 = pack2<T>(s[2 * j + 1][0], s[2 * j + 1][1]);
                const uint32_t a3 = pack2<T>(s[2 * j + 1][2], s[2 * j + 1][3]);
                mma_16x8x16<T>(o, a0, a1, a2, a3, vb_[2 * j], vb_[2 * j + 1]);
                mma_16x8x16<T>(l, a0, a1, a2, a3, ones, ones);   // denominator from the SAME rounded P the numerator uses
            }
        }
    }
    if (!active) return;
    const float i0 = 1.0f / l[0], i1 = 1.0f / l[2];
    T* ob = out + (size_t)n * S * C + head * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g) * C) = pack2<T>(o[0] * i0, o[1] * i0);
    *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g + 8) * C) = pack2<T>(o[2] * i1, o[3] * i1);
}

int launch_attention_mma(int dt, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s) {
    PD_REQUIRE(dt == DT_BF16 || dt == DT_F16, "attention_mma takes bf16 or fp16 activations");
    PD_REQUIRE(d == 8, "attention kernels implement attention_head_dim == 8 (the shipped configs)");
    PD_REQUIRE(S % 64 == 0 && C % 8 == 0, "attention_mma needs S % 64 == 0");
    static int poly_every = -1;
    if (poly_every < 0) {
        const char* e = getenv("PHENDIFF_B200_ATTN_POLY");   // 0: all exponentials on MUFU; 4 (default): one in four on the FMA pipes
        poly_every = e ? atoi(e) : 4;
        if (poly_every != 0 && poly_every != 2 && poly_every != 3 && poly_every != 4 && poly_every != 8) poly_every = 4;
    }
    dim3 grid((S + AT_WARPS * 16 - 1) / (AT_WARPS * 16), C / 8, N);
#define PD_ATT(PE) PD_DISPATCH_HALF(dt, T, (attention_mma_kernel<T, PE><<<grid, AT_WARPS * 32, 0, s>>>((const T*)qkv, S, C, (T*)out)))
    switch (poly_every) {
        case 0: PD_ATT(0); break;
        case 2: PD_ATT(2); break;
        case 3: PD_ATT(3); break;
        case 8: PD_ATT(8); break;
        default: PD_ATT(4); break;
    }
#undef PD_ATT
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pd
