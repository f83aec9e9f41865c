// phendiff_b200 — fused flash-style self-attention for head_dim 8 (diffusers Attention + AttnProcessor2_0 core:
// softmax(q k^T / sqrt(8)) v), bf16 in / fp32 softmax / bf16 out, on packed qkv (N, S, 3C).
//
// Shape note: d = 8 is below the tcgen05 bf16 K step (16) and P.V has N = 8; the tensor work is 4.5 % of the
// forward FLOPs while the exp / max / rescale work dominates.  The warp-level tensor-core
// shapes fit d = 8 exactly (QK^T: m16n8k8, P.V: m16n8k16 with the S accumulator fragment reused as the A operand),
// so this round's kernel is a register-resident flash kernel on those; K and V^T of one head live in shared memory.
#include "pd_kernels.h"
#include <type_traits>

namespace pd {

constexpr int AT_TK = 1024;        // keys per shared-memory tile
constexpr int AT_VPITCH = AT_TK + 8;  // padded V^T row pitch (elements): conflict-free B-fragment reads
constexpr int AT_WARPS = 8;        // 16 queries per warp

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attention_mma_kernel(const T* __restrict__ qkv, int S, int C,
                                                                      T* __restrict__ out) {
    __shared__ __align__(16) T Ks[AT_TK * 8];
    __shared__ __align__(16) T Vt[8 * AT_VPITCH];
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
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4] = {0.f, 0.f, 0.f, 0.f};

    for (int k0 = 0; k0 < S; k0 += AT_TK) {
        const int tk = min(AT_TK, S - k0);
        __syncthreads();
        for (int j = threadIdx.x; j < tk; j += blockDim.x) {
            const T* kp = base + (size_t)(k0 + j) * rowp + C;
            uint4 kv = *reinterpret_cast<const uint4*>(kp);
            uint4 vv = *reinterpret_cast<const uint4*>(kp + C);
            *reinterpret_cast<uint4*>(&Ks[j * 8]) = kv;
            const T* ve = reinterpret_cast<const T*>(&vv);
#pragma unroll
            for (int d = 0; d < 8; ++d) Vt[d * AT_VPITCH + j] = ve[d];
        }
        __syncthreads();
        if (!active) continue;
        for (int c0 = 0; c0 < tk; c0 += 64) {
            float s[8][4];
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[(c0 + kb * 8 + g) * 8 + 2 * t]);
                mma_16x8x8<T>(s[kb], qa0, qa1, b0);
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
            l0 *= corr0; l1 *= corr1;
            o[0] *= corr0; o[1] *= corr0; o[2] *= corr1; o[3] *= corr1;
            const float ms0 = nm0 * sl, ms1 = nm1 * sl;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
                s[kb][0] = ex2(s[kb][0] * sl - ms0); s[kb][1] = ex2(s[kb][1] * sl - ms0);
                s[kb][2] = ex2(s[kb][2] * sl - ms1); s[kb][3] = ex2(s[kb][3] * sl - ms1);
                l0 += s[kb][0] + s[kb][1];
                l1 += s[kb][2] + s[kb][3];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t a0 = pack2<T>(s[2 * j][0], s[2 * j][1]);
                const uint32_t a1 = pack2<T>(s[2 * j][2], s[2 * j][3]);
                const uint32_t a2 = pack2<T>(s[2 * j + 1][0], s[2 * j + 1][1]);
                const uint32_t a3 = pack2<T>(s[2 * j + 1][2], s[2 * j + 1][3]);
                const T* vp = &Vt[g * AT_VPITCH + c0 + 16 * j + 2 * t];
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vp);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vp + 8);
                mma_16x8x16<T>(o, a0, a1, a2, a3, b0, b1);
            }
        }
    }
    if (!active) return;
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    T* ob = out + (size_t)n * S * C + head * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g) * C) = pack2<T>(o[0] * i0, o[1] * i0);
    *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g + 8) * C) = pack2<T>(o[2] * i1, o[3] * i1);
}

int launch_attention_mma(int dt, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s) {
    PD_REQUIRE(dt == DT_BF16 || dt == DT_F16, "attention_mma takes bf16 or fp16 activations");
    PD_REQUIRE(d == 8, "attention kernels implement attention_head_dim == 8 (the shipped configs)");
    PD_REQUIRE(S % 64 == 0 && C % 8 == 0, "attention_mma needs S % 64 == 0");
    dim3 grid((S + AT_WARPS * 16 - 1) / (AT_WARPS * 16), C / 8, N);
    PD_DISPATCH_HALF(dt, T, (attention_mma_kernel<T><<<grid, AT_WARPS * 32, 0, s>>>((const T*)qkv, S, C, (T*)out)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pd
