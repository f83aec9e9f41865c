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
#include "pd_attn_common.cuh"

namespace pd {

constexpr int AT_WARPS = 8;        // 16 queries per warp, 128 per CTA
constexpr int AT_MAXS = 1024;      // keys resident in shared memory per pass

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
                                                                      float sl, T* __restrict__ out) {
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
    // sl = log2(e) / sqrt(8) divided by whatever factor the caller already folded into q
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

static int launch_attention_chunked(int dt, const void* qkv, int N, int S, int C, int d, float sl, void* out, cudaStream_t s) {
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
#define PD_ATT(PE) PD_DISPATCH_HALF(dt, T, (attention_mma_kernel<T, PE><<<grid, AT_WARPS * 32, 0, s>>>((const T*)qkv, S, C, sl, (T*)out)))
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


// =====================================================================================================================
// v3: one CTA per (image, head) — K / V^T of the head are staged into shared memory ONCE (fragment order, conflict-free
// 128-bit fragment loads) and every warp walks its 16-query blocks over them.  ncu on v2 (profiles/r1h_ncu_attention.md):
// MUFU 63 %, legacy HMMA pipe 38 % (a m16n8k8 costs as much pipe time as a m16n8k16), ALU 39 %, FMA 32 %, issue-bound at
// ~10.8 cycles per (lane, score).  What v3 removes from the per-score instruction stream:
//   * the scale FFMA: log2(e)/sqrt(d) is folded into the q rows of the fused qkv weight at finalize (pd_api.cu), so
//     scores leave the tensor core in log2 units;
//   * the max subtraction and the zero-initialisation of the S accumulators: -m rides in as the C operand of the QK^T MMA;
//   * the per-step running max, its shuffles and the rescale of (o, l): m is the exact row max of key block 0 and stays
//     fixed; fp32 accumulators absorb a stale (low) m, and P only overflows 16-bit storage if a later score exceeds m
//     by 2^16 — detected as a non-finite row sum, upon which the warp redoes that query block with the exact online
//     softmax (same code as block 0);
//   * half of the MUFU work: every other (row, key-octet) pair of scores is exponentiated on the FMA/ALU pipes, for
//     fp16 in PACKED half2 arithmetic (range reduction by the 1039 magic add, degree-3 minimax polynomial, exponent
//     field built from the magic sum: 12 instructions per 2 scores, 2.4e-4 rms relative error ~ the fp16 rounding of P).
// Expected balance per 64-key step and warp (issue ~146, MUFU 128, HMMA 128 cycles) vs ~346 measured for v2.
// =====================================================================================================================
constexpr int AH_WARPS = 8;
#ifndef AH_MIN_CTAS
#define AH_MIN_CTAS 4
#endif
constexpr float AH_SL = PD_ATTN_QFOLD;   // log2(e) / sqrt(8)

template <typename T>
__device__ __forceinline__ void mma_16x8x8_c(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, const float (&c)[4]) {
    if (std::is_same<T, bf16>::value)
        asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
            : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
            : "r"(a0), "r"(a1), "r"(b0), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
    else
        asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
            : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
            : "r"(a0), "r"(a1), "r"(b0), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

// One exact online-softmax step over a 64-key block (scores in log2 units): block 0 of every query block, and every block
// of a query block whose fast pass overflowed.
template <typename T>
__device__ __forceinline__ void attn_exact_step(const uint32_t* Kblk, const uint32_t* Vblk, int lane, uint32_t qa0, uint32_t qa1,
                                                uint32_t ones, float& m0, float& m1, float (&o)[4], float (&l)[4]) {
    const uint4* kf = reinterpret_cast<const uint4*>(Kblk) + lane;
    const uint4 kA = kf[0], kB = kf[32];
    const uint32_t kb_[8] = {kA.x, kA.y, kA.z, kA.w, kB.x, kB.y, kB.z, kB.w};
    float s[8][4];
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) mma_16x8x8_c<T>(s[kb], qa0, qa1, kb_[kb], zero4);
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
    const float corr0 = ex2(m0 - nm0), corr1 = ex2(m1 - nm1);
    m0 = nm0; m1 = nm1;
    o[0] *= corr0; o[1] *= corr0; o[2] *= corr1; o[3] *= corr1;
    l[0] *= corr0; l[1] *= corr0; l[2] *= corr1; l[3] *= corr1;
    uint32_t pa[8][2];
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        pa[kb][0] = pack2<T>(ex2(s[kb][0] - m0), ex2(s[kb][1] - m0));
        pa[kb][1] = pack2<T>(ex2(s[kb][2] - m1), ex2(s[kb][3] - m1));
    }
    const uint4* vf = reinterpret_cast<const uint4*>(Vblk) + lane;
    const uint4 vA = vf[0], vB = vf[32];
    const uint32_t vb_[8] = {vA.x, vA.y, vA.z, vA.w, vB.x, vB.y, vB.z, vB.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        mma_16x8x16<T>(o, pa[2 * j][0], pa[2 * j][1], pa[2 * j + 1][0], pa[2 * j + 1][1], vb_[2 * j], vb_[2 * j + 1]);
        mma_16x8x16<T>(l, pa[2 * j][0], pa[2 * j][1], pa[2 * j + 1][0], pa[2 * j + 1][1], ones, ones);
    }
}

// POLY_MASK: bit 2*kb (2*kb+1) set = the score pair of accumulator rows g (g+8) of key octet kb goes to the polynomial
template <typename T, uint32_t POLY_MASK>
__global__ void __launch_bounds__(AH_WARPS * 32, AH_MIN_CTAS) attention_head_kernel(const T* __restrict__ qkv, int S, int C, float qmul,
                                                                          T* __restrict__ out, const uint8_t* __restrict__ redo_flags) {
    extern __shared__ __align__(16) uint32_t ah_smem[];
    uint32_t* Kf = ah_smem;             // [S/64][2][32 lanes][4]: plane h, lane (g,t), word i = K[key 64 blk + 8 (4h+i) + g][dims 2t, 2t+1]
    uint32_t* Vf = ah_smem + S * 4;     // [S/64][2][32 lanes][4]: plane p, lane (g,t), word i = V[keys 64 blk + 16 (2p + (i>>1)) + 8 (i&1) + 2t, +1][dim g]
    const int n = blockIdx.z, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const size_t rowp = (size_t)3 * C;
    const T* base = qkv + (size_t)n * S * rowp + head * 8;
    // repair mode (after the tcgen05 kernel): only the 32-query groups it flagged (16-bit P overflow) are recomputed, exactly
    const uint8_t* redo = redo_flags ? redo_flags + ((size_t)n * gridDim.y + head) * (size_t)(S >> 7) * 4 : nullptr;
    if (redo) {
        bool any = false;
        for (int i = threadIdx.x; i < (S >> 7) * 4; i += blockDim.x) any = any || redo[i] != 0;
        if (!__syncthreads_or(any)) return;
    }

    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const T* kp = base + (size_t)j * rowp + C;
        const uint4 kv = *reinterpret_cast<const uint4*>(kp);
        const uint4 vv = *reinterpret_cast<const uint4*>(kp + C);
        const int blk = j >> 6, r = j & 63;
        {
            const int kb = r >> 3, gg = r & 7;
            uint32_t* dst = Kf + blk * 256 + (kb >> 2) * 128 + gg * 16 + (kb & 3);
            dst[0] = kv.x; dst[4] = kv.y; dst[8] = kv.z; dst[12] = kv.w;      // lane (gg, tt) at word (gg*4 + tt)*4
        }
        {
            const int jj = r >> 4, h = (r >> 3) & 1, tt = (r >> 1) & 3, e = r & 1;
            T* dst = reinterpret_cast<T*>(Vf + blk * 256 + (jj >> 1) * 128 + tt * 4 + (jj & 1) * 2 + h) + e;
            const T* ve = reinterpret_cast<const T*>(&vv);
#pragma unroll
            for (int d = 0; d < 8; ++d) dst[d * 32] = ve[d];                    // lane (d, tt): 16 words = 32 halves per dim
        }
    }
    __syncthreads();

    const uint32_t ones = std::is_same<T, bf16>::value ? 0x3F803F80u : 0x3C003C00u;
    constexpr bool kRecentre = std::is_same<T, f16>::value && (POLY_MASK & 0xFFFFu) != 0u && (POLY_MASK & 0xFFFFu) != 0xFFFFu;
    const int nblk = S >> 6, nqb = S >> 4;
    for (int qb = blockIdx.x * AH_WARPS + warp; qb < nqb; qb += gridDim.x * AH_WARPS) {
        const int q0 = qb * 16;
        uint32_t qa0 = *reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g) * rowp + 2 * t);
        uint32_t qa1 = *reinterpret_cast<const uint32_t*>(base + (size_t)(q0 + g + 8) * rowp + 2 * t);
        if (qmul != 1.0f) {   // q not pre-scaled by the caller (test entry, SIMT qkv projection)
            const T* h0 = reinterpret_cast<const T*>(&qa0);
            const T* h1 = reinterpret_cast<const T*>(&qa1);
            qa0 = pack2<T>(to_f(h0[0]) * qmul, to_f(h0[1]) * qmul);
            qa1 = pack2<T>(to_f(h1[0]) * qmul, to_f(h1[1]) * qmul);
        }
        float o[4] = {0.f, 0.f, 0.f, 0.f}, l[4] = {0.f, 0.f, 0.f, 0.f};
        float m0 = -INFINITY, m1 = -INFINITY;
        if (redo && !redo[(qb >> 3) * 4 + ((qb & 7) >> 1)]) continue;
        bool need_exact = redo != nullptr;
        if (!redo) {
        // key block 0: exact step, fixes the row maxima for the fast steps
        attn_exact_step<T>(Kf, Vf, lane, qa0, qa1, ones, m0, m1, o, l);
        {
            // C operand of the QK^T MMAs: scores arrive as s - m.  The quad is produced BY an MMA (0 * 0 + c) so that ptxas keeps
            // it in four consecutive registers for the whole loop instead of re-assembling it with moves before every HMMA.
            const float cinit[4] = {-m0, -m0, -m1, -m1};
            float cq[4];
            mma_16x8x8_c<T>(cq, 0u, 0u, 0u, cinit);
            // 32 keys per half step (one 128-bit K fragment load, one V fragment load): short live ranges keep the kernel at
            // 64 registers = 8 warps per scheduler, which is what lets the MUFU / HMMA / FMA pipes overlap
            // (tools/microbench/pipes.cu: the same instruction mix runs 1.34x faster at 8 than at 6 warps per scheduler)
            const uint4* kf = reinterpret_cast<const uint4*>(Kf + 256) + lane;
            const uint4* vf = reinterpret_cast<const uint4*>(Vf + 256) + lane;
            // largest P seen on the MUFU path since the last re-centring (fp16 only): a stale row max costs the packed-half
            // polynomial input precision (x is rounded to fp16 before the range reduction), so when the sampled P exceeds
            // 2^2 the row is re-centred on it — o and l are scaled by 1/P, m grows by log2(P).  Sampling the MUFU half of the
            // scores is enough for precision; overflow proper is still caught by the non-finite check below.
            float pm0 = 0.f, pm1 = 0.f;
#pragma unroll 2
            for (int hb = 2; hb < 2 * nblk; ++hb, kf += 32, vf += 32) {
                const uint4 kA = *kf;
                const uint32_t kb_[4] = {kA.x, kA.y, kA.z, kA.w};
                uint32_t pa[4][2];
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    float s[4];
                    mma_16x8x8_c<T>(s, qa0, qa1, kb_[kb], cq);
                    if ((POLY_MASK >> (2 * kb)) & 1u) pa[kb][0] = ex2_pair_poly<T>(s[0], s[1]);
                    else {
                        const float p0 = ex2(s[0]), p1 = ex2(s[1]);
                        if (kRecentre) pm0 = fmaxf(pm0, fmaxf(p0, p1));
                        pa[kb][0] = pack2<T>(p0, p1);
                    }
                    if ((POLY_MASK >> (2 * kb + 1)) & 1u) pa[kb][1] = ex2_pair_poly<T>(s[2], s[3]);
                    else {
                        const float p2 = ex2(s[2]), p3 = ex2(s[3]);
                        if (kRecentre) pm1 = fmaxf(pm1, fmaxf(p2, p3));
                        pa[kb][1] = pack2<T>(p2, p3);
                    }
                }
                const uint4 vA = *vf;
                const uint32_t vb_[4] = {vA.x, vA.y, vA.z, vA.w};
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_16x8x16<T>(o, pa[2 * j][0], pa[2 * j][1], pa[2 * j + 1][0], pa[2 * j + 1][1], vb_[2 * j], vb_[2 * j + 1]);
                    mma_16x8x16<T>(l, pa[2 * j][0], pa[2 * j][1], pa[2 * j + 1][0], pa[2 * j + 1][1], ones, ones);
                }
                if (kRecentre && (hb & 1) && __any_sync(0xffffffffu, fmaxf(pm0, pm1) > 4.0f)) {
                    pm0 = fmaxf(pm0, __shfl_xor_sync(0xffffffffu, pm0, 1));
                    pm0 = fmaxf(pm0, __shfl_xor_sync(0xffffffffu, pm0, 2));
                    pm1 = fmaxf(pm1, __shfl_xor_sync(0xffffffffu, pm1, 1));
                    pm1 = fmaxf(pm1, __shfl_xor_sync(0xffffffffu, pm1, 2));
                    // shift by the power of two below the sampled maximum (exact scaling; never re-centre downwards)
                    const int d0 = (__float_as_int(fminf(fmaxf(pm0, 1.0f), 1.0e30f)) >> 23) - 127;
                    const int d1 = (__float_as_int(fminf(fmaxf(pm1, 1.0f), 1.0e30f)) >> 23) - 127;
                    const float r0 = __int_as_float((127 - d0) << 23), r1 = __int_as_float((127 - d1) << 23);
                    o[0] *= r0; o[1] *= r0; o[2] *= r1; o[3] *= r1;
                    l[0] *= r0; l[1] *= r0; l[2] *= r1; l[3] *= r1;
                    m0 += (float)d0; m1 += (float)d1;
                    const float cnew[4] = {-m0, -m0, -m1, -m1};
                    mma_16x8x8_c<T>(cq, 0u, 0u, 0u, cnew);
                    pm0 = pm1 = 0.f;
                }
            }
        }
        const bool bad = !(fabsf(l[0]) <= 3.0e38f) || !(fabsf(l[2]) <= 3.0e38f) || !(fabsf(o[0]) <= 3.0e38f) ||
                         !(fabsf(o[1]) <= 3.0e38f) || !(fabsf(o[2]) <= 3.0e38f) || !(fabsf(o[3]) <= 3.0e38f);
        need_exact = __any_sync(0xffffffffu, bad);
        }
        if (need_exact) {
            // a score beyond the block-0 maximum by 2^16 overflowed the 16-bit P (or the input holds inf / NaN): redo this
            // query block with the exact online softmax
            m0 = m1 = -INFINITY;
            o[0] = o[1] = o[2] = o[3] = 0.f;
            l[0] = l[1] = l[2] = l[3] = 0.f;
#pragma unroll 1
            for (int blk = 0; blk < nblk; ++blk) attn_exact_step<T>(Kf + blk * 256, Vf + blk * 256, lane, qa0, qa1, ones, m0, m1, o, l);
        }
        const float i0 = 1.0f / l[0], i1 = 1.0f / l[2];
        T* ob = out + (size_t)n * S * C + head * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g) * C) = pack2<T>(o[0] * i0, o[1] * i0);
        *reinterpret_cast<uint32_t*>(ob + (size_t)(q0 + g + 8) * C) = pack2<T>(o[2] * i1, o[3] * i1);
    }
}

template <typename T, uint32_t MASK>
static int launch_head(const void* qkv, int N, int S, int C, float qmul, void* out, dim3 grid, size_t smem, cudaStream_t s,
                       const uint8_t* redo = nullptr) {
    static size_t attr_dev[PD_MAX_DEVICES] = {0};
    size_t& attr = attr_dev[pd_cur_dev()];
    if (smem > 48 * 1024 && smem > attr) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(attention_head_kernel<T, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    attention_head_kernel<T, MASK><<<grid, AH_WARPS * 32, smem, s>>>((const T*)qkv, S, C, qmul, (T*)out, redo);
    return 0;
}

// kernel launches one launch_attention_mma call makes (the tcgen05 variant is followed by its repair pass)
int attention_mma_launches(int S) {
    const char* e = getenv("PHENDIFF_B200_ATTN_KERNEL");
    const bool tc = e && e[0] == 't' && e[1] == 'c';
    if (!e || (tc && e[2] == '3')) return (S % 128 == 0 && (size_t)S * 32 <= 200 * 1024 && attention_tc3_supported(S, 8)) ? 2 : 1;
    return (tc && S % 128 == 0 && (size_t)S * 32 <= 200 * 1024 && attention_tc_smem_bytes(S) <= 110 * 1024) ? 2 : 1;
}

// qfold: the factor the caller already folded into q (1 = raw q; AH_SL = the finalize-time fold of pd_api.cu)
int launch_attention_mma(int dt, const void* qkv, int N, int S, int C, int d, float qfold, void* out, cudaStream_t s, int force_variant) {
    PD_REQUIRE(dt == DT_BF16 || dt == DT_F16, "attention_mma takes bf16 or fp16 activations");
    PD_REQUIRE(d == 8, "attention kernels implement attention_head_dim == 8 (the shipped configs)");
    PD_REQUIRE(S % 64 == 0 && C % 8 == 0, "attention_mma needs S % 64 == 0");
    static int variant = -1, polyv = -1;
    if (variant < 0) {
        const char* e = getenv("PHENDIFF_B200_ATTN_KERNEL");
        // default "tc3" (pd_attn_tc3.cu: persistent tcgen05 / TMEM / TMA kernel, three independent softmax streams; shapes it does
        // not take — S < 384, S not a multiple of 128, raw q — fall through to "v3"); "v3": warp-level head-resident mma.sync kernel;
        // "tc" / "tc2": the first two tcgen05 kernels; "v2": chunked warp-level kernel (any S)
        variant = !e ? 6 : ((e[0] == 'v' && e[1] == '2') ? 2 : ((e[0] == 't' && e[1] == 'c') ? (e[2] == '3' ? 6 : (e[2] == '2' ? 5 : 4)) : 3));
        const char* pe = getenv("PHENDIFF_B200_ATTN_POLYPAIRS");   // score pairs per 16 (v3) / per 8 (tc) on the FMA/ALU pipes
        polyv = pe ? atoi(pe) : -1;
    }
    const size_t smem = (size_t)S * 32;
    const int var = force_variant ? force_variant : variant;
    if (var == 2 || smem > 200 * 1024) return launch_attention_chunked(dt, qkv, N, S, C, d, AH_SL / qfold, out, s);
    // tc3 stages q by TMA exactly as it lies in memory: it needs the finalize-time fold (qfold == AH_SL); raw q takes the v3 kernel
    const bool tc3 = var == 6 && S % 128 == 0 && attention_tc3_supported(S, C) && fabsf(AH_SL / qfold - 1.0f) < 1e-6f;
    if (tc3 || ((var == 4 || var == 5) && S % 128 == 0 && attention_tc_smem_bytes(S) <= 110 * 1024)) {
        // flags: one byte per (image, head, 128-query tile, warp); grow-only scratch owned by the library
        static uint8_t* flags_dev[PD_MAX_DEVICES] = {nullptr};
        static size_t flags_cap_dev[PD_MAX_DEVICES] = {0};
        uint8_t*& flags = flags_dev[pd_cur_dev()];
        size_t& flags_cap = flags_cap_dev[pd_cur_dev()];
        const size_t need = (size_t)N * (C / 8) * (S >> 7) * 4;
        if (need > flags_cap) {
            if (flags) PD_CHECK_CUDA(cudaFree(flags));
            PD_CHECK_CUDA(cudaMalloc(&flags, need));
            flags_cap = need;
        }
        const int tc_pp = polyv < 0 ? (dt == DT_F16 ? 4 : 2) : polyv;   // fp16: half of the exponent pairs on the packed-half polynomial
        int rc = tc3 ? launch_attention_tc3(dt, qkv, N, S, C, out, flags, tc_pp, s)
                 : (var == 5 ? launch_attention_tc2(dt, qkv, N, S, C, AH_SL / qfold, out, flags, tc_pp, s)
                             : launch_attention_tc(dt, qkv, N, S, C, AH_SL / qfold, out, flags, tc_pp, s));
        if (rc) return rc;
        dim3 grid(1, C / 8, N);
        PD_DISPATCH_HALF(dt, T, (launch_head<T, 0x0000u>(qkv, N, S, C, AH_SL / qfold, out, grid, smem, s, flags)));
        PD_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const float qmul = AH_SL / qfold;
    const int heads = C / 8;
    int split = 1;
    while (split < 8 && (size_t)heads * N * split < 1184 && (S >> 4) / (split * 2) >= AH_WARPS) split *= 2;
    dim3 grid(split, heads, N);
    int pp = polyv;
    if (pp < 0) pp = (dt == DT_F16) ? 6 : 4;   // measured flat between 4 and 8 (profiles/r1k_attention.md); the sum of pipe cycles is what counts
#define PD_AH(MASK) PD_DISPATCH_HALF(dt, T, (launch_head<T, MASK>(qkv, N, S, C, qmul, out, grid, smem, s)))
    switch (pp) {
        case 0: PD_AH(0x0000u); break;
        case 4: PD_AH(0x4242u); break;
        case 6: PD_AH(0x6262u); break;
        case 7: PD_AH(0x6662u); break;
        case 10: PD_AH(0xE6E6u); break;
        case 12: PD_AH(0xEEEEu); break;
        default: PD_AH(0x6666u); break;
    }
#undef PD_AH
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pd
