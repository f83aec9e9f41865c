// phendiff_b200 — pieces shared by the two tcgen05 convolution kernels (pd_conv_tc.cu, pd_conv_halo.cu):
// UMMA shared-memory / instruction descriptors, the fused epilogue (bias + time-embedding row + residual + scale ->
// 16-bit NHWC store + GroupNorm chunk statistics of the stored values), and the conv_out epilogue that applies the
// DDIM / inverse-DDIM update to x_t in place.
#pragma once
#include "pd_kernels.h"
#include <type_traits>

namespace pd {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;
constexpr int TC_THREADS = 256;        // per-tap kernel: 4 control warps + 4 epilogue warps
constexpr int HALO_THREADS = 384;      // halo kernel: 4 control warps + 8 epilogue warps
constexpr int HALO_THREADS_GN = 512;   // + 4 warps that apply GroupNorm + SiLU to the raw halo tiles in shared memory

struct TcEpi {
    const float* bias;          // (Cout) or null
    const float* addvec;        // (rows, addvec_stride) or null
    const int32_t* addvec_row;  // (N) or null
    int addvec_stride;
    const void* residual;       // NHWC, same geometry as out, or null
    float out_scale;
    void* out;                  // NHWC 16-bit
    double* stats;              // (N, Cout/stats_cw, 2) fp64 or null
    int stats_cw;               // channels per statistics chunk: 4 or 2
    int Cout;
    // TC_MODE_DDIM
    float* model_out;           // NCHW fp32 or null
    float* x_t;                 // NCHW fp32, updated in place, or null
    pd_step_coeffs_t step;
    int c_valid;                // real output channels (<= 16)
    int plane;                  // Ho*Wo
    int img_off;                // first image of this launch: model_out / x_t / cfg are indexed by (img - img_off)
    CfgEpi cfg;                 // classifier-free guidance combine ahead of the scheduler update, or uncond == null
};

// K-major SWIZZLE_128B operand descriptor (PTX "matrix descriptor"): start address >> 4, LBO unused for swizzled
// K-major, SBO = byte stride between 8-row groups, version 1 (sm_100), base offset = phase of the first row inside the
// 1024-byte swizzle pattern when the start address is not 1024-aligned, layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t sbo_bytes = 1024, uint32_t base_off = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// The same descriptor as (high word, low word): the high word (SBO, version, layout; base offset 0) is constant per operand,
// the low word is the 16-byte-granular shared-memory address, so stepping through taps / K slices / ring stages is one 32-bit add.
__host__ __device__ constexpr uint32_t sw128_desc_hi(uint32_t sbo_bytes) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// instruction descriptor, kind::f16: D fp32 (bit 4), A/B format at bits 7/10 (0 = fp16, 1 = bf16), both K-major,
// N >> 3 at bits 17.., M >> 4 at bits 24..
template <typename T, int BLOCK_N> __device__ __forceinline__ constexpr uint32_t make_idesc() {
    constexpr uint32_t fmt = std::is_same<T, bf16>::value ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// Sum NV per-lane values across the 32 lanes of a warp with a transpose-reduce butterfly (NV-1 shuffles + 1 instead of
// 5*NV): on return v[0] of lane L holds the full sum of value index (L >> (5 - log2(NV))) ... see callers.
template <int NV> __device__ __forceinline__ float warp_transpose_reduce(float (&v)[NV], int lane) {
    static_assert(NV == 16 || NV == 32, "NV must be 16 or 32");
    int m = 16;
#pragma unroll
    for (int k = NV / 2; k >= 1; k >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < k; ++i) {
            const float keep = up ? v[i + k] : v[i];
            const float send = up ? v[i] : v[i + k];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
        m >>= 1;
    }
    if (NV == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// Epilogue of one accumulator row (= one output pixel) x 32 consecutive output channels starting at col0.
// All 32 lanes of the calling warp must belong to the same image `img` when e.stats != null.
// STAGED = false: 16-bit values go straight to global memory (each lane writes its own pixel row: strided 16-byte stores).
// STAGED = true:  they go to the lane's 128-byte row of a SWIZZLE_128B shared-memory slab (`srow`, 16-byte chunks
//                 chunk_base..chunk_base+3) that one TMA store then writes out fully coalesced; the residual arrives the same
//                 way (`res4`: this lane's 32 residual channels, already transposed through shared memory).
template <typename T, bool STAGED, int CW>
__device__ __forceinline__ void tc_epilogue_chunk32(const TcEpi& e, const uint32_t (&r)[32], int col0, size_t pix, int img,
                                                    int lane, uint8_t* srow = nullptr, int chunk_base = 0,
                                                    const uint4* res4 = nullptr, bool use_res = false) {
    T* orow = STAGED ? nullptr : reinterpret_cast<T*>(e.out) + pix * e.Cout + col0;
    const T* rrow = (!STAGED && e.residual) ? reinterpret_cast<const T*>(e.residual) + pix * e.Cout + col0 : nullptr;
    const float* av = nullptr;
    if (e.addvec) av = e.addvec + (size_t)(e.addvec_row ? e.addvec_row[img] : img) * e.addvec_stride + col0;
    const float* bs = e.bias ? e.bias + col0 : nullptr;
    // CW == 4: sv[0..7] sums of the eight 4-channel chunks, sv[8..15] sums of squares; CW == 2: [0..15] sums, [16..31] squares
    float sv[64 / CW];
    const bool st4 = CW == 4 && e.stats != nullptr, st2 = CW == 2 && e.stats != nullptr;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g8 * 8 + i]);
        if (bs) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + g8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bs + g8 * 8 + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (av) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(av + g8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(av + g8 * 8 + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (STAGED) {
            if (use_res) {
                float rv[8];
                unpack8<T>(res4[g8], rv);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += rv[i];
            }
        } else if (rrow) {
            float rv[8];
            load8(rrow + g8 * 8, rv);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += rv[i];
        }
        if (e.out_scale != 1.0f) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= e.out_scale;
        }
        if (STAGED) store8(reinterpret_cast<T*>(srow + (((chunk_base + g8) ^ (lane & 7)) << 4)), v);
        else store8(orow + g8 * 8, v);
        // GroupNorm statistics from the fp32 values (before the 16-bit rounding of the store: the difference averages out
        // over the thousands of elements of a group and keeps 8 conversions per vector off the epilogue's critical path)
        if (CW == 4) {
            sv[g8 * 2 + 0] = (v[0] + v[1]) + (v[2] + v[3]);
            sv[g8 * 2 + 1] = (v[4] + v[5]) + (v[6] + v[7]);
            sv[8 + g8 * 2 + 0] = (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
            sv[8 + g8 * 2 + 1] = (v[4] * v[4] + v[5] * v[5]) + (v[6] * v[6] + v[7] * v[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sv[g8 * 4 + j] = v[2 * j] + v[2 * j + 1];
                sv[16 + g8 * 4 + j] = v[2 * j] * v[2 * j] + v[2 * j + 1] * v[2 * j + 1];
            }
        }
    }
    if (st4) {
        const float tot = warp_transpose_reduce<64 / CW>(sv, lane);
        if ((lane & 1) == 0) {
            const int idx = lane >> 1;   // 0..7: sum of chunk idx, 8..15: sum of squares of chunk idx-8
            atomicAdd(e.stats + ((size_t)img * (e.Cout >> 2) + (col0 >> 2) + (idx & 7)) * 2 + (idx >> 3), (double)tot);
        }
    } else if (st2) {
        const float tot = warp_transpose_reduce<64 / CW>(sv, lane);
        atomicAdd(e.stats + ((size_t)img * (e.Cout >> 1) + (col0 >> 1) + (lane & 15)) * 2 + (lane >> 4), (double)tot);
    }
}

// conv_out epilogue: one pixel x (up to 16) output channels -> NCHW fp32 model output and / or in-place x_t update
__device__ __forceinline__ void tc_epilogue_ddim(const TcEpi& e, const uint32_t (&r)[16], int img, int hw) {
    // all x_t loads are issued before the first store: the compiler cannot prove that a store to x_t[.., c, ..] does not feed
    // the load of channel c + 1, and would otherwise serialise three global round trips per pixel (ncu r1r: conv_out spent
    // 9 k cycles per tile at 6 % tensor activity, long-scoreboard bound)
    float xv[16], uv[16];
    const size_t base = (size_t)(img - e.img_off) * e.c_valid * e.plane + hw;
    if (e.x_t) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < e.c_valid) xv[c] = e.x_t[base + (size_t)c * e.plane];
    }
    const bool guided = e.cfg.uncond != nullptr;
    float gw = 0.f;
    if (guided) {
        gw = __ldg(e.cfg.w + (img - e.img_off));
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < e.c_valid) uv[c] = e.cfg.uncond[base + (size_t)c * e.plane];
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        if (c < e.c_valid) {
            float m = __uint_as_float(r[c]) + (e.bias ? __ldg(e.bias + c) : 0.f);
            if (guided) m = (e.cfg.eqn == 0 ? uv[c] : m) + gw * (m - uv[c]);
            const size_t idx = base + (size_t)c * e.plane;
            if (e.model_out) e.model_out[idx] = m;
            if (e.x_t) e.x_t[idx] = ddim_update(e.step, xv[c], m, 0.f, nullptr);
        }
    }
}

// ---- host-side helpers shared by both kernels ---------------------------------------------------------------------
int tc_encode_map(CUtensorMap* tm, int dt, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128 = true);
int tc_num_sms();

enum { TC_KIND_TAP = 0, TC_KIND_HALO = 1 };
struct ConvTapPlan;
struct ConvHaloPlan;
struct ConvTcPlan {
    int kind;
    ConvTapPlan* tap = nullptr;
    ConvHaloPlan* halo = nullptr;
};
int conv_tap_plan_create(const ConvTcDesc& d, ConvTapPlan** out);
void conv_tap_plan_destroy(ConvTapPlan* p);
int conv_tap_launch(const ConvTapPlan* p, cudaStream_t s);
int conv_halo_plan_create(const ConvTcDesc& d, ConvHaloPlan** out);
void conv_halo_plan_destroy(ConvHaloPlan* p);
int conv_halo_launch(const ConvHaloPlan* p, cudaStream_t s, const ConvTcLaunch* extra);

}  // namespace pd
