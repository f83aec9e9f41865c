// phendiff_b200 — flash-style self-attention for head_dim 8 on tcgen05 / TMEM (diffusers Attention + AttnProcessor2_0 core,
// SURVEY A.2): softmax(q k^T / sqrt(8)) v on packed qkv (N, S, 3C), 16-bit in / fp32 accumulate / 16-bit out.
//
// Why not mma.sync: the warp-level kernel (pd_attn_mma.cu) is bound by the SUM of its pipe occupancies — every warp
// interleaves HMMA (8 cycles of the legacy tensor pipe per instruction), MUFU.EX2 (8 cycles) and packed-half FMA work in one
// in-order stream, and ncu shows the warps queueing on whichever pipe is busy (profiles/r1k_attention.md: 41 % of the hot
// loop's stall samples sit on HMMA.16816, no pipe above 55 %).  Here the matrix products leave the warps altogether: one
// thread issues tcgen05.mma, the 128 softmax threads only see TMEM loads / stores and the exponentials.
//
// One CTA per (image, head); two CTAs per SM (256 TMEM columns each), so one CTA's staging overlaps the other's math.
//   shared memory (S tokens, SWIZZLE_128B K-major tiles, written by the CTA's own threads):
//     Q', K' : [S/512][128 rows][128 B]: the 32-byte K-step js of row r holds token (4 t + js) * 128 + r as 16 halves
//              {x0..x7, a0, a1, 0 x 6}: K' has a = (1, 1); Q' has a = (-m_hi, -m_lo) once the row max m of the first key
//              tile is known, so that S = q.k - m comes out of the tensor core with NO per-score instruction
//              (hi + lo 16-bit split: |error| <= 2^-22 |m|);
//     V'^T   : [S/64][16 rows][64 keys]: rows 0..7 = v dims, row 8 = ones (the PV product's column 8 is the softmax
//              denominator, from the same rounded P the numerator uses), rows 9..15 = 0.
//   TMEM: three S/P buffers of 64 fp32 columns (S = 128 queries x 64 keys; P overwrites columns 0..31 as packed 16-bit
//         pairs [two halves of 16 columns, at columns 0 and 32] = the A operand of the PV product, read straight from TMEM) + two O slots of 16 columns (O accumulates
//         over all key tiles of a query tile, the softmax threads read it once per query tile).
//   warps 0-3: softmax (thread = one query row: row max / exponent pairs are thread-local, no shuffles);
//   warp 4: MMA issuer; warp 5: TMEM allocator; warps 4-7 also stage.
// Softmax numerics follow pd_attn_mma.cu v3: scores arrive in log2 units (q pre-scaled at finalize), the row max is fixed
// after key tile 0, half of the exponent pairs run on the FMA / ALU pipes in packed fp16 (fp16 storage) or fp32 (bf16
// storage); a query tile whose 16-bit P overflowed (non-finite denominator) is flagged and recomputed by the exact
// warp-level kernel (launch_attention_fix).
#include "pd_attn_common.cuh"

namespace pd {

constexpr int ATC_SM_WARPS = 8;                       // softmax warps: two warpgroups, alternating key tiles
constexpr int ATC_THREADS = (ATC_SM_WARPS + 2) * 32;  // + MMA issuer + TMEM allocator
constexpr int ATC_TMEM_COLS = 256;
#ifndef ATC_WAIT_PV
#define ATC_WAIT_PV 0
#endif
constexpr int ATC_S_COLS = 64;
constexpr int ATC_O_COL0 = 192;

__device__ __forceinline__ void umma_f16kind_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32b_x16b(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// mbarrier wait of the softmax / issuer loops: same bounded-wait contract as mbar_wait (a protocol bug traps instead of hanging
// the box), but the 64-bit clock is read once per 64 polls — in the r3b capture the watchdog arithmetic of the shared helper was
// 8 % of this kernel's issued instructions, competing with the exponentials for issue slots.
__device__ __forceinline__ void atc_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 64; ++k)
            if (mbar_try_wait(bar, parity)) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("phendiff_b200: attention_tc mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}

// K-major SWIZZLE_128B descriptor, 8-row groups 1024 B apart (same encoding as the convolution kernels)
__device__ __forceinline__ uint64_t atc_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <typename T, int N> __device__ __forceinline__ constexpr uint32_t atc_idesc() {
    constexpr uint32_t fmt = std::is_same<T, bf16>::value ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <typename T> __device__ __forceinline__ uint32_t atc_pack_raw(T lo, T hi) {
    return (uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
}

// byte offset of the 16-byte chunk `chunk` (0..7) of row r (0..127) inside one [128 rows][128 B] swizzled tile
__device__ __forceinline__ uint32_t atc_tile_off(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

// exponentials of 32 scores (one thread's row x 32 keys) -> 16 packed pairs; PP of every 8 pairs go to the polynomial
template <typename T, int PP>
__device__ __forceinline__ void atc_exp32(const uint32_t (&s)[32], float sub, uint32_t* p) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float x0 = __uint_as_float(s[2 * r]) - sub, x1 = __uint_as_float(s[2 * r + 1]) - sub;
        const bool poly = ((r * PP) & 7) < PP && PP > 0;
        p[r] = poly ? ex2_pair_poly<T>(x0, x1) : pack2<T>(ex2(x0), ex2(x1));
    }
}
template <typename T, int PP>
__device__ __forceinline__ void atc_exp32_nosub(const uint32_t (&s)[32], uint32_t* p) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float x0 = __uint_as_float(s[2 * r]), x1 = __uint_as_float(s[2 * r + 1]);
        const bool poly = ((r * PP) & 7) < PP && PP > 0;
        p[r] = poly ? ex2_pair_poly<T>(x0, x1) : pack2<T>(ex2(x0), ex2(x1));
    }
}

template <typename T, int PP>
__global__ void __launch_bounds__(ATC_THREADS, 2) attention_tc_kernel(const T* __restrict__ qkv, int S, int C, float qmul,
                                                                      T* __restrict__ out, uint8_t* __restrict__ flags) {
    extern __shared__ uint8_t atc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int ntile = (S + 511) >> 9;                 // [128 x 128 B] tiles of Q' / K'
    uint8_t* smQ = smem;
    uint8_t* smK = smQ + (size_t)ntile * 16384;
    uint8_t* smV = smK + (size_t)ntile * 16384;       // S / 64 blocks of 2 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smV + (size_t)S * 32);
    uint64_t* sfull = bars;          // [3] MMA -> softmax: S tile landed
    uint64_t* pready = bars + 3;     // [3] softmax -> MMA: P written
    uint64_t* pvdone = bars + 6;     // [3] MMA -> MMA: PV product has consumed P (the buffer may take the next S)
    uint64_t* ofull = bars + 9;      // [2] MMA -> softmax: O of a query tile complete
    uint64_t* oread = bars + 11;     // [2] softmax -> MMA: O slot read
    uint64_t* qmready = bars + 13;   // [1] softmax -> MMA: -m written into Q'
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int n = blockIdx.y, head = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t rowp = (size_t)3 * C;
    const T* base = qkv + (size_t)n * S * rowp + head * 8;
    const int nqt = S >> 7, ntl = S >> 6, total = nqt * ntl;

    if (warp == ATC_SM_WARPS && lane == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(&sfull[i], 1); mbar_init(&pready[i], 256); mbar_init(&pvdone[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&ofull[i], 1); mbar_init(&oread[i], 128); }
        mbar_init(qmready, 128);
        fence_barrier_init();
    }
    if (warp == ATC_SM_WARPS + 1) {
        tmem_alloc(tmem_slot, ATC_TMEM_COLS);
        tmem_relinquish();
    }
    // ---- stage Q', K', V'^T (all threads) ----
    {
        const T one = from_f<T>(1.0f), zero = from_f<T>(0.0f);
        const uint32_t ones2 = atc_pack_raw<T>(one, one);
        for (int tok = threadIdx.x; tok < S; tok += ATC_THREADS) {
            const T* tp = base + (size_t)tok * rowp;
            uint4 qv = *reinterpret_cast<const uint4*>(tp);
            const uint4 kv = *reinterpret_cast<const uint4*>(tp + C);
            const uint4 vv = *reinterpret_cast<const uint4*>(tp + 2 * C);
            if (qmul != 1.0f) {
                float f[8];
                unpack8<T>(qv, f);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] *= qmul;
                qv.x = pack2<T>(f[0], f[1]); qv.y = pack2<T>(f[2], f[3]); qv.z = pack2<T>(f[4], f[5]); qv.w = pack2<T>(f[6], f[7]);
            }
            const int blk = tok >> 7, r = tok & 127, t = blk >> 2, js = blk & 3;
            uint8_t* qrow = smQ + (size_t)t * 16384;
            uint8_t* krow = smK + (size_t)t * 16384;
            *reinterpret_cast<uint4*>(qrow + atc_tile_off(r, 2 * js)) = qv;
            *reinterpret_cast<uint4*>(qrow + atc_tile_off(r, 2 * js + 1)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(krow + atc_tile_off(r, 2 * js)) = kv;
            *reinterpret_cast<uint4*>(krow + atc_tile_off(r, 2 * js + 1)) = make_uint4(ones2, 0u, 0u, 0u);
            const int vb = tok >> 6, kk = tok & 63;
            uint8_t* vblk = smV + (size_t)vb * 2048;
            const T* ve = reinterpret_cast<const T*>(&vv);
#pragma unroll
            for (int d = 0; d < 8; ++d)
                *reinterpret_cast<T*>(vblk + d * 128 + ((((kk >> 3) ^ d)) << 4) + (kk & 7) * 2) = ve[d];
        }
        // rows 8..15 of every V'^T block: ones, then zeros (uniform rows: the swizzle does not matter)
        const uint4 o4 = make_uint4(ones2, ones2, ones2, ones2), z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < ntl * 64; i += ATC_THREADS) {
            const int vb = i >> 6, c = i & 63;                       // 64 16-byte chunks = rows 8..15
            *reinterpret_cast<uint4*>(smV + (size_t)vb * 2048 + 1024 + c * 16) = (c < 8) ? o4 : z4;
        }
        (void)zero;
    }
    fence_proxy_async();          // generic-proxy writes above are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == ATC_SM_WARPS) {
        if (elect_one()) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idS = atc_idesc<T, 64>(), idPV = atc_idesc<T, 16>();
            const uint32_t q0 = smem_u32(smQ), k0 = smem_u32(smK), v0 = smem_u32(smV);
            // tile counters are advanced incrementally: ncu's source view (profiles/r3b_attention_tc.md) showed the
            // software division by the runtime ntl (I2F / MUFU.RCP / F2I chains) of `i / ntl`, `i % 3` on every tile
            int s_qt = 0, s_j = 0, s_buf = 0;       // the S tile issue_S() issues next
            auto issue_S = [&]() {
                const int qt = s_qt, j = s_j, buf = s_buf;
                if (++s_j == ntl) { s_j = 0; ++s_qt; }
                if (++s_buf == 3) s_buf = 0;
                if (j == 1) { atc_wait(qmready, qt & 1); tc_fence_after(); }
                const int kb = j >> 1;                                                  // 128-token block of the key tile
                const uint64_t a = atc_desc(q0 + (uint32_t)((qt >> 2) * 16384 + (qt & 3) * 32));
                const uint64_t b = atc_desc(k0 + (uint32_t)((kb >> 2) * 16384 + (j & 1) * 8192 + (kb & 3) * 32));
                umma_f16kind(tmem_base + (uint32_t)(buf * ATC_S_COLS), a, b, idS, 0u);
                umma_commit(&sfull[buf]);
            };
            const int pro = total < 3 ? total : 3;
            for (int i = 0; i < pro; ++i) issue_S();
            int qt = 0, j = 0, buf = 0;
            uint32_t use = 0;                       // parity of the buffer's use count: (i / 3) & 1
            for (int i = 0; i < total; ++i) {
                atc_wait(&pready[buf], use);
                tc_fence_after();
                if (j == 0 && qt >= 2) { atc_wait(&oread[qt & 1], (uint32_t)((qt >> 1) - 1) & 1u); tc_fence_after(); }
                const uint32_t d = tmem_base + (uint32_t)(ATC_O_COL0 + 16 * (qt & 1));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_f16kind_ts(d, tmem_base + (uint32_t)(buf * ATC_S_COLS + 32 * (kk >> 1) + 8 * (kk & 1)), atc_desc(v0 + (uint32_t)(j * 2048 + kk * 32)), idPV,
                                    (j | kk) ? 1u : 0u);
                if (j == ntl - 1) umma_commit(&ofull[qt & 1]);
                if (i + 3 < total) {
                    // The next S tile lands on the buffer whose P the PV product above still reads.  tcgen05.mma operations of
                    // one thread execute in issue order, so the write-after-read needs no barrier; ATC_WAIT_PV keeps the
                    // conservative variant (commit + wait) for A/B checks.
                    if (ATC_WAIT_PV) {
                        umma_commit(&pvdone[buf]);
                        atc_wait(&pvdone[buf], use);
                        tc_fence_after();
                    }
                    issue_S();
                }
                if (++j == ntl) { j = 0; ++qt; }
                if (++buf == 3) { buf = 0; use ^= 1u; }
            }
        }
    } else if (warp < ATC_SM_WARPS) {
        // ===================== softmax: two threads per query row (one per warpgroup, 32 keys of every tile each) =====================
        const int wg = warp >> 2, wq = warp & 3;          // TMEM lanes 32 wq .. 32 wq + 31 are this warp's
        const int row = wq * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16);
        float m = 0.f;
        auto finish_qtile = [&](int qt) {
            atc_wait(&ofull[qt & 1], (uint32_t)(qt >> 1) & 1u);
            tc_fence_after();
            uint32_t o[16];
            tmem_ld_32x32b_x16b(lane_base + (uint32_t)(ATC_O_COL0 + 16 * (qt & 1)), o);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&oread[qt & 1]);
            const float l = __uint_as_float(o[8]);
            bool bad = !(fabsf(l) <= 3.0e38f) || !(l > 0.f);
            float v[8];
            const float inv = 1.0f / l;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float oc = __uint_as_float(o[c]);
                bad = bad || !(fabsf(oc) <= 3.0e38f);
                v[c] = oc * inv;
            }
            store8(out + ((size_t)n * S + qt * 128 + row) * C + head * 8, v);
            const bool anybad = __any_sync(0xffffffffu, bad);
            if (lane == 0) flags[(((size_t)n * gridDim.x + head) * nqt + qt) * 4 + wq] = anybad ? 1 : 0;
        };
        int qt = 0, j = 0, buf = 0;
        uint32_t use = 0;                           // (i / 3) & 1, advanced incrementally (no division by the runtime ntl)
        for (int i = 0; i < total; ++i) {
            atc_wait(&sfull[buf], use);
            tc_fence_after();
            // both warpgroups work on the SAME tile: warpgroup wg takes its keys 32 wg .. 32 wg + 31 (S columns 32 wg .. 32 wg + 31; its P
            // overwrites the first 16 of those), so two whole tiles stay queued ahead of the softmax threads on the three buffers
            const uint32_t ts = lane_base + (uint32_t)(buf * ATC_S_COLS);
            uint32_t s[32], p[16];
            if (j == 0) {
                // exact row max of key tile 0 -> Q' (so every later S tile arrives as s - m); this tile subtracts in registers.
                // Both warpgroups read all 64 columns and get the same m; warpgroup 0 publishes it.
                tmem_ld_32x32b_x32(ts + (uint32_t)(32 * (wg ^ 1)), s);
                tmem_ld_wait();
                float mx = __uint_as_float(s[0]);
#pragma unroll
                for (int c = 1; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
                tmem_ld_32x32b_x32(ts + (uint32_t)(32 * wg), s);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
                m = mx;
                if (wg == 0) {
                    const T mh = from_f<T>(-m);
                    const T ml = from_f<T>(-m - to_f(mh));
                    *reinterpret_cast<uint4*>(smQ + (size_t)(qt >> 2) * 16384 + atc_tile_off(row, 2 * (qt & 3) + 1)) =
                        make_uint4(atc_pack_raw<T>(mh, ml), 0u, 0u, 0u);
                    fence_proxy_async();
                    mbar_arrive(qmready);
                }
                atc_exp32<T, PP>(s, m, p);
                // the other warpgroup may still be reading these S columns for its max: P is stored only after both are done
                asm volatile("bar.sync 1, 256;" ::: "memory");
            } else {
                tmem_ld_32x32b_x32(ts + (uint32_t)(32 * wg), s);
                tmem_ld_wait();
                atc_exp32_nosub<T, PP>(s, p);
            }
            tmem_st_32x32b_x16(ts + (uint32_t)(32 * wg), p);      // P of keys 32 wg.. lands on the first 16 of this warpgroup's own S columns
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&pready[buf]);
            // O of the PREVIOUS query tile is collected one key tile into this one: waiting for it right after its last P
            // would deadlock (the MMA thread may be parked on qmready of the next query tile, which this thread signals)
            if (j == 1 && qt >= 1 && wg == 1) finish_qtile(qt - 1);
            if (++j == ntl) { j = 0; ++qt; }
            if (++buf == 3) { buf = 0; use ^= 1u; }
        }
        if (wg == 1) finish_qtile(nqt - 1);      // warpgroup 1 collects every O
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ATC_SM_WARPS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, ATC_TMEM_COLS);
    }
}

size_t attention_tc_smem_bytes(int S) { return (size_t)((S + 511) >> 9) * 32768 + (size_t)S * 32 + 256 + 1024; }

template <typename T, int PP>
static int launch_tc(const void* qkv, int N, int S, int C, float qmul, void* out, uint8_t* flags, cudaStream_t s) {
    const size_t smem = attention_tc_smem_bytes(S);
    static size_t attr_dev[PD_MAX_DEVICES] = {0};
    size_t& attr = attr_dev[pd_cur_dev()];
    if (smem > attr) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<T, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid(C / 8, N);
    attention_tc_kernel<T, PP><<<grid, ATC_THREADS, smem, s>>>((const T*)qkv, S, C, qmul, (T*)out, flags);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_attention_tc(int dt, const void* qkv, int N, int S, int C, float qmul, void* out, uint8_t* flags, int poly_pairs, cudaStream_t s) {
#define PD_ATC(PP) PD_DISPATCH_HALF(dt, T, { return launch_tc<T, PP>(qkv, N, S, C, qmul, out, flags, s); })
    switch (poly_pairs) {
        case 0: PD_ATC(0); break;
        case 2: PD_ATC(2); break;
        case 3: PD_ATC(3); break;
        case 5: PD_ATC(5); break;
        default: PD_ATC(4); break;
    }
#undef PD_ATC
    return 0;
}

}  // namespace pd
