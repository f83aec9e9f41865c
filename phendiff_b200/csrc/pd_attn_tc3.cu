// phendiff_b200 — flash-style self-attention for head_dim 8 on tcgen05 / TMEM / TMA, third design ("tc3"):
// softmax(q k^T / sqrt(8)) v on packed qkv (N, S, 3C) -> (N, S, C); 16-bit in, fp32 accumulate, 16-bit out
// (diffusers Attention + AttnProcessor2_0 core as instantiated by cond_unet_2d.py:166-227).
//
// What the measurements of the first two tcgen05 kernels said (profiles/r3_attention_notes.md, profiles/r4a_*): they were
// latency-starved, not throughput-bound — 1.4 eligible warps per scheduler, the two softmax warpgroups of a CTA sharing one
// stream of S tiles behind one issuer round trip per tile, a per-thread scatter staging phase, and TMEM reads nowhere near a
// limit (tools/microbench/tmem_ld.cu: 470 .. 900 B/clk/SM against the 64 B/clk/SM the exponentials need).  This kernel
// removes the coupling:
//   * PERSISTENT, one CTA per SM, 512 TMEM columns: three softmax warpgroups, each owning its OWN stream of query tiles
//     (128 rows x all keys) with a private double-buffered S/P tile (2 x 64 columns) and a private double-buffered O
//     accumulator (2 x 16 columns) — no barrier between warpgroups, no shared tile;
//   * one MMA-issuer thread multiplexes the three streams with non-blocking barrier probes; S for step s+1 is already in TMEM
//     while the warpgroup exponentiates step s, and S for s+2 is issued right behind the PV product of s (tcgen05 ops of one
//     thread execute in order, so the write-after-read on the buffer needs no barrier);
//   * Q, K, V of a head arrive by TMA exactly as they lie in memory — [token][8 halves] = 16-byte rows, NO swizzle — into a
//     2-deep ring of heads (the next head loads while this one computes).  The UMMA "interleave" (no-swizzle) canonical
//     layouts make those packed rows directly usable (CUTLASS cute/atom/mma_traits_sm100.hpp:169-199):
//        K-major  ((8,m),(T,2)):((1T,SBO),(1,LBO)) : 8 tokens x 16 B = one core matrix, SBO = 128 B between 8-token groups,
//                                                    LBO = distance to the SECOND 16-byte K chunk — which we point at a
//                                                    small separate "augmentation" block instead of interleaving it;
//        MN-major ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) : V as [key][8 dims] is the MN-major B operand of the PV product
//                                                    (LBO = 128 B between 8-key groups, SBO = distance to the second
//                                                    8-wide N chunk = a constant ones block -> column 8 of O is the softmax
//                                                    denominator, summed from the same rounded P the numerator uses).
//     Augmentation blocks: K' = (k, 1, 1, 0..) for every key; Q' = (q, -m_hi, -m_lo, 0..) once the row max m of the first key
//     tile is known (written by the softmax threads, 2 KB per warpgroup), zeros before: S = q.k - m leaves the tensor core
//     with no per-score instruction (hi + lo 16-bit split: |error| <= 2^-22 |m|).
//   * softmax thread = one query row (TMEM lane), 64 scores per tile in two x32 halves with the second half's tcgen05.ld in
//     flight under the first half's exponentials; P goes back to TMEM as packed 16-bit pairs = the A operand of the PV
//     product.  Exponentials: `PP` of every 8 pairs on the FMA / ALU pipes (packed-half cubic), the rest on MUFU.EX2.
// Numerics are those of pd_attn_mma.cu v3 / pd_attn_tc.cu: scores in log2 units (q pre-scaled at finalize), row max fixed
// after key tile 0, a query tile whose 16-bit P overflowed (non-finite or non-positive denominator) is flagged and recomputed
// by the exact warp-level kernel (repair pass in launch_attention_mma).
#include "pd_attn_common.cuh"
#include "pd_tc_common.cuh"

namespace pd {

constexpr int A3_NWG = 3;                          // softmax warpgroups = independent query-tile streams
__host__ __device__ constexpr int a3_threads(int sp) { return 128 + 128 * A3_NWG * sp; }   // warps 0-3: TMA producer + one MMA issuer per stream
constexpr int A3_TMEM_COLS = 512;
constexpr int A3_WG_COLS = 160;                    // S0 [0,64) | S1 [64,128) | O0 [128,144) | O1 [144,160)
constexpr int A3_KT = 64;                          // keys per S tile

struct A3Params {
    CUtensorMap tm;        // qkv viewed as {3C, S, N} 16-bit, box {8, tok_box, 1}, no swizzle
    int S, C, heads, items, tok_box;
    void* out;
    uint8_t* flags;
    int swap_k, swap_mn;   // probe knobs: exchange the LBO / SBO roles of the K-major / MN-major descriptors
    int nws;               // active streams (diagnosis knob PHENDIFF_B200_ATTN_TC3_STREAMS; default A3_NWG)
};

__device__ __forceinline__ void a3_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void a3_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void a3_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void a3_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead of
// spinning — a spinning softmax warp takes issue slots from the two other streams' warps on its scheduler (the r4d capture had
// 2.5 try_wait executions per wait and the spin loop among the top-sampled instructions)
__device__ __forceinline__ bool a3_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
// bounded wait (a protocol bug must trap, not hang the box); the 64-bit clock is read once per 64 polls
__device__ __forceinline__ void a3_wait(uint64_t* bar, uint32_t parity) {
    if (a3_try_wait_hint(bar, parity, 2000u)) return;
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 64; ++k)
            if (a3_try_wait_hint(bar, parity, 2000u)) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("phendiff_b200: attention_tc3 mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// the same for a whole warp whose lanes all need the barrier: ONE lane polls (a 32-lane SYNCS.PHASECHK costs ~270 cycles even on
// a completed barrier, r4h trace), the others park at the convergence barrier
__device__ __forceinline__ void a3_wait_warp(uint64_t* bar, uint32_t parity) {
    if (elect_one()) {
        // non-blocking probes first: try_wait is a potentially-blocking instruction (r4h trace: ~300 cycles on an already completed
        // phase when issued right behind a tcgen05.commit)
        bool ok = mbar_test(bar, parity);
#pragma unroll 1
        for (int k = 0; k < 4096 && !ok; ++k) ok = mbar_test(bar, parity);
        if (!ok) a3_wait(bar, parity);
    }
    __syncwarp();
}

// bare spin on a 32-bit shared barrier address (issuer warps only: the softmax warps keep the bounded waits, so a protocol bug
// still traps the grid instead of hanging it)
__device__ __forceinline__ void a3_spin(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tA3_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra A3_DONE;\n\tbra A3_WAIT;\n\tA3_DONE:\n\t}\n"
        ::"r"(bar), "r"(parity) : "memory");
}

// no-swizzle ("interleave") shared-memory matrix descriptor: start >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 | version 1
__device__ __forceinline__ uint64_t a3_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor, kind::f16: D fp32, A/B format (0 fp16, 1 bf16), B major (bit 16: 1 = MN-major), N >> 3, M = 128
template <typename T, int N, bool B_MN> __device__ __forceinline__ constexpr uint32_t a3_idesc() {
    constexpr uint32_t fmt = std::is_same<T, bf16>::value ? 1u : 0u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
template <typename T> __device__ __forceinline__ uint32_t a3_pack_raw(T lo, T hi) {
    return (uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
}

// exponentials of 32 scores (one thread's row x 32 keys) -> 16 packed pairs; PP of every 8 pairs go to the polynomial
template <typename T, int PP, bool SUB>
__device__ __forceinline__ void a3_exp32(const uint32_t (&s)[32], float sub, uint32_t (&p)[16]) {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        float x0 = __uint_as_float(s[2 * r]), x1 = __uint_as_float(s[2 * r + 1]);
        if (SUB) { x0 -= sub; x1 -= sub; }
        if (PP == 9) { p[r] = pack2<T>(fmaf(fabsf(x0), 1e-4f, 1e-3f), fmaf(fabsf(x1), 1e-4f, 1e-3f)); continue; }   // diagnosis only (PHENDIFF_B200_ATTN_POLYPAIRS=9): no exponentials, wrong results
        const bool poly = PP > 0 && ((r * PP) & 7) < PP;
        p[r] = poly ? ex2_pair_poly<T>(x0, x1) : pack2<T>(ex2(x0), ex2(x1));
    }
}

// SP = softmax warps per TMEM lane quarter of a stream: 1 = a thread owns a query row and all 64 columns of a tile; 2 = two
// threads (in two warpgroups) own a row and 32 columns each — twice the warps per scheduler (6 instead of 3) for latency cover,
// half the registers per thread, one extra exchange of the row max per query tile.
template <typename T, int PP, int SP>
__global__ void __launch_bounds__(a3_threads(SP), 1) attention_tc3_kernel(const __grid_constant__ A3Params p) {
    constexpr int A3_THREADS = a3_threads(SP);
    extern __shared__ uint8_t a3_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a3_smem_raw) + 127) & ~(uintptr_t)127);
    const int S = p.S, C = p.C;
    const uint32_t opb = (uint32_t)S * 16u;            // bytes of one operand (Q, K or V) of one head
    const uint32_t stage_bytes = 3u * opb;
    uint8_t* sm_kaug = smem + 2 * (size_t)stage_bytes;  // [64 keys][16 B] = (1, 1, 0 x 6)
    uint8_t* sm_vaug = sm_kaug + 1024;                  // [16 keys][16 B] = (1, 0 x 7)
    uint8_t* sm_zero = sm_vaug + 256;                   // [128 rows][16 B] zeros: Q augmentation of key tile 0
    uint8_t* sm_qaug = sm_zero + 2048;                  // [3 warpgroups][128 rows][16 B] = (-m_hi, -m_lo, 0 x 6)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_qaug + A3_NWG * 2048);
    uint64_t* qkv_full = bars;                 // [2] TMA -> MMA
    uint64_t* qkv_empty = bars + 2;            // [2] MMA -> TMA (one arrival per query tile of the head)
    uint64_t* sfull = bars + 4;                // [3][2] MMA -> softmax: S tile landed
    uint64_t* pready = bars + 10;              // [3][2] softmax -> MMA: P written, S buffer consumed
    uint64_t* ofull = bars + 16;               // [3][2] MMA -> softmax: O of a query tile complete
    uint64_t* oread = bars + 22;               // [3][2] softmax -> MMA: O slot read
    uint64_t* qmready = bars + 28;             // [3]    softmax -> MMA: -m written into the Q augmentation
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 31);
    float* sm_xmax = reinterpret_cast<float*>(bars + 32);   // [3 streams][2 halves][128 rows]: row max exchange (SP == 2)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nqt = S >> 7, ntl = S / A3_KT;
#ifdef A3_TRACE
    __shared__ long long tr_soft[40][5];    // stream 0, warp 4 lane 0: after sfull wait / after loads / after exps / after st wait / after arrive
    __shared__ long long tr_iss[40][4];     // issuer of stream 0: after pready wait / after PV issue / after S issue
    __shared__ long long tr_w[40][4][2];    // stream 0, warps 4..7 lane 0: sfull wake / pready arrive
    __shared__ long long tr_mma[40][6];     // issuer of stream 0: clock after each PV MMA, after the commits
#endif
    const int my_items = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_T = my_items * nqt;        // query tiles of this CTA, in order (item-major)

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&qkv_full[i], 1); mbar_init(&qkv_empty[i], (uint32_t)nqt); }
        for (int i = 0; i < A3_NWG * 2; ++i) { mbar_init(&sfull[i], 1); mbar_init(&pready[i], 4 * SP); mbar_init(&ofull[i], 1); mbar_init(&oread[i], 4); }
        for (int i = 0; i < A3_NWG; ++i) mbar_init(&qmready[i], 4);   // ONE arrival per warp (after __syncwarp): 128 per-thread arrivals on
                                                                       // one barrier serialise in the SYNCS unit (~350 cycles per step, r4 trace)
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) prefetch_tmap(&p.tm);
    if (warp == 2) {
        tmem_alloc(tmem_slot, A3_TMEM_COLS);
        tmem_relinquish();
    }
    {   // constant augmentation blocks (generic-proxy writes, read by the tensor core's async proxy)
        const T one = from_f<T>(1.0f), zero = from_f<T>(0.0f);
        const uint32_t ones2 = a3_pack_raw<T>(one, one), one1 = a3_pack_raw<T>(one, zero);
        for (int i = threadIdx.x; i < 64; i += A3_THREADS) *reinterpret_cast<uint4*>(sm_kaug + i * 16) = make_uint4(ones2, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < 16; i += A3_THREADS) *reinterpret_cast<uint4*>(sm_vaug + i * 16) = make_uint4(one1, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < 128 * (1 + A3_NWG); i += A3_THREADS) *reinterpret_cast<uint4*>(sm_zero + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer: Q, K, V of one (image, head) per stage =====================
            for (int it = 0; it < my_items; ++it) {
                const int stage = it & 1;
                a3_wait(&qkv_empty[stage], (uint32_t)(((it >> 1) & 1) ^ 1));
                const int item = (int)blockIdx.x + it * (int)gridDim.x;
                const int n = item / p.heads, head = item - n * p.heads;
                mbar_arrive_expect_tx(&qkv_full[stage], stage_bytes);
                uint8_t* dst = smem + (size_t)stage * stage_bytes;
                for (int which = 0; which < 3; ++which)
                    for (int t0 = 0; t0 < S; t0 += p.tok_box)
                        tma_load_3d(&p.tm, &qkv_full[stage], dst + (size_t)which * opb + (size_t)t0 * 16, which * C + head * 8, t0, n);
            }
        }
    } else if (warp < 4) {
        // ===================== MMA issuers: one WARP per stream (warps 1, 2, 3) =====================
        // History (cycle traces, profiles/r4_attention_tc3.md): a single thread multiplexing the three streams spent ~1450
        // cycles of scalar work per tile step (2.8 ms per launch); one issuer THREAD per stream still ~1400 (inside an
        // `if (lane == 0)` region ptxas treats every tcgen05 operand as divergent and wraps each UTCHMMA / UTCBAR in an
        // ELECT + 4 x R2UR.BROADCAST loop); a whole warp with bounded, watchdogged waits ~1000 (300 SASS instructions and a
        // dozen branches per step).  This version is written for instruction count: the WHOLE warp runs warp-uniform code
        // (stream index from a shuffle so that ptxas can prove it), the inner loop over key tiles is unrolled by two so that
        // buffer indices and barrier addresses are constants, descriptors are running 32-bit words, waits are bare try_wait
        // spins (the softmax warps keep the bounded waits: a protocol bug still traps instead of hanging).
        const int g = __shfl_sync(0xffffffffu, warp, 0) - 1;
        const int nTg = (g >= 0 && g < p.nws && total_T > g) ? (total_T - g + p.nws - 1) / p.nws : 0;
        if (nTg > 0) {
            constexpr uint32_t idS = a3_idesc<T, A3_KT, false>(), idPV = a3_idesc<T, 16, true>();
            constexpr uint32_t HI_K = (128u >> 4) | (1u << 14);          // K-major operands: SBO = 128 B, descriptor version 1
            const uint32_t sm0 = smem_u32(smem), kaug = smem_u32(sm_kaug), vaug = smem_u32(sm_vaug), zaug = smem_u32(sm_zero);
            const uint32_t qaug = smem_u32(sm_qaug) + (uint32_t)g * 2048u;
            const uint32_t tS = tmem_base + (uint32_t)(g * A3_WG_COLS), tO = tS + 128u;
            const uint32_t bar_sfull = smem_u32(&sfull[g * 2]), bar_pready = smem_u32(&pready[g * 2]);
            const uint32_t bar_ofull = smem_u32(&ofull[g * 2]), bar_oread = smem_u32(&oread[g * 2]);
            const uint32_t bar_qm = smem_u32(&qmready[g]), bar_full = smem_u32(qkv_full), bar_empty = smem_u32(qkv_empty);
            auto lo_k = [](uint32_t start, uint32_t aug) { return ((start & 0x3FFFFu) >> 4) | ((((aug - start) >> 4) & 0x3FFFu) << 16); };
            // one S tile: S[buf] = Q' K'^T (fresh accumulator), then its completion is committed to sfull[buf]
            auto mma_S = [&](int buf, uint32_t a_lo, uint32_t b_lo) {
                if (elect_one()) {
                    umma_f16kind(tS + (uint32_t)(buf * 64), desc64(HI_K, a_lo), desc64(HI_K, b_lo), idS, 0u);
                    umma_commit_addr(bar_sfull + buf * 8);
                }
                __syncwarp();
            };
            // one PV product: O (+)= P[buf] V'(64 keys) as four K = 16 MMAs; MN-major B: LBO = 128 B between 8-key groups, SBO =
            // distance to the ones block (running: + 256 B start, - 256 B SBO per 16 keys)
            auto mma_PV = [&](int buf, uint32_t d, uint32_t v_lo, uint32_t v_hi, bool first, bool last, uint32_t ob, uint32_t eb) {
                if (elect_one()) {
                    const uint32_t a0 = tS + (uint32_t)(buf * 64);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        a3_mma_ts(d, a0 + (uint32_t)(32 * (kk >> 1) + 8 * (kk & 1)), desc64(v_hi - 16u * kk, v_lo + 16u * kk), idPV, (!first || kk) ? 1u : 0u);
                    if (last) { umma_commit_addr(ob); umma_commit_addr(eb); }
                }
                __syncwarp();
            };
            int it = 0, qt = g, items_seen = 0;
            while (qt >= nqt) { qt -= nqt; ++it; }
            // S tile 0 of the first query tile (zero augmentation: needs only the head's Q / K / V)
            a3_spin(bar_full + (it & 1) * 8, (uint32_t)(it >> 1) & 1u);
            items_seen = it + 1;
            tc_fence_after();
            uint32_t stage0 = sm0 + (uint32_t)(it & 1) * stage_bytes;
            uint32_t qs = stage0 + (uint32_t)qt * 2048u;
            mma_S(0, lo_k(qs, zaug), lo_k(stage0 + opb, kaug));
            for (int k = 0; k < nTg; ++k) {
                // this query tile: item `it` (stage it & 1), tile qt; S(k, 0) is in flight or done in buffer 0 (ntl is even)
                const uint32_t a_aug = lo_k(qs, qaug);
                uint32_t b_lo = lo_k(stage0 + opb + A3_KT * 16u, kaug);                 // K' of key tile 1
                const uint32_t vs0 = stage0 + 2u * opb;
                uint32_t v_lo = ((vs0 & 0x3FFFFu) >> 4) | ((128u >> 4) << 16);
                uint32_t v_hi = (((vaug - vs0) >> 4) & 0x3FFFu) | (1u << 14);
                const uint32_t d = tO + (uint32_t)(16 * (k & 1));
                const uint32_t ob = bar_ofull + (k & 1) * 8, eb = bar_empty + (it & 1) * 8;
                // next query tile of this stream (its S tile 0 is issued two steps before this one ends)
                int it_n = it, qt_n = qt + p.nws;
                while (qt_n >= nqt) { qt_n -= nqt; ++it_n; }
                const bool has_next = k + 1 < nTg;
                const uint32_t stage_n = sm0 + (uint32_t)(it_n & 1) * stage_bytes;
                const uint32_t qs_n = stage_n + (uint32_t)qt_n * 2048u;
                // S(k, 1) needs -m of this query tile
                a3_spin(bar_qm, (uint32_t)k & 1u);
                tc_fence_after();
                mma_S(1, a_aug, b_lo);
                b_lo += 64u - (64u << 16);
                if (k >= 2) a3_spin(bar_oread + (k & 1) * 8, (uint32_t)((k >> 1) - 1) & 1u);   // O slot of query tile k - 2 has been read
                const uint32_t par0 = (uint32_t)(k * (ntl >> 1)) & 1u;                  // parity of buffer use (k ntl / 2 + j / 2)
#pragma unroll 1
                for (int j = 0; j < ntl; j += 2) {
                    const uint32_t par = (par0 + (uint32_t)(j >> 1)) & 1u;
                    // ---- even step j: buffer 0 ----
#ifdef A3_TRACE
                    const int sT = k * ntl + j;
                    const bool trI = blockIdx.x == 0 && g == 0 && lane == 0 && sT >= 20 && sT < 60;
                    if (trI) tr_iss[sT - 20][3] = clock64();
#endif
                    a3_spin(bar_pready, par);
#ifdef A3_TRACE
                    if (trI) tr_iss[sT - 20][0] = clock64();
#endif
                    tc_fence_after();
                    mma_PV(0, d, v_lo, v_hi, j == 0, false, ob, eb);
#ifdef A3_TRACE
                    if (trI) tr_iss[sT - 20][1] = clock64();
#endif
                    v_lo += 64u; v_hi -= 64u;
                    if (j + 2 < ntl) {
                        mma_S(0, a_aug, b_lo);                                          // S(k, j + 2)
                        b_lo += 64u - (64u << 16);
                    } else if (has_next) {
                        if (it_n == items_seen) {                                       // first touch of the next item by this stream
                            a3_spin(bar_full + (it_n & 1) * 8, (uint32_t)(it_n >> 1) & 1u);
                            ++items_seen;
                            tc_fence_after();
                        }
                        mma_S(0, lo_k(qs_n, zaug), lo_k(stage_n + opb, kaug));          // S(k + 1, 0)
                    }
#ifdef A3_TRACE
                    if (trI) tr_iss[sT - 20][2] = clock64();
                    if (trI) tr_iss[sT + 1 - 20][3] = clock64();
#endif
                    // ---- odd step j + 1: buffer 1 ----
                    a3_spin(bar_pready + 8, par);
#ifdef A3_TRACE
                    if (trI) tr_iss[sT + 1 - 20][0] = clock64();
#endif
                    tc_fence_after();
                    mma_PV(1, d, v_lo, v_hi, false, j + 2 >= ntl, ob, eb);
                    v_lo += 64u; v_hi -= 64u;
#ifdef A3_TRACE
                    if (trI) tr_iss[sT + 1 - 20][1] = clock64();
#endif
                    if (j + 3 < ntl) {
                        mma_S(1, a_aug, b_lo);                                          // S(k, j + 3)
                        b_lo += 64u - (64u << 16);
                    }
#ifdef A3_TRACE
                    if (trI) tr_iss[sT + 1 - 20][2] = clock64();
#endif
                }
                it = it_n; qt = qt_n; stage0 = stage_n; qs = qs_n;
            }
        }

    } else if (warp >= 4) {
        // ===================== softmax: stream g = SP warpgroups; thread = one query row x (64 / SP) columns of every tile =====================
        const int sw = warp - 4;
        const int g = sw / (4 * SP), half = (sw >> 2) % SP, wq = warp & 3;   // TMEM lanes 32 wq .. 32 wq + 31 are this warp's
        const int row = wq * 32 + lane;
        constexpr int NC = 64 / SP;                                           // columns of a tile per thread
        const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * A3_WG_COLS);
        uint8_t* my_qaug = sm_qaug + g * 2048 + row * 16;
        T* out = reinterpret_cast<T*>(p.out);
        const int nTg = (g < p.nws && total_T > g) ? (total_T - g + p.nws - 1) / p.nws : 0;
        auto finish_qtile = [&](int k) {
            const int Tq = g + p.nws * k, it = Tq / nqt, qt = Tq - it * nqt;
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            const int n = item / p.heads, head = item - n * p.heads;
            a3_wait(&ofull[g * 2 + (k & 1)], (uint32_t)(k >> 1) & 1u);
            tc_fence_after();
            uint32_t o[16];
            a3_ld_x16(lane_base + (uint32_t)(128 + 16 * (k & 1)), o);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&oread[g * 2 + (k & 1)]);
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
            if (lane == 0) p.flags[(((size_t)n * p.heads + head) * nqt + qt) * 4 + wq] = anybad ? 1 : 0;
        };
        int s = 0;                                                  // step counter of this stream
        for (int k = 0; k < nTg; ++k) {
            float m = 0.f;
            for (int j = 0; j < ntl; ++j, ++s) {
                const int buf = s & 1;
                a3_wait(&sfull[g * 2 + buf], (uint32_t)(s >> 1) & 1u);
                tc_fence_after();
#ifdef A3_TRACE
                const bool trc = blockIdx.x == 0 && warp == 4 && lane == 0 && s >= 20 && s < 60;
                if (trc) tr_soft[s - 20][0] = clock64();
                const bool trw = blockIdx.x == 0 && g == 0 && lane == 0 && s >= 20 && s < 60;
                if (trw) tr_w[s - 20][wq][0] = clock64();
#endif
                const uint32_t ts = lane_base + (uint32_t)(buf * 64 + half * NC);
                uint32_t s0[32], p0[16];
                tmem_ld_32x32b_x32(ts, s0);
                tmem_ld_wait();
                if (SP == 1) {
                    uint32_t s1[32], p1[16];
                    tmem_ld_32x32b_x32(ts + 32u, s1);               // in flight under the first half's exponentials
                    if (j == 0) {
                        // exact row max of key tile 0 -> Q augmentation (every later S tile of this query tile arrives as s - m);
                        // this tile subtracts in registers
                        tmem_ld_wait();
                        float mx = __uint_as_float(s0[0]);
#pragma unroll
                        for (int c = 1; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(s0[c]));
#pragma unroll
                        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(s1[c]));
                        m = mx;
                        const T mh = from_f<T>(-m);
                        const T ml = from_f<T>(-m - to_f(mh));
                        *reinterpret_cast<uint4*>(my_qaug) = make_uint4(a3_pack_raw<T>(mh, ml), 0u, 0u, 0u);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&qmready[g]);
                        a3_exp32<T, PP, true>(s0, m, p0);
                        a3_st_x16(ts, p0);
                        a3_exp32<T, PP, true>(s1, m, p1);
                        a3_st_x16(ts + 32u, p1);
                    } else {
                        a3_exp32<T, PP, false>(s0, 0.f, p0);
                        a3_st_x16(ts, p0);                          // P of keys 0..31 lands on the first 16 of their own S columns
                        tmem_ld_wait();
#ifdef A3_TRACE
                        if (trc) tr_soft[s - 20][1] = clock64();
#endif
                        a3_exp32<T, PP, false>(s1, 0.f, p1);
                        a3_st_x16(ts + 32u, p1);
#ifdef A3_TRACE
                        if (trc) tr_soft[s - 20][2] = clock64();
#endif
                    }
                } else {
                    if (j == 0) {
                        float mx = __uint_as_float(s0[0]);
#pragma unroll
                        for (int c = 1; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(s0[c]));
                        // the other 32 columns of this row belong to the thread `row` of the stream's other warpgroup
                        sm_xmax[(g * 2 + half) * 128 + row] = mx;
                        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(128 * SP) : "memory");
                        m = fmaxf(mx, sm_xmax[(g * 2 + (half ^ 1)) * 128 + row]);
                        if (half == 0) {
                            const T mh = from_f<T>(-m);
                            const T ml = from_f<T>(-m - to_f(mh));
                            *reinterpret_cast<uint4*>(my_qaug) = make_uint4(a3_pack_raw<T>(mh, ml), 0u, 0u, 0u);
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&qmready[g]);
                        }
                        a3_exp32<T, PP, true>(s0, m, p0);
                    } else {
                        a3_exp32<T, PP, false>(s0, 0.f, p0);
                    }
                    a3_st_x16(ts, p0);                              // P of this thread's 32 keys lands on the first 16 of their own S columns
                }
                a3_st_wait();
#ifdef A3_TRACE
                if (trc) tr_soft[s - 20][3] = clock64();
#endif
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&pready[g * 2 + buf]);
#ifdef A3_TRACE
                if (trc) tr_soft[s - 20][4] = clock64();
                if (trw) tr_w[s - 20][wq][1] = clock64();
#endif
                // O of the PREVIOUS query tile of this stream: its last PV product was issued when this stream handed over the
                // last P, a whole tile ago
                if (j == 0 && k >= 1 && half == 0) finish_qtile(k - 1);
            }
        }
        if (nTg > 0 && half == 0) finish_qtile(nTg - 1);
    }
    tc_fence_before();
    __syncthreads();
#ifdef A3_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int i = 0; i < 40; ++i) {
            printf("step %2d soft(w4..7): wake %lld %lld %lld %lld arrive %lld %lld %lld %lld || iss: top %lld spin-done %lld (+%lld) PV-done +%lld S-done +%lld\n", 20 + i,
                   tr_w[i][0][0] - tr_soft[0][0], tr_w[i][1][0] - tr_soft[0][0], tr_w[i][2][0] - tr_soft[0][0], tr_w[i][3][0] - tr_soft[0][0],
                   tr_w[i][0][1] - tr_soft[0][0], tr_w[i][1][1] - tr_soft[0][0], tr_w[i][2][1] - tr_soft[0][0], tr_w[i][3][1] - tr_soft[0][0],
                   tr_iss[i][3] - tr_soft[0][0], tr_iss[i][0] - tr_soft[0][0], tr_iss[i][0] - tr_iss[i][3], tr_iss[i][1] - tr_iss[i][0], tr_iss[i][2] - tr_iss[i][1]);
        }
    }
#endif
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, A3_TMEM_COLS);
    }
}

size_t attention_tc3_smem_bytes(int S) { return (size_t)2 * 3 * S * 16 + 1024 + 256 + 2048 + A3_NWG * 2048 + 32 * 8 + A3_NWG * 2 * 128 * 4 + 128; }

bool attention_tc3_supported(int S, int C) {
    // >= 3 query tiles per head: every stream then has work in every head, which the head-ring barrier parities rely on
    return S % 128 == 0 && S >= 128 * A3_NWG && C % 8 == 0 && attention_tc3_smem_bytes(S) <= 227 * 1024;
}

template <typename T, int PP, int SP>
static int launch_tc3(const A3Params& p, size_t smem, int grid, cudaStream_t s) {
    static size_t attr_dev[PD_MAX_DEVICES] = {0};
    size_t& attr = attr_dev[pd_cur_dev()];
    if (smem > attr) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(attention_tc3_kernel<T, PP, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    attention_tc3_kernel<T, PP, SP><<<grid, a3_threads(SP), smem, s>>>(p);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// q must already carry log2(e) / sqrt(d) (the finalize-time fold of pd_api.cu): TMA stages the rows as they are
int launch_attention_tc3(int dt, const void* qkv, int N, int S, int C, void* out, uint8_t* flags, int poly_pairs, cudaStream_t s) {
    PD_REQUIRE(attention_tc3_supported(S, C), "attention_tc3: S must be a multiple of 128 and fit two heads in shared memory");
    A3Params p;
    memset(&p, 0, sizeof(p));
    p.S = S; p.C = C; p.heads = C / 8; p.items = N * p.heads; p.out = out; p.flags = flags;
    p.tok_box = S < 256 ? S : 256;
    PD_REQUIRE(S % p.tok_box == 0, "attention_tc3: S must be a multiple of the TMA token box");
    {
        const char* e = getenv("PHENDIFF_B200_ATTN_TC3_SWAP");   // probe knob: bit 0 swaps LBO / SBO of the K-major descriptors, bit 1 of the MN-major one
        const int sw = e ? atoi(e) : 0;
        p.swap_k = sw & 1; p.swap_mn = (sw >> 1) & 1;
        const char* ns = getenv("PHENDIFF_B200_ATTN_TC3_STREAMS");
        p.nws = ns ? std::min(A3_NWG, std::max(1, atoi(ns))) : A3_NWG;
    }
    const uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)S, (uint64_t)N};
    const uint64_t strides[2] = {(uint64_t)3 * C * 2, (uint64_t)S * 3 * C * 2};
    const uint32_t box[3] = {8, (uint32_t)p.tok_box, 1};
    int rc = tc_encode_map(&p.tm, dt, qkv, 3, dims, strides, box, false);
    if (rc) return rc;
    const size_t smem = attention_tc3_smem_bytes(S);
    const int grid = std::min(p.items, tc_num_sms());
    static const int split = [] { const char* e = getenv("PHENDIFF_B200_ATTN_TC3_SPLIT"); return e ? atoi(e) : 2; }();
#define PD_A3(PP) PD_DISPATCH_HALF(dt, T, { return split == 1 ? launch_tc3<T, PP, 1>(p, smem, grid, s) : launch_tc3<T, PP, 2>(p, smem, grid, s); })
    switch (poly_pairs) {
        case 0: PD_A3(0); break;
        case 2: PD_A3(2); break;
        case 3: PD_A3(3); break;
        case 5: PD_A3(5); break;
        case 8: PD_A3(8); break;
        case 9: PD_A3(9); break;
        default: PD_A3(4); break;
    }
#undef PD_A3
    return 0;
}

}  // namespace pd
