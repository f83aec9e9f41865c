// phendiff_b200 — implicit-GEMM convolution / linear on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// GEMM view: M = N*Ho*Wo output pixels, Ncol = Cout, K = k*k*C (+ Csc of a fused 1x1 shortcut).
//   A (activations, NHWC bf16): never materialised as im2col.  An M tile is a box of 128 output pixels
//     (Wt x Ht x Nt); for filter tap (r,s) the A tile is the SAME box shifted by (r-pad, s-pad) in the input, i.e. one
//     4-D tiled TMA load {64 ch, Wt, Ht, Nt} whose out-of-bounds rows/columns the TMA unit zero-fills (= conv padding).
//     Each pixel is one 128-byte row (64 bf16), written with the 128B swizzle -> exactly the canonical K-major
//     SWIZZLE_128B UMMA operand.  Stride-2 convs (Downsample2D) use a 5-D "phase" view of the input
//     {(b,c), W/2, a, H/2, N} so that tap (r,s) is again a plain box.
//   B (weights): (Cout, Ktot) bf16 K-major, Ktot ordered (tap, channel) then the shortcut channels; 2-D TMA boxes.
//   D: fp32 accumulator in TMEM, 128 lanes x BLOCK_N columns, double buffered so the epilogue of tile i overlaps the
//     MMAs of tile i+1.
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> +bias +time-embedding row +residual ->
// bf16 -> global).
//
// This is the per-tap kernel: it serves the stride-2 convolutions (Downsample2D) and shapes the halo kernel
// (pd_conv_halo.cu, which re-uses one activation tile for all taps) does not tile; conv_tc_plan_create picks.
#include "pd_tc_common.cuh"
#include <algorithm>
#include <cstring>
#include <string>

namespace pd {

struct ConvTcParams {
    CUtensorMap tmA;    // main activation view (4-D, or 5-D phase view when stride2)
    CUtensorMap tmS1;   // shortcut source 1 (4-D, output resolution)
    CUtensorMap tmS2;   // shortcut source 2
    CUtensorMap tmB;    // weights (Ktot, Cout)
    CUtensorMap tmBh;   // 2-CTA kernel: the same with a box of BLOCK_N / 2 rows (each CTA of a pair loads its half of the N tile)
    CUtensorMap tmOut;  // 2-CTA kernel: output as (Cout, pixels), box {64 channels, 32 pixels}: one TMA store per epilogue slab
    int ksize, pad, stride2, C;
    int kb_main;        // 64-channel blocks per tap
    int kb_s1, kb_s2;   // 64-channel blocks of the shortcut segments
    int num_kb;
    int Wt, Ht, Nt, tilesW, tilesH;
    int Ho, Wo, Cout;
    int m_tiles, n_tiles;
    TcEpi epi;
};

struct ConvTapPlan {
    ConvTcParams p;
    int dt;
    int block_n;
    int grid;
    size_t smem;
    bool pair = false;   // 1x1 layers on the 2-CTA kernel (conv_pair_kernel)
};

constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KiB

template <int BLOCK_N> struct TcCfg {
    static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;   // power of two for 64/128/256
    static constexpr size_t SMEM = (size_t)STAGES * (TC_A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, typename T>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    using Cfg = TcCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = smem;
    uint8_t* smB = smem + (size_t)STAGES * TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * (TC_A_BYTES + Cfg::B_BYTES));
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmA);
        prefetch_tmap(&p.tmB);
        if (p.kb_s1) prefetch_tmap(&p.tmS1);
        if (p.kb_s2) prefetch_tmap(&p.tmS2);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.m_tiles * p.n_tiles;
    const int tiles_per_group = p.tilesW * p.tilesH;

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
                const int grp = m_tile / tiles_per_group, rem = m_tile - grp * tiles_per_group;
                const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
                const int n0 = grp * p.Nt, h0 = th * p.Ht, w0 = tw * p.Wt;
                const int ntap = p.ksize * p.ksize;
                const int kb_taps = ntap * p.kb_main;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], TC_A_BYTES + Cfg::B_BYTES);
                    uint8_t* dstA = smA + (size_t)stage * TC_A_BYTES;
                    uint8_t* dstB = smB + (size_t)stage * Cfg::B_BYTES;
                    if (kb < kb_taps) {
                        const int tap = kb / p.kb_main, c0 = (kb - tap * p.kb_main) * TC_BLOCK_K;
                        const int r = tap / p.ksize, s = tap - r * p.ksize;
                        if (!p.stride2) {
                            tma_load_4d(&p.tmA, &full[stage], dstA, c0, w0 + s - p.pad, h0 + r - p.pad, n0);
                        } else {
                            // input row 2*ho + q, q = r - pad: phase a = q mod 2, half-resolution offset floor(q/2)
                            const int qr = r - p.pad, qs = s - p.pad;
                            const int ar = qr & 1, as = qs & 1;
                            const int dr = (qr - ar) / 2, ds = (qs - as) / 2;
                            tma_load_5d(&p.tmA, &full[stage], dstA, as * p.C + c0, w0 + ds, ar, h0 + dr, n0);
                        }
                    } else if (kb < kb_taps + p.kb_s1) {
                        tma_load_4d(&p.tmS1, &full[stage], dstA, (kb - kb_taps) * TC_BLOCK_K, w0, h0, n0);
                    } else {
                        tma_load_4d(&p.tmS2, &full[stage], dstA, (kb - kb_taps - p.kb_s1) * TC_BLOCK_K, w0, h0, n0);
                    }
                    tma_load_2d(&p.tmB, &full[stage], dstB, kb * TC_BLOCK_K, n_tile * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // ===================== MMA issuer (single thread) =====================
            constexpr uint32_t idesc = make_idesc<T, BLOCK_N>();
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smA + (size_t)stage * TC_A_BYTES);
                    const uint32_t b_addr = smem_u32(smB + (size_t)stage * Cfg::B_BYTES);
                    const uint64_t a_desc = make_sw128_desc(a_addr);
                    const uint64_t b_desc = make_sw128_desc(b_addr);
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 address
                        umma_f16kind(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);                       // smem slot free when these MMAs retire
                    if (kb == p.num_kb - 1) umma_commit(&tfull[as]);  // accumulator ready for the epilogue
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
            const int grp = m_tile / tiles_per_group, rem = m_tile - grp * tiles_per_group;
            const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
            const int per_img = p.Wt * p.Ht;
            const int nn = row / per_img, rr = row - nn * per_img;
            const int hh = rr / p.Wt, ww = rr - hh * p.Wt;
            const int img = grp * p.Nt + nn;
            const size_t pix = ((size_t)img * p.Ho + (th * p.Ht + hh)) * p.Wo + (tw * p.Wt + ww);
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + chunk * 32), r);
                tmem_ld_wait();
                if (p.epi.stats_cw == 2) tc_epilogue_chunk32<T, false, 2>(p.epi, r, n_tile * BLOCK_N + chunk * 32, pix, img, lane);
                else tc_epilogue_chunk32<T, false, 4>(p.epi, r, n_tile * BLOCK_N + chunk * 32, pix, img, lane);
            }
            tc_fence_before();
            mbar_arrive(&tempty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps, plan, launch
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int tc_encode_map(CUtensorMap* tm, int dt, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    PFN_encodeTiled enc = get_encode();
    PD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bd[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bd[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(tm, dt == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bd, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return 2;
    }
    return 0;
}

static bool tile_geometry(int N, int Ho, int Wo, int* Wt, int* Ht, int* Nt) {
    int wt = std::min(Wo, TC_BLOCK_M);
    if (wt <= 0 || TC_BLOCK_M % wt != 0 || Wo % wt != 0) return false;
    int ht = std::min(Ho, TC_BLOCK_M / wt);
    if (ht <= 0 || (TC_BLOCK_M / wt) % ht != 0 || Ho % ht != 0) return false;
    int nt = TC_BLOCK_M / (wt * ht);
    if (nt > 1 && (ht != Ho || wt != Wo)) return false;
    if (N % nt != 0) return false;
    *Wt = wt; *Ht = ht; *Nt = nt;
    return true;
}

bool conv_tc_supported(const ConvTcDesc& d, std::string* why);
bool conv_tc_can_emit_stats(const ConvTcDesc& d);
static bool tile_geometry(int N, int Ho, int Wo, int* Wt, int* Ht, int* Nt);

// ---------------------------------------------------------------------------------------------------------------------
// 1x1 / linear layers on a CTA PAIR (tcgen05 cta_group::2, thread-block cluster of two SMs): M = 256 output pixels x N = 256 output
// channels per MMA.  Each CTA of the pair loads ITS 128-pixel A tile and ITS half (128 rows) of the weight tile, so the L2 -> SM stream
// of the weights and the shared-memory operand reads per SM are halved (the qkv / out-proj projections were bound there: DESIGN §3.4).
// Only the leader (cluster rank 0) issues MMAs; the peer's TMA completions are relayed to the leader by one thread (local mbarrier wait
// -> remote arrive through the cluster shared window); tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs; the
// peer's epilogue warps release the accumulator on the leader's barrier.  Each CTA runs the usual epilogue on its own TMEM lanes.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int PAIR_STAGES = 2, PAIR_N = 256, PAIR_KSUB = 2;       // a ring slot holds PAIR_KSUB 64-channel K blocks: 8 MMAs per barrier round
constexpr int PAIR_B_BYTES = (PAIR_N / 2) * TC_BLOCK_K * 2;        // 16 KiB: this CTA's half of the weight tile, one K block
constexpr int PAIR_SLOT_BYTES = PAIR_KSUB * (TC_A_BYTES + PAIR_B_BYTES);
constexpr int PAIR_EPI_WARPS = 8, PAIR_THREADS = 128 + 32 * PAIR_EPI_WARPS;   // two epilogue warps per TMEM lane quarter, alternating slabs
constexpr int PAIR_EPI_BYTES = PAIR_EPI_WARPS * 2 * 4096;          // two swizzled 32-row x 128-byte staging slabs per epilogue warp: the TMA
                                                                   // store of one slab overlaps the TMEM load / arithmetic of the next
constexpr size_t PAIR_SMEM = (size_t)PAIR_STAGES * PAIR_SLOT_BYTES + PAIR_EPI_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier that receives arrivals from the peer CTA (cluster-scope acquire), bounded like mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok = 0;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) {
            printf("phendiff_b200: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void umma_f16kind_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair when they complete
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) conv_pair_kernel(const __grid_constant__ ConvTcParams p) {
    constexpr int STAGES = PAIR_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // slot layout: [A k0][A k1][B k0][B k1]
    uint8_t* smEpi = smem + (size_t)STAGES * PAIR_SLOT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smEpi + PAIR_EPI_BYTES);
    uint64_t* full = bars;                    // leader: local TMA bytes + the peer's relay (2 arrivals); peer: local TMA bytes
    uint64_t* empty = bars + STAGES;          // multicast commit: slot free in both CTAs
    uint64_t* tfull = bars + 2 * STAGES;      // [2] multicast commit: accumulator ready
    uint64_t* tempty = bars + 2 * STAGES + 2; // [2] leader: 8 local + 8 remote epilogue warps released the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (warp == 0 && elect_one()) { prefetch_tmap(&p.tmA); prefetch_tmap(&p.tmBh); }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], rank == 0 ? 2 : 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * PAIR_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile = (pair of consecutive M tiles) x N tile; this CTA's M tile = 2 * pair + rank
    const int m_pairs = p.m_tiles / 2;
    const int total = m_pairs * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int tiles_per_group = p.tilesW * p.tilesH;
    const int nslots = p.num_kb / PAIR_KSUB;      // ring rounds per tile

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer (both CTAs) =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < total; tile += nclusters) {
                const int n_tile = tile % p.n_tiles, m_tile = 2 * (tile / p.n_tiles) + (int)rank;
                const int grp = m_tile / tiles_per_group, rem = m_tile - grp * tiles_per_group;
                const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
                const int n0 = grp * p.Nt, h0 = th * p.Ht, w0 = tw * p.Wt;
                const int brow = n_tile * PAIR_N + (int)rank * (PAIR_N / 2);
                for (int ks = 0; ks < nslots; ++ks) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], PAIR_SLOT_BYTES);
                    uint8_t* slot = smem + (size_t)stage * PAIR_SLOT_BYTES;
#pragma unroll
                    for (int j = 0; j < PAIR_KSUB; ++j) {
                        const int k0 = (ks * PAIR_KSUB + j) * TC_BLOCK_K;
                        tma_load_4d(&p.tmA, &full[stage], slot + (size_t)j * TC_A_BYTES, k0, w0, h0, n0);
                        tma_load_2d(&p.tmBh, &full[stage], slot + (size_t)PAIR_KSUB * TC_A_BYTES + (size_t)j * PAIR_B_BYTES, k0, brow);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        if (rank == 1 && elect_one()) {
            // ===================== peer: relay "my slot landed" to the leader's full barrier =====================
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t remote0 = mapa_u32(smem_u32(&full[0]), 0);
            for (int tile = cluster_id; tile < total; tile += nclusters)
                for (int ks = 0; ks < nslots; ++ks) {
                    mbar_wait(&full[stage], phase);
                    mbar_arrive_remote(remote0 + (uint32_t)stage * 8u);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        if (rank == 0 && elect_one()) {
            // ===================== MMA issuer (leader CTA, single thread) =====================
            constexpr uint32_t fmt = std::is_same<T, bf16>::value ? 1u : 0u;
            constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(PAIR_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            constexpr uint32_t HI = sw128_desc_hi(1024);
            const uint32_t base_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int tile = cluster_id; tile < total; tile += nclusters, ++iter) {
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait_cluster(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * PAIR_N);
                for (int ks = 0; ks < nslots; ++ks) {
                    mbar_wait_cluster(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_lo = base_lo + (uint32_t)stage * (PAIR_SLOT_BYTES >> 4);
                    const uint32_t b_lo = a_lo + ((PAIR_KSUB * TC_A_BYTES) >> 4);
#pragma unroll
                    for (int j = 0; j < PAIR_KSUB; ++j)
#pragma unroll
                        for (int k = 0; k < TC_BLOCK_K / 16; ++k)
                            umma_f16kind_pair(d_tmem, desc64(HI, a_lo + (uint32_t)j * (TC_A_BYTES >> 4) + 2u * k),
                                              desc64(HI, b_lo + (uint32_t)j * (PAIR_B_BYTES >> 4) + 2u * k), idesc, (ks | j | k) != 0 ? 1u : 0u);
                    umma_commit_pair(&empty[stage]);
                    if (ks == nslots - 1) umma_commit_pair(&tfull[as]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (4 warps per CTA, this CTA's 128 accumulator rows) =====================
        // staged like the halo kernel's: 64 columns of a warp's 32 rows go through a swizzled 32 x 128-byte slab and leave as ONE TMA
        // store (the warp's 32 rows are 32 consecutive pixels: tiles span whole image rows); the residual comes in the same way
        const int ew = warp - 4, q = ew & 3, slab0 = ew >> 2;
        constexpr int SLAB_STEP = PAIR_EPI_WARPS / 4;
        const int row = q * 32 + lane;
        const uint32_t remote_tempty = mapa_u32(smem_u32(&tempty[0]), 0);
        uint8_t* buf0 = smEpi + ew * 8192;
        int sb = 0;
        const bool has_res = p.epi.residual != nullptr;
        int iter = 0;
        for (int tile = cluster_id; tile < total; tile += nclusters, ++iter) {
            const int n_tile = tile % p.n_tiles, m_tile = 2 * (tile / p.n_tiles) + (int)rank;
            const int grp = m_tile / tiles_per_group, rem = m_tile - grp * tiles_per_group;
            const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
            const int per_img = p.Wt * p.Ht;
            const int nn = row / per_img, rr = row - nn * per_img;
            const int hh = rr / p.Wt, ww = rr - hh * p.Wt;
            const int img = grp * p.Nt + nn;
            const size_t pix = ((size_t)img * p.Ho + (th * p.Ht + hh)) * p.Wo + (tw * p.Wt + ww);
            const size_t pix0 = pix - lane;          // first pixel of this warp's 32 consecutive rows
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PAIR_N);
#pragma unroll 1
            for (int slab = slab0; slab < PAIR_N / 64; slab += SLAB_STEP) {
                const int col0 = n_tile * PAIR_N + slab * 64;
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(t_addr + (uint32_t)(slab * 64), r0);
                tmem_ld_32x32b_x32(t_addr + (uint32_t)(slab * 64 + 32), r1);
                uint4 res[8];
                if (has_res) {
                    // coalesced: each load instruction covers 4 pixel rows x 128 B; lane -> (row (lane >> 3) + 4k, 16-byte chunk lane & 7)
                    const T* rbase = reinterpret_cast<const T*>(p.epi.residual) + col0 + (lane & 7) * 8;
#pragma unroll
                    for (int k = 0; k < 8; ++k) res[k] = __ldg(reinterpret_cast<const uint4*>(rbase + (pix0 + (lane >> 3) + 4 * k) * p.epi.Cout));
                }
                uint8_t* buf = buf0 + sb * 4096;
                sb ^= 1;
                if (lane == 0) bulk_wait_read<1>();   // the store issued two slabs ago (same buffer) has finished reading it
                __syncwarp();
                if (has_res) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int i = (lane >> 3) + 4 * k;
                        *reinterpret_cast<uint4*>(buf + i * 128 + (((lane & 7) ^ (i & 7)) << 4)) = res[k];
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) res[c] = *reinterpret_cast<const uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4));
                    __syncwarp();
                }
                tmem_ld_wait();
                if (slab + SLAB_STEP >= PAIR_N / 64) {
                    // this warp's share of the accumulator is in registers: release it before the arithmetic / stores of its last slab
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(remote_tempty + (uint32_t)as * 8u);   // the leader's barrier (rank 0 maps to itself)
                }
                if (p.epi.stats_cw == 2) {
                    tc_epilogue_chunk32<T, true, 2>(p.epi, r0, col0, pix, img, lane, buf + lane * 128, 0, res, has_res);
                    tc_epilogue_chunk32<T, true, 2>(p.epi, r1, col0 + 32, pix, img, lane, buf + lane * 128, 4, res + 4, has_res);
                } else {
                    tc_epilogue_chunk32<T, true, 4>(p.epi, r0, col0, pix, img, lane, buf + lane * 128, 0, res, has_res);
                    tc_epilogue_chunk32<T, true, 4>(p.epi, r1, col0 + 32, pix, img, lane, buf + lane * 128, 4, res + 4, has_res);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&p.tmOut, buf, col0, (int)pix0);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait_read<0>();   // shared memory must outlive the last store's reads
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // both CTAs are done with the pair's TMEM and with each other's barriers
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool conv_pair_supported(const ConvTcDesc& d) {
    // opt-in (PHENDIFF_B200_LIN2CTA=1, read at plan time): measured on par with the halo kernel's 1x1 mode (out-proj -8 %, qkv +6 %,
    // profiles/r8b_pair_kernel.md) — these layers are bound on the output side (epilogue), not by the operand port
    const char* e = getenv("PHENDIFF_B200_LIN2CTA");
    if (!e || atoi(e) == 0) return false;
    if (d.ksize != 1 || d.stride != 1 || d.upsample || d.gn_coef || d.x2 || d.C2 || d.Csc1 || d.Csc2 || d.mode != TC_MODE_STD) return false;
    if (d.Cout % PAIR_N != 0 || d.C % (64 * PAIR_KSUB) != 0) return false;
    if (!conv_tc_supported(d, nullptr)) return false;
    int Wt, Ht, Nt;
    if (!tile_geometry(d.N, d.Ho, d.Wo, &Wt, &Ht, &Nt)) return false;
    const int m_tiles = (d.N / Nt) * (d.Wo / Wt) * (d.Ho / Ht);
    if (m_tiles % 2 != 0) return false;
    if (Wt != d.Wo || (Wt * Ht) % 32 != 0) return false;      // a warp's 32 accumulator rows = 32 consecutive pixels of one image (TMA-store slabs)
    if ((size_t)d.N * d.Ho * d.Wo > 0x7fffffffull) return false;
    if (d.stats_out && !conv_tc_can_emit_stats(d)) return false;
    return true;
}

static int pick_block_n(int Cout) {
    if (Cout % 256 == 0) return 256;
    if (Cout % 128 == 0) return 128;
    if (Cout % 64 == 0) return 64;
    return 0;
}

bool conv_tc_supported(const ConvTcDesc& d, std::string* why) {
    auto no = [&](const char* m) { if (why) *why = m; return false; };
    if (d.dt != DT_BF16 && d.dt != DT_F16) return no("tcgen05 path takes bf16 or fp16 activations");
    if (d.upsample || d.mode != TC_MODE_STD) return no("the per-tap kernel has no upsample / conv_out mode");
    if (d.stats_out && d.stats_cw != 4 && d.stats_cw != 2) return no("fused statistics need a chunk width of 4 or 2");
    if (d.C % 64 != 0 || d.Csc1 % 64 != 0 || d.Csc2 % 64 != 0) return no("channel counts must be multiples of 64");
    if (pick_block_n(d.Cout) == 0) return no("Cout must be a multiple of 64");
    if (!(d.ksize == 1 || d.ksize == 3)) return no("kernel size must be 1 or 3");
    if (d.stride == 2) {
        if (d.ksize != 3 || d.H % 2 || d.W % 2) return no("stride-2 needs a 3x3 kernel and even H, W");
        if (d.Ho != d.H / 2 || d.Wo != d.W / 2) return no("stride-2 output shape");
        if (d.pad != 0 && d.pad != 1) return no("stride-2 pad must be 0 or 1");
        if (d.Csc1 || d.Csc2) return no("no shortcut segment on stride-2 convs");
    } else if (d.stride == 1) {
        if (d.Ho != d.H || d.Wo != d.W || d.pad != d.ksize / 2) return no("stride-1 convs must be 'same'");
    } else return no("stride must be 1 or 2");
    int Wt, Ht, Nt;
    if (!tile_geometry(d.N, d.Ho, d.Wo, &Wt, &Ht, &Nt)) return no("output extent does not tile into 128-pixel boxes");
    return true;
}

int conv_tap_plan_create(const ConvTcDesc& d, ConvTapPlan** out) {
    std::string why;
    PD_REQUIRE(conv_tc_supported(d, &why), ("conv_tc: unsupported shape: " + why).c_str());
    ConvTapPlan* pl = new ConvTapPlan();
    ConvTcParams& p = pl->p;
    memset(&p, 0, sizeof(p));
    p.ksize = d.ksize; p.pad = d.pad; p.stride2 = d.stride == 2; p.C = d.C;
    p.kb_main = d.C / 64; p.kb_s1 = d.Csc1 / 64; p.kb_s2 = d.Csc2 / 64;
    p.num_kb = d.ksize * d.ksize * p.kb_main + p.kb_s1 + p.kb_s2;
    tile_geometry(d.N, d.Ho, d.Wo, &p.Wt, &p.Ht, &p.Nt);
    p.tilesW = d.Wo / p.Wt; p.tilesH = d.Ho / p.Ht;
    p.Ho = d.Ho; p.Wo = d.Wo; p.Cout = d.Cout;
    p.m_tiles = (d.N / p.Nt) * p.tilesW * p.tilesH;
    pl->dt = d.dt;
    pl->pair = conv_pair_supported(d);
    pl->block_n = pl->pair ? PAIR_N : pick_block_n(d.Cout);
    p.n_tiles = d.Cout / pl->block_n;
    TcEpi& e = p.epi;
    e.bias = d.bias; e.addvec = d.addvec; e.addvec_row = d.addvec_row; e.addvec_stride = d.addvec_stride;
    e.residual = d.residual; e.out_scale = d.out_scale; e.out = d.out; e.Cout = d.Cout;
    e.stats = d.stats_out; e.stats_cw = d.stats_cw;
    PD_REQUIRE(!d.stats_out || conv_tc_can_emit_stats(d), "conv_tc: fused statistics need 32-row groups inside one image");
    const uint64_t C = d.C, H = d.H, W = d.W, N = d.N;
    int rc = 0;
    if (!p.stride2) {
        uint64_t dims[4] = {C, W, H, N};
        uint64_t st[3] = {C * 2, W * C * 2, H * W * C * 2};
        uint32_t box[4] = {64, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Nt};
        rc = tc_encode_map(&p.tmA, d.dt, d.x, 4, dims, st, box);
    } else {
        uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, N};
        uint64_t st[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
        uint32_t box[5] = {64, (uint32_t)p.Wt, 1, (uint32_t)p.Ht, (uint32_t)p.Nt};
        rc = tc_encode_map(&p.tmA, d.dt, d.x, 5, dims, st, box);
    }
    if (rc) { delete pl; return rc; }
    const void* scs[2] = {d.sc1, d.sc2};
    const int cscs[2] = {d.Csc1, d.Csc2};
    CUtensorMap* tms[2] = {&p.tmS1, &p.tmS2};
    for (int i = 0; i < 2; ++i) {
        if (!cscs[i]) continue;
        uint64_t Cs = cscs[i], Ho = d.Ho, Wo = d.Wo;
        uint64_t dims[4] = {Cs, Wo, Ho, N};
        uint64_t st[3] = {Cs * 2, Wo * Cs * 2, Ho * Wo * Cs * 2};
        uint32_t box[4] = {64, (uint32_t)p.Wt, (uint32_t)p.Ht, (uint32_t)p.Nt};
        rc = tc_encode_map(tms[i], d.dt, scs[i], 4, dims, st, box);
        if (rc) { delete pl; return rc; }
    }
    {
        const uint64_t Ktot = (uint64_t)p.num_kb * 64;
        uint64_t dims[2] = {Ktot, (uint64_t)d.Cout};
        uint64_t st[1] = {Ktot * 2};
        uint32_t box[2] = {64, (uint32_t)pl->block_n};
        rc = tc_encode_map(&p.tmB, d.dt, d.wmat, 2, dims, st, box);
        if (rc) { delete pl; return rc; }
        if (pl->pair) {
            uint32_t boxh[2] = {64, (uint32_t)(PAIR_N / 2)};
            rc = tc_encode_map(&p.tmBh, d.dt, d.wmat, 2, dims, st, boxh);
            if (rc) { delete pl; return rc; }
            uint64_t od[2] = {(uint64_t)d.Cout, (uint64_t)d.N * d.Ho * d.Wo};
            uint64_t os[1] = {(uint64_t)d.Cout * 2};
            uint32_t ob[2] = {64, 32};
            rc = tc_encode_map(&p.tmOut, d.dt, d.out, 2, od, os, ob);
            if (rc) { delete pl; return rc; }
        }
    }
    pl->grid = std::min(p.m_tiles * p.n_tiles, tc_num_sms());
    pl->smem = pl->block_n == 256 ? TcCfg<256>::SMEM : (pl->block_n == 128 ? TcCfg<128>::SMEM : TcCfg<64>::SMEM);
    if (pl->pair) {
        pl->grid = std::min(p.m_tiles * p.n_tiles, tc_num_sms()) & ~1;     // whole CTA pairs
        pl->smem = PAIR_SMEM;
    }
    *out = pl;
    return 0;
}

void conv_tap_plan_destroy(ConvTapPlan* p) { delete p; }

template <int BLOCK_N, typename T>
static int launch_tc(const ConvTapPlan* pl, cudaStream_t s) {
    static bool attr_set_dev[PD_MAX_DEVICES] = {false};
    bool& attr_set = attr_set_dev[pd_cur_dev()];
    if (!attr_set) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)TcCfg<BLOCK_N>::SMEM));
        attr_set = true;
    }
    conv_tc_kernel<BLOCK_N, T><<<pl->grid, TC_THREADS, TcCfg<BLOCK_N>::SMEM, s>>>(pl->p);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static long long g_pair_launches = 0;
long long conv_pair_launch_count() { return g_pair_launches; }

template <typename T>
static int launch_pair(const ConvTapPlan* pl, cudaStream_t s) {
    static bool attr_set_dev[PD_MAX_DEVICES] = {false};
    bool& attr_set = attr_set_dev[pd_cur_dev()];
    if (!attr_set) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(conv_pair_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM));
        attr_set = true;
    }
    conv_pair_kernel<T><<<pl->grid, PAIR_THREADS, PAIR_SMEM, s>>>(pl->p);
    g_pair_launches++;
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int conv_tap_launch(const ConvTapPlan* pl, cudaStream_t s) {
    if (pl->pair) { PD_DISPATCH_HALF(pl->dt, T, { return launch_pair<T>(pl, s); }); }
    PD_DISPATCH_HALF(pl->dt, T, {
        switch (pl->block_n) {
            case 256: return launch_tc<256, T>(pl, s);
            case 128: return launch_tc<128, T>(pl, s);
            case 64: return launch_tc<64, T>(pl, s);
        }
    });
    set_error("conv_tc: bad block_n");
    return 1;
}

bool conv_tc_can_emit_stats(const ConvTcDesc& d) {
    int Wt, Ht, Nt;
    if (!tile_geometry(d.N, d.Ho, d.Wo, &Wt, &Ht, &Nt)) return false;
    return (Wt * Ht) % 32 == 0;   // a warp's 32 accumulator rows stay inside one image
}

int tc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// ---- dispatch between the two kernels ----------------------------------------------------------------------------------
int conv_tc_plan_create(const ConvTcDesc& d, ConvTcPlan** out) {
    ConvTcPlan* pl = new ConvTcPlan();
    int rc;
    if (conv_pair_supported(d)) {
        pl->kind = TC_KIND_TAP;
        rc = conv_tap_plan_create(d, &pl->tap);
    } else if (conv_halo_supported(d, nullptr)) {
        pl->kind = TC_KIND_HALO;
        rc = conv_halo_plan_create(d, &pl->halo);
    } else {
        pl->kind = TC_KIND_TAP;
        rc = conv_tap_plan_create(d, &pl->tap);
    }
    if (rc) { delete pl; return rc; }
    *out = pl;
    return 0;
}
void conv_tc_plan_destroy(ConvTcPlan* p) {
    if (!p) return;
    if (p->tap) conv_tap_plan_destroy(p->tap);
    if (p->halo) conv_halo_plan_destroy(p->halo);
    delete p;
}
int conv_tc_launch(const ConvTcPlan* p, cudaStream_t s, const ConvTcLaunch* extra) {
    return p->kind == TC_KIND_HALO ? conv_halo_launch(p->halo, s, extra) : conv_tap_launch(p->tap, s);
}

}  // namespace pd
