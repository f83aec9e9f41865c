// phendiff_b200 — implicit-GEMM convolution on tcgen05 + TMEM with HALO tiles: the activation tile is loaded from L2
// ONCE per 64-channel block and reused by all filter taps.
//
// Why: the per-tap kernel (pd_conv_tc.cu) re-loads a shifted copy of the same activation box for each of the 9 taps, so
// every CTA pulls (128 + BLOCK_N) x 128 B from L2 per 64-deep K block: 32..43 MAC/B.  ncu (profiles/r1a_ncu_conv_tc.md)
// shows both of its variants pinned at ~12 TB/s of L2->SM traffic with the tensor pipe 40 % active: the path is bounded
// by L2 delivery, not by the MMA rate.  Here one TMA box {64 ch, Wt+2, Ht+2} (a 16x8-pixel output tile plus its halo,
// 23 KB) lands in shared memory with the 128-byte swizzle, and tap (r,s) is the SAME buffer addressed from a start that
// is shifted by (r*(Wt+2)+s) pixel rows: 8 consecutive pixels of one image row form one 8x128 B swizzle atom, and output
// row h+1 is exactly (Wt+2)*128 B further, i.e. a uniform stride-between-8-row-groups (SBO = 1280 B) in the UMMA descriptor.
// Measured on B200 (tools/gpu_halo_probe.sh, profiles/r1b_halo_probe.md): the tensor core applies the 128-byte swizzle to
// the ABSOLUTE shared-memory address bits (like TMA does when writing), so a start shifted by whole 128-byte rows and an
// SBO that is not a multiple of 1024 both address the right data with the descriptor's base-offset field left at 0
// (setting base offset = row phase breaks it).  A-traffic drops 6.2x.
//
// GEMM view: M = N*Ho*Wo pixels (tiles of Ht=16 x Wt=8), Ncol = Cout, K = taps*C (+ shortcut C).
//   segments of K: [main: taps x C] [1x1 shortcut over source 1] [1x1 shortcut over source 2]   (conv2 + conv_shortcut
//   of a ResnetBlock2D run as ONE GEMM; the concat of the up blocks is never materialised for the shortcut)
//   upsample mode: nearest-2x + 3x3 conv (Upsample2D) = four 2x2 sub-pixel phase convs on the low-res input (2.25x
//   fewer MACs, no 4x tensor): phase (a,b) reads rows {i-1+a, i+a}, cols {j-1+b, j+b} and writes pixel (2i+a, 2j+b).
//   TC_MODE_DDIM: conv_out (Cout = 3 padded to 16); the epilogue applies the scheduler update to x_t in place.
// Warp roles (384 threads, 1 CTA/SM, persistent): warp 0 TMA producer of weight tiles, warp 3 TMA producer of halo tiles,
// warp 1 MMA issuer (one thread), warp 2 TMEM allocator, warps 4-11 epilogue (two per TMEM lane quarter; TMA-store of swizzled 64-channel slabs).  Rings: A (halo tiles, SA stages) and B (weight tiles, SB stages) are decoupled: one A
// stage lives through `taps` B stages.  The fp32 accumulator is double buffered in TMEM.
//
// GN variant (template flag, 512 threads): the conv's input is GroupNorm(+SiLU) of concat(x, x2) and the normalised tensor
// is NEVER written to HBM.  The producer loads the RAW halo tile; four extra warps (12-15) apply y = silu(x * a[n,c] + b[n,c])
// to it in place in shared memory (per-image per-channel coefficients from gn_coef_kernel; halo pixels outside the image
// stay zero = the conv's zero padding of the NORMALISED tensor), fence the generic-proxy writes towards the async proxy
// and hand the stage to the MMA issuer through a third barrier ring (readyA).  The transform of stage k+1 runs under the
// 36 MMAs of stage k.  Registers are re-split with setmaxnreg (control 96, epilogue 168, transform 80 per thread).
#include "pd_tc_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

namespace pd {

struct HaloParams {
    CUtensorMap tmA, tmA2, tmS1, tmS2, tmB;   // tmA2: second main-segment source (channel concat), GN variant only
    CUtensorMap tmOut;             // output: 4-D {C, Wo, Ho, N} box {64, 8, 4, 1}; upsample: 5-D {C, 2, W, 2, N*H} box {64, 1, 8, 1, 4}
    int ntaps, kw, pitch_px;
    int a_bytes_main, a_bytes_sc, a_stage_bytes, SA, SB;
    int kb_main, kb_s1, kb_s2, C;
    int tilesW, tilesH;
    int off_h, off_w;              // box origin relative to the tile origin (-pad); upsample adds the phase
    int m_tiles, n_tiles, phases, b_rows_per_phase;
    int Ho, Wo;                    // output extent (2x the tile-space extent in upsample mode)
    int upsample;
    int mt;                        // M halves per tile: 1 = 16x8 pixels, 2 = 16x16 pixels (two accumulators share every weight tile)
    int kb_a1;                     // 64-channel blocks of the main segment that come from tmA (the rest from tmA2)
    int sc_per_stage;              // 1x1-shortcut K blocks that share one A stage (1 or 2): a shortcut block is ONE tap (~0.5 k cycles
                                   // of MMAs), shorter than a TMA round trip, so two of them travel per stage where shared memory allows
    const float2* gn_coef;         // GN variant: (N, C) (scale, shift) of the fused GroupNorm, SiLU follows; else null
    int H, W;                      // input extent (halo mask of the GN variant)
    int gn_tanh;                   // SiLU of the GN variant through tanh.approx (1 MUFU / element) instead of ex2 + rcp
    int tile_begin, tile_end;      // tile range of this launch (default: 0 .. m_tiles * phases * n_tiles)
    TcEpi epi;
};

struct ConvHaloPlan {
    HaloParams p;
    int dt, block_n, grid, mode, gn;
    size_t smem;
};

constexpr int HL_WT = 8, HL_HT = 16;
constexpr int HL_EPI_BYTES = 8 * 4096;   // one 4 KB output slab per epilogue warp

// epilogue warps that have work: two per TMEM lane quarter when the tile has >= 2 slabs of 64 output channels
template <int BLOCK_N, int MODE> struct HaloEpiWarps { static constexpr int value = (MODE == TC_MODE_STD && BLOCK_N >= 128) ? 8 : 4; };

template <int R> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// y = silu(x * a + b) on 8 packed 16-bit values.  TANH = false: cf = (a0,b0,a1,b1) pairs, silu(y) = y / (1 + 2^(-y log2 e))
// (MUFU.EX2 + MUFU.RCP per element).  TANH = true: cf holds (a/2, b/2), silu(y) = h + h tanh(h) with h = y/2: ONE MUFU per
// element; tanh.approx.f32 is good to 2^-11 relative, i.e. the same size as the fp16 rounding of the stored activation.
template <typename T, bool TANH> __device__ __forceinline__ uint4 gn_silu8(const uint4& raw, const float4 (&cf)[4]) {
    float v[8];
    unpack8<T>(raw, v);
    const float a[8] = {cf[0].x, cf[0].z, cf[1].x, cf[1].z, cf[2].x, cf[2].z, cf[3].x, cf[3].z};
    const float b[8] = {cf[0].y, cf[0].w, cf[1].y, cf[1].w, cf[2].y, cf[2].w, cf[3].y, cf[3].w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float y = fmaf(v[i], a[i], b[i]);
        if (TANH) {
            float t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(y));
            v[i] = fmaf(y, t, y);
        } else {
            v[i] = silu<false>(y);
        }
    }
    uint4 o;
    o.x = pack2<T>(v[0], v[1]); o.y = pack2<T>(v[2], v[3]); o.z = pack2<T>(v[4], v[5]); o.w = pack2<T>(v[6], v[7]);
    return o;
}

// GroupNorm + SiLU of one raw halo tile in place (128 threads; see the GN-variant note in the file header).  Thread ->
// 16-byte chunk column `col` of rows r0, r0+16, ...; two rows (16 values) are in flight per iteration so that the MUFU
// latency of one hides under the other.  Rows outside the image keep their TMA zero fill: they are the conv's padding.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T, bool TANH>
__device__ __forceinline__ void gn_transform_tile(uint32_t col, int r0, int nrows, int pitch, int qstep, int rstep, int hy, int hx,
                                                  int gy0, int gx0, int H, int W, const float4 (&cf)[4]) {
#pragma unroll 1
    for (int rho = r0; rho < nrows; rho += 32) {
        const bool ok0 = (unsigned)(gy0 + hy) < (unsigned)H && (unsigned)(gx0 + hx) < (unsigned)W;
        hy += qstep; hx += rstep;
        if (hx >= pitch) { hx -= pitch; ++hy; }
        const bool in1 = rho + 16 < nrows;
        const bool ok1 = in1 && (unsigned)(gy0 + hy) < (unsigned)H && (unsigned)(gx0 + hx) < (unsigned)W;
        hy += qstep; hx += rstep;
        if (hx >= pitch) { hx -= pitch; ++hy; }
        const uint32_t q0 = col + (uint32_t)rho * 128u, q1 = q0 + 2048u;
        const uint4 v0 = lds128(q0);
        const uint4 v1 = in1 ? lds128(q1) : make_uint4(0u, 0u, 0u, 0u);
        const uint4 o0 = gn_silu8<T, TANH>(v0, cf);
        const uint4 o1 = gn_silu8<T, TANH>(v1, cf);
        if (ok0) sts128(q0, o0);
        if (ok1) sts128(q1, o1);
    }
}

template <int BLOCK_N, typename T, int MODE, int CW, int MT, bool GN>
__global__ void __launch_bounds__(GN ? HALO_THREADS_GN : HALO_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
    constexpr int EPI_WARPS = HaloEpiWarps<BLOCK_N, MODE>::value;
    constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
    constexpr int TMEM_COLS = (2 * MT * BLOCK_N < 32) ? 32 : 2 * MT * BLOCK_N;
    static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
    constexpr int TILE_W = HL_WT * MT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int SA = p.SA, SB = p.SB;
    uint8_t* smA = smem;
    uint8_t* smB = smem + (size_t)SA * p.a_stage_bytes;
    uint8_t* smEpi = smB + (size_t)SB * B_BYTES;   // 8 epilogue warps x 4 KB output slab (TC_MODE_STD)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smEpi + (MODE == TC_MODE_STD ? HL_EPI_BYTES : 0));
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + SA;
    uint64_t* fullB = emptyA + SA;
    uint64_t* emptyB = fullB + SB;
    uint64_t* tfull = emptyB + SB;
    uint64_t* tempty = tfull + 2;
    uint64_t* readyA = tempty + 2;   // GN variant: stage transformed (or passed through) by warps 12-15
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(readyA + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&p.tmA);
        prefetch_tmap(&p.tmB);
        if (p.kb_a1 < p.kb_main) prefetch_tmap(&p.tmA2);
        if (p.kb_s1) prefetch_tmap(&p.tmS1);
        if (p.kb_s2) prefetch_tmap(&p.tmS2);
        if (MODE == TC_MODE_STD) prefetch_tmap(&p.tmOut);
    }
    if (warp == 1 && elect_one()) {
        for (int i = 0; i < SA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < SB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * EPI_WARPS); }
        if (GN) for (int i = 0; i < SA; ++i) mbar_init(&readyA[i], 128);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int total_tiles = p.tile_end;   // [tile_begin, tile_end): all tiles, or the image range of a conv_out launch
    const int tiles_per_img = p.tilesW * p.tilesH;

    // GN variant: 512 threads start at 128 registers each; every role branch opens with the setmaxnreg of its warpgroup
    // (control 96, transform 80, epilogue 168) so that ptxas allocates each region against its own budget
    if (warp < 4) {
      if (GN) setmaxnreg_dec<96>();
      if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer, weight tiles (B ring) =====================
            int sb = 0;
            uint32_t phb = 0;
            for (int tile = p.tile_begin + blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int phase = (tile / p.n_tiles) % p.phases;
                const int brow = phase * p.b_rows_per_phase + n_tile * BLOCK_N;
                int kcol = 0;
                for (int seg = 0; seg < 3; ++seg) {
                    const int nkb = seg == 0 ? p.kb_main : (seg == 1 ? p.kb_s1 : p.kb_s2);
                    const int ntap = seg == 0 ? p.ntaps : 1;
                    const int segC = nkb * TC_BLOCK_K;
                    for (int cb = 0; cb < nkb; ++cb) {
                        for (int tap = 0; tap < ntap; ++tap) {
                            mbar_wait(&emptyB[sb], phb ^ 1);
                            mbar_arrive_expect_tx(&fullB[sb], B_BYTES);
                            tma_load_2d(&p.tmB, &fullB[sb], smB + (size_t)sb * B_BYTES, kcol + tap * segC + cb * TC_BLOCK_K, brow);
                            if (++sb == SB) { sb = 0; phb ^= 1; }
                        }
                    }
                    kcol += ntap * segC;
                }
            }
        }
      } else if (warp == 3) {
        if (elect_one()) {
            // ===================== TMA producer, activation halo tiles (A ring) =====================
            // Its own thread: an A stage is requested the moment the MMAs that read its previous content retire, not when
            // the weight producer gets round to it (with one producer the request went out only SB taps before the
            // tile was needed, which left no room for the GN variant's in-place transform).
            int sa = 0;
            uint32_t pha = 0;
            for (int tile = p.tile_begin + blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int t2 = tile / p.n_tiles;
                const int phase = t2 % p.phases, m_tile = t2 / p.phases;
                const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
                const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
                const int h0 = th * HL_HT, w0 = tw * TILE_W;
                const int oh = p.off_h + (p.upsample ? (phase >> 1) : 0), ow = p.off_w + (p.upsample ? (phase & 1) : 0);
                for (int seg = 0; seg < 3; ++seg) {
                    const int nkb = seg == 0 ? p.kb_main : (seg == 1 ? p.kb_s1 : p.kb_s2);
                    const int per = seg == 0 ? 1 : p.sc_per_stage;   // K blocks per stage
                    for (int cb = 0; cb < nkb; cb += per) {
                        mbar_wait(&emptyA[sa], pha ^ 1);
                        uint8_t* dstA = smA + (size_t)sa * p.a_stage_bytes;
                        if (seg == 0) {
                            mbar_arrive_expect_tx(&fullA[sa], p.a_bytes_main);
                            if (cb < p.kb_a1) tma_load_4d(&p.tmA, &fullA[sa], dstA, cb * TC_BLOCK_K, w0 + ow, h0 + oh, img);
                            else tma_load_4d(&p.tmA2, &fullA[sa], dstA, (cb - p.kb_a1) * TC_BLOCK_K, w0 + ow, h0 + oh, img);
                        } else {
                            const int nb = min(per, nkb - cb);
                            mbar_arrive_expect_tx(&fullA[sa], nb * p.a_bytes_sc);
                            for (int b = 0; b < nb; ++b)
                                tma_load_4d(seg == 1 ? &p.tmS1 : &p.tmS2, &fullA[sa], dstA + (size_t)b * p.a_bytes_sc, (cb + b) * TC_BLOCK_K, w0, h0, img);
                        }
                        if (++sa == SA) { sa = 0; pha ^= 1; }
                    }
                }
            }
        }
      } else if (warp == 1) {
        if (elect_one()) {
            // ===================== MMA issuer (single thread) =====================
            // Everything a tap needs is an add away: ncu on conv_out (profiles/r1y_ncu_conv_out.md) showed this thread — not a
            // barrier — as the critical path of the kernel at ~510 cycles of dependent scalar work per tap (an integer division
            // by the runtime kernel width, 64-bit descriptor assembly, generic->shared conversions and constant-bank reloads),
            // i.e. as long as the four N = 256 MMAs of a tap keep the tensor pipe busy.  Descriptors are now (constant high
            // word, running low word = address >> 4), barrier addresses are 32-bit shared addresses advanced by 8.
            constexpr uint32_t idesc = make_idesc<T, BLOCK_N>();
            constexpr uint32_t B_HI = sw128_desc_hi(1024u);
            constexpr uint32_t B_LO_STEP = (uint32_t)B_BYTES >> 4;
            const int kw = p.kw, ntaps = p.ntaps, pitch_main = p.pitch_px;
            const int nkbs[3] = {p.kb_main, p.kb_s1, p.kb_s2};
            const int sc_per_stage = p.sc_per_stage;
            const uint32_t sc_blk_lo = (uint32_t)p.a_bytes_sc >> 4;
            const uint32_t a_lo0 = (smem_u32(smA) & 0x3FFFFu) >> 4, a_lo_step = (uint32_t)p.a_stage_bytes >> 4;
            const uint32_t b_lo0 = (smem_u32(smB) & 0x3FFFFu) >> 4;
            const uint32_t bar_fullA = smem_u32(GN ? readyA : fullA), bar_emptyA = smem_u32(emptyA);
            const uint32_t bar_fullB = smem_u32(fullB), bar_emptyB = smem_u32(emptyB);
            const uint32_t bar_tfull = smem_u32(tfull), bar_tempty = smem_u32(tempty);
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            int iter = 0;
            for (int tile = p.tile_begin + blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
                const int as = iter & 1;
                const uint32_t aphase = (iter >> 1) & 1;
                mbar_wait_addr(bar_tempty + as * 8, aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * MT * BLOCK_N);
                uint32_t accum = 0;
#pragma unroll 1
                for (int seg = 0; seg < 3; ++seg) {
                    const int nkb = nkbs[seg];
                    const int pitch = seg == 0 ? pitch_main : TILE_W;
                    const uint32_t a_hi = sw128_desc_hi((uint32_t)pitch * 128u);   // SBO = one pixel row of the tile per 8-row group step
                    // a shortcut stage holds up to sc_per_stage K blocks of ONE tap each: it runs through the same loop as the taps
                    // of a main block, stepping a whole block (a_bytes_sc) instead of one pixel row and never wrapping
                    const int per = seg == 0 ? 1 : sc_per_stage;
                    const uint32_t tap_step = seg == 0 ? 8u : sc_blk_lo;
                    const int wrap_at = seg == 0 ? kw : (1 << 30);
                    const uint32_t row_wrap = (uint32_t)(pitch - kw) * 8u;
#pragma unroll 1
                    for (int cb = 0; cb < nkb; cb += per) {
                        const int ntap = seg == 0 ? ntaps : min(per, nkb - cb);
                        mbar_wait_addr(bar_fullA + sa * 8, pha);
                        tc_fence_after();
                        uint32_t a_lo = a_lo0 + (uint32_t)sa * a_lo_step;   // + 8 per pixel row of 128 B
                        int sx = 0;
#pragma unroll 1
                        for (int tap = 0; tap < ntap; ++tap) {
                            mbar_wait_addr(bar_fullB + sb * 8, phb);
                            tc_fence_after();
                            const uint32_t b_lo = b_lo0 + (uint32_t)sb * B_LO_STEP;
                            // K slice outer, tile half inner: consecutive MMAs alternate between the two accumulators instead of
                            // chaining four dependent accumulations into the same TMEM columns
#pragma unroll
                            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
#pragma unroll
                                for (int half = 0; half < MT; ++half)
                                    // the second 16x8-pixel half starts 8 pixel rows (one 1024-byte swizzle atom) further in the same halo tile
                                    umma_f16kind(d_tmem + (uint32_t)(half * BLOCK_N), desc64(a_hi, a_lo + (uint32_t)(half * HL_WT * 8 + 2 * k)),
                                                 desc64(B_HI, b_lo + (uint32_t)(2 * k)), idesc, (accum | (uint32_t)k) ? 1u : 0u);
                            }
                            accum = 1;
                            umma_commit_addr(bar_emptyB + sb * 8);
                            if (++sb == SB) { sb = 0; phb ^= 1; }
                            a_lo += tap_step;                                  // next tap: one pixel to the right ...
                            if (++sx == wrap_at) { sx = 0; a_lo += row_wrap; }   // ... or the first pixel of the next halo row
                        }
                        umma_commit_addr(bar_emptyA + sa * 8);
                        if (++sa == SA) { sa = 0; pha ^= 1; }
                    }
                }
                umma_commit_addr(bar_tfull + as * 8);
            }
        }
      }
    } else if (warp < 12) {
      if (GN) setmaxnreg_inc<168>();
      if (warp < 4 + EPI_WARPS) {
        // ===================== epilogue (8 warps: two per TMEM lane quarter, alternating 64-channel slabs) =====================
        // Accumulator row = output pixel.  A warp owns 32 rows = a 4 x 8-pixel box of the tile; it stages 64 output channels
        // at a time as a SWIZZLE_128B slab (32 x 128 B) in its private shared-memory buffer and one elected lane writes it
        // with a single TMA store (fully coalesced; per-lane stores would touch 32 separate lines per instruction).  A
        // residual tile is read coalesced and transposed through the same buffer.  Two warps per scheduler hide each
        // other's TMEM-load, shuffle and global-load latencies.
        const int ew = warp - 4;
        const int q = ew & 3, slab0 = ew >> 2;
        constexpr int SLAB_STEP = EPI_WARPS / 4;
        const int row = q * 32 + lane;
        const int hh = row >> 3, ww = row & 7;
        uint8_t* buf = smEpi + ew * 4096;
        int iter = 0;
        for (int tile = p.tile_begin + blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int n_tile = tile % p.n_tiles;
            const int t2 = tile / p.n_tiles;
            const int phase = t2 % p.phases, m_tile = t2 / p.phases;
            const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
            const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
            const int h0 = th * HL_HT, w0 = tw * TILE_W;
            const int as = iter & 1;
            const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * MT * BLOCK_N);
            if (MODE == TC_MODE_DDIM) {
#pragma unroll
                for (int half = 0; half < MT; ++half) {
                    const int oh = h0 + hh, ow = w0 + half * HL_WT + ww;
                    const int hw = oh * p.Wo + ow;
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(t_addr + (uint32_t)(half * BLOCK_N), r);
                    tmem_ld_wait();
                    if (half == MT - 1) {
                        tc_fence_before();
                        mbar_arrive(&tempty[as]);      // accumulators are in registers: the MMA warp may overwrite this TMEM buffer
                    }
                    tc_epilogue_ddim(p.epi, r, img, hw);
                }
            } else {
                constexpr int SLABS = BLOCK_N >= 64 ? BLOCK_N / 64 : 1, ITEMS = MT * SLABS;
#pragma unroll 1
                for (int item = slab0; item < ITEMS; item += SLAB_STEP) {
                    const int half = item / SLABS, slab = item - half * SLABS;
                    const int wh = w0 + half * HL_WT;             // first output column of this 16x8-pixel half
                    int oh = h0 + hh, ow = wh + ww;
                    if (p.upsample) { oh = 2 * oh + (phase >> 1); ow = 2 * ow + (phase & 1); }
                    const size_t pix = (size_t)img * p.Ho * p.Wo + (size_t)oh * p.Wo + ow;
                    const int col0 = n_tile * BLOCK_N + slab * 64;
                    const uint32_t t_item = t_addr + (uint32_t)(half * BLOCK_N + slab * 64);
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(t_item, r0);
                    tmem_ld_32x32b_x32(t_item + 32u, r1);
                    uint4 res[8];
                    const bool has_res = p.epi.residual != nullptr;
                    if (has_res) {
                        // coalesced: each load instruction covers 4 pixel rows x 128 B; lane -> (row (lane>>3)+4k, 16-byte chunk lane&7)
                        const T* rbase = reinterpret_cast<const T*>(p.epi.residual) + col0 + (lane & 7) * 8;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int i = (lane >> 3) + 4 * k;
                            const size_t rp = (size_t)img * p.Ho * p.Wo + (size_t)(h0 + 4 * q + (i >> 3)) * p.Wo + (wh + (i & 7));
                            res[k] = __ldg(reinterpret_cast<const uint4*>(rbase + rp * p.epi.Cout));
                        }
                    }
                    if (lane == 0) bulk_wait_read<0>();   // this warp's previous TMA store has finished reading the buffer
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
                    if (item + SLAB_STEP >= ITEMS) {   // last item of this warp: its share of the accumulators is in registers
                        tc_fence_before();
                        mbar_arrive(&tempty[as]);
                    }
                    tc_epilogue_chunk32<T, true, CW>(p.epi, r0, col0, pix, img, lane, buf + lane * 128, 0, res, has_res);
                    tc_epilogue_chunk32<T, true, CW>(p.epi, r1, col0 + 32, pix, img, lane, buf + lane * 128, 4, res + 4, has_res);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.upsample) tma_store_5d(&p.tmOut, buf, col0, phase & 1, wh, phase >> 1, img * (p.Ho >> 1) + h0 + 4 * q);
                        else tma_store_4d(&p.tmOut, buf, col0, wh, h0 + 4 * q, img);
                        bulk_commit();
                    }
                }
            }
        }
        if (MODE == TC_MODE_STD && lane == 0) bulk_wait_read<0>();   // shared memory must outlive the last store's reads
      }
    } else if (GN) {
        setmaxnreg_dec<80>();
        // ===================== GroupNorm + SiLU on the raw halo tile, in place (GN variant) =====================
        // 128 threads; thread -> 16-byte chunk j of rows r0, r0+16, ...  Rows are 128 B and stage bases 1024-aligned, so the
        // swizzle phase (row & 7) of a thread's rows is constant and so is its logical chunk = 8 channels: the coefficients
        // are loaded once per stage.
        const int t = threadIdx.x - 384;
        const int j = t & 7, r0 = t >> 3;
        const int lc = j ^ (r0 & 7);
        const int pitch = p.pitch_px;
        const int nrows = pitch * (HL_HT + 2);
        const int qstep = 16 / pitch, rstep = 16 - qstep * pitch;   // row + 16 -> (hy + qstep, hx + rstep) before the carry
        int hy0 = 0, hx0 = r0;
        while (hx0 >= pitch) { hx0 -= pitch; ++hy0; }
        int sa = 0;
        uint32_t pha = 0;
        // the coefficient table is written by the kernel this grid may have overtaken (programmatic dependent launch, see launch_halo)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int tile = p.tile_begin + blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int t2 = tile / p.n_tiles;
            const int m_tile = t2 / p.phases;
            const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
            const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
            const int gy0 = th * HL_HT + p.off_h, gx0 = tw * TILE_W + p.off_w;
            const float2* crow = p.gn_coef + (size_t)img * p.C + lc * 8;
            for (int seg = 0; seg < 3; ++seg) {
                const int nkb = seg == 0 ? p.kb_main : (seg == 1 ? p.kb_s1 : p.kb_s2);
                const int per = seg == 0 ? 1 : p.sc_per_stage;
                for (int cb = 0; cb < nkb; cb += per) {
                    float4 cf[4];
                    if (seg == 0) {
                        const float4* cp = reinterpret_cast<const float4*>(crow + cb * TC_BLOCK_K);
#pragma unroll
                        for (int i = 0; i < 4; ++i) cf[i] = __ldg(cp + i);
                        if (p.gn_tanh) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) { cf[i].x *= 0.5f; cf[i].y *= 0.5f; cf[i].z *= 0.5f; cf[i].w *= 0.5f; }
                        }
                    }
                    mbar_wait(&fullA[sa], pha);
                    if (seg == 0) {
                        const uint32_t col = smem_u32(smA + (size_t)sa * p.a_stage_bytes + j * 16);
                        if (p.gn_tanh) gn_transform_tile<T, true>(col, r0, nrows, pitch, qstep, rstep, hy0, hx0, gy0, gx0, p.H, p.W, cf);
                        else gn_transform_tile<T, false>(col, r0, nrows, pitch, qstep, rstep, hy0, hx0, gy0, gx0, p.H, p.W, cf);
                        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    }
                    mbar_arrive(&readyA[sa]);
                    if (++sa == SA) { sa = 0; pha ^= 1; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int halo_block_n(const ConvTcDesc& d) {
    if (d.mode == TC_MODE_DDIM) return d.Cout <= 16 ? 16 : 0;
    // 1x1 / linear layers have K = C only: a 128x256 tile re-streams 384 KB from L2 per 33 MFLOP (the qkv projection sat at
    // 660 TFLOP/s on the crossbar limit); 256 pixels x 128 channels on the dual-accumulator tile moves the same bytes for 2x the work
    static const int lin128 = [] { const char* e = getenv("PHENDIFF_B200_HALO_LIN128"); return e ? atoi(e) : 1; }();
    if (lin128 && d.ksize == 1 && !d.upsample && d.Cout % 128 == 0 && d.W % (2 * HL_WT) == 0) return 128;
    if (d.Cout % 256 == 0) return 256;
    if (d.Cout % 128 == 0) return 128;
    if (d.Cout % 64 == 0) return 64;
    return 0;
}

bool conv_halo_supported(const ConvTcDesc& d, std::string* why) {
    auto no = [&](const char* m) { if (why) *why = m; return false; };
    if (const char* off = getenv("PHENDIFF_B200_HALO")) if (off[0] == '0') return no("halo kernel disabled by PHENDIFF_B200_HALO=0");
    if (d.dt != DT_BF16 && d.dt != DT_F16) return no("tcgen05 path takes bf16 or fp16 activations");
    if (d.C % 64 != 0 || d.C2 % 64 != 0 || d.Csc1 % 64 != 0 || d.Csc2 % 64 != 0) return no("channel counts must be multiples of 64");
    if (d.C2 < 0 || d.C2 >= d.C + (d.C == 0)) return no("second main-segment source must leave channels for the first");
    if (d.gn_coef && (d.ksize != 3 || d.upsample)) return no("fused GroupNorm: plain 3x3 convolution only");
    if (d.gn_coef && d.mode == TC_MODE_STD && halo_block_n(d) < 128) return no("fused GroupNorm: Cout must be a multiple of 128");
    if (halo_block_n(d) == 0) return no("Cout must be a multiple of 64 (or <= 16 for the conv_out mode)");
    if (d.stride != 1) return no("halo kernel takes stride-1 convolutions");
    if (!(d.ksize == 1 || d.ksize == 3)) return no("kernel size must be 1 or 3");
    if (d.pad != d.ksize / 2) return no("convolution must be 'same'");
    if (d.upsample) {
        if (d.ksize != 3 || d.Ho != 2 * d.H || d.Wo != 2 * d.W || d.Csc1 || d.Csc2) return no("upsample mode: 3x3, 2x output, no shortcut");
    } else if (d.Ho != d.H || d.Wo != d.W) return no("output extent must equal the input extent");
    if (d.W % HL_WT != 0 || d.H % HL_HT != 0) return no("extent must tile into 16x8-pixel boxes");
    if (d.mode == TC_MODE_DDIM && (d.upsample || d.Csc1 || d.residual || d.addvec || d.stats_out)) return no("conv_out mode takes a plain 3x3 conv");
    if (d.stats_out && d.stats_cw != 4 && d.stats_cw != 2) return no("fused statistics need a chunk width of 4 or 2");
    return true;
}

int conv_halo_plan_create(const ConvTcDesc& d, ConvHaloPlan** out) {
    std::string why;
    PD_REQUIRE(conv_halo_supported(d, &why), ("conv_halo: unsupported shape: " + why).c_str());
    ConvHaloPlan* pl = new ConvHaloPlan();
    HaloParams& p = pl->p;
    memset(&p, 0, sizeof(p));
    pl->dt = d.dt; pl->mode = d.mode; pl->block_n = halo_block_n(d); pl->gn = d.gn_coef != nullptr;
    p.gn_coef = d.gn_coef; p.H = d.H; p.W = d.W;
    {   // PHENDIFF_B200_GN_SILU=exp|tanh: which SiLU the GN variant's transform warps evaluate.  Default tanh: one MUFU per
        // element instead of two, +2.5 % on the whole path, and the same end-to-end parity (fp16 per-step eps error 2.0e-3 vs
        // 2.0e-3, DDIB PSNR 74.1 dB both; profiles/r1w_gn_fusion.md)
        const char* e = getenv("PHENDIFF_B200_GN_SILU");
        p.gn_tanh = e ? (std::string(e) == "tanh") : 1;
    }
    const int kh = d.upsample ? 2 : d.ksize, kw = kh;
    p.ntaps = kh * kw; p.kw = kw;
    // probe knobs for the descriptor semantics (see file header; defaults = the measured-correct variant):
    // PHENDIFF_B200_HALO_PITCH=pow2 pads the halo row to 16 pixels (SBO 2048).  (The r1b probe also had a knob that set the
    // descriptor's base-offset field to the row phase of the shifted start; it addressed the wrong rows and was removed.)
    const char* pk = getenv("PHENDIFF_B200_HALO_PITCH");
    const bool pow2 = pk && std::string(pk) == "pow2";
    // M halves per tile: layers with <= 128 output channels per tile are bound by the L2 -> SM stream of the weight tiles
    // (profiles/r1k_ncu_conv_halo128.md: 7.8 TB/s of crossbar reads, tensor pipe 47 % active): a 16x16-pixel tile feeds two
    // accumulators from every weight tile and halves that stream.  TMEM holds 2 (double buffer) x mt x block_n fp32 columns.
    int mt = 1;
    if (pl->block_n <= 128 && d.W % (2 * HL_WT) == 0 && !pow2) mt = 2;   // conv_out (N = 16) too: its cost is the per-tap issue overhead
    if (const char* e = getenv("PHENDIFF_B200_HALO_MT")) mt = (atoi(e) == 2 && mt == 2) ? 2 : 1;
    p.mt = mt;
    const int tile_w = HL_WT * mt;
    p.pitch_px = (kw == 1) ? tile_w : (pow2 ? 16 : tile_w + kw - 1);
    const int rows = HL_HT + kh - 1;
    p.a_bytes_main = 128 * p.pitch_px * rows;
    p.a_bytes_sc = 128 * tile_w * HL_HT;
    p.sc_per_stage = 1;
    p.a_stage_bytes = ((std::max(p.a_bytes_main, p.a_bytes_sc) + 1023) / 1024) * 1024;
    p.C = d.C; p.kb_main = d.C / 64; p.kb_s1 = d.Csc1 / 64; p.kb_s2 = d.Csc2 / 64;
    p.kb_a1 = (d.C - d.C2) / 64;
    p.tilesW = d.W / tile_w; p.tilesH = d.H / HL_HT;
    p.off_h = p.off_w = d.upsample ? -1 : -d.pad;
    p.m_tiles = d.N * p.tilesW * p.tilesH;
    p.phases = d.upsample ? 4 : 1;
    const int cout_rows = d.mode == TC_MODE_DDIM ? 16 : d.Cout;
    p.b_rows_per_phase = cout_rows;
    p.n_tiles = cout_rows / pl->block_n;
    p.Ho = d.Ho; p.Wo = d.Wo; p.upsample = d.upsample;
    TcEpi& e = p.epi;
    e.bias = d.bias; e.addvec = d.addvec; e.addvec_row = d.addvec_row; e.addvec_stride = d.addvec_stride;
    e.residual = d.residual; e.out_scale = d.out_scale; e.out = d.out; e.stats = d.stats_out; e.stats_cw = d.stats_cw;
    e.Cout = d.Cout; e.c_valid = d.Cout; e.plane = d.Ho * d.Wo;
    // shared memory budget: B ring as deep as fits beside SA halo stages
    const int b_bytes = pl->block_n * 128;
    const int epi_bytes = d.mode == TC_MODE_STD ? HL_EPI_BYTES : 0;
    const int budget = 227 * 1024 - 1024 - 512 - epi_bytes;
    p.SA = (pl->block_n >= 256 || mt == 2) ? 2 : 3;
    if (pl->block_n == 16) p.SA = 4;
    // 1x1 / linear layers: an A stage lives for ONE weight tile (~0.5 k cycles of MMAs), less than a TMA round trip, so the ring
    // has to be deep in stages, not in bytes.  PHENDIFF_B200_HALO_LIN_SA overrides.
    if (d.ksize == 1 && d.mode == TC_MODE_STD && d.C >= 128) {   // (conv_in's K = 64 GEMM is one block per tile: HBM-bound, measured no gain)
        static const int lin_sa = [] { const char* e = getenv("PHENDIFF_B200_HALO_LIN_SA"); return e ? atoi(e) : 4; }();
        p.SA = std::min(4, std::max(2, lin_sa));
    }
    if (d.gn_coef) {
        // TMA latency + transform must fit under (SA - 1) K-blocks of MMAs; the 41 KB stages of the dual-accumulator tile take
        // the longest to transform and leave room for a third stage beside a 4-deep weight ring
        if (mt == 2 && pl->block_n != 16) p.SA = 3;
        if (const char* e = getenv("PHENDIFF_B200_HALO_GN_SA")) p.SA = std::min(4, std::max(2, atoi(e)));
    }
    if (d.Csc1 >= 128 || d.Csc2 >= 128) {
        // two shortcut blocks per A stage where that keeps the stage count and a 4-deep weight ring: the 16x8-pixel tiles
        // (2 x 16 KB beside a 23 KB halo tile).  Same-box A/B (profiles/r2f): fused conv2+shortcut layers +8..14 %.  On the
        // 16x16-pixel tiles it would cost the GN variant its third A stage (2 x 64 KB): measured slower, not done.
        // PHENDIFF_B200_HALO_SC2=0 keeps one block per stage.
        static const int sc2 = [] { const char* e = getenv("PHENDIFF_B200_HALO_SC2"); return e ? atoi(e) : 1; }();
        const int stage2 = ((std::max(p.a_bytes_main, 2 * p.a_bytes_sc) + 1023) / 1024) * 1024;
        if (sc2 && stage2 <= 48 * 1024 && (budget - p.SA * stage2) / b_bytes >= 4) { p.sc_per_stage = 2; p.a_stage_bytes = stage2; }
    }
    p.SB = std::min(16, (budget - p.SA * p.a_stage_bytes) / b_bytes);
    if (p.SB < 2) { delete pl; set_error("conv_halo: shared memory budget too small"); return 1; }
    pl->smem = (size_t)p.SA * p.a_stage_bytes + (size_t)p.SB * b_bytes + epi_bytes + 1024 + 512;
    const uint64_t C = d.C, H = d.H, W = d.W, N = d.N;
    int rc;
    {
        const void* srcs[2] = {d.x, d.x2};
        const uint64_t cs[2] = {C - (uint64_t)d.C2, (uint64_t)d.C2};
        CUtensorMap* tms[2] = {&p.tmA, &p.tmA2};
        for (int i = 0; i < 2; ++i) {
            if (!cs[i]) continue;
            uint64_t dims[4] = {cs[i], W, H, N};
            uint64_t st[3] = {cs[i] * 2, W * cs[i] * 2, H * W * cs[i] * 2};
            uint32_t box[4] = {64, (uint32_t)p.pitch_px, (uint32_t)rows, 1};
            if ((rc = tc_encode_map(tms[i], d.dt, srcs[i], 4, dims, st, box))) { delete pl; return rc; }
        }
    }
    const void* scs[2] = {d.sc1, d.sc2};
    const int cscs[2] = {d.Csc1, d.Csc2};
    CUtensorMap* tms[2] = {&p.tmS1, &p.tmS2};
    for (int i = 0; i < 2; ++i) {
        if (!cscs[i]) continue;
        uint64_t Cs = cscs[i];
        uint64_t dims[4] = {Cs, W, H, N};
        uint64_t st[3] = {Cs * 2, W * Cs * 2, H * W * Cs * 2};
        uint32_t box[4] = {64, (uint32_t)tile_w, HL_HT, 1};
        if ((rc = tc_encode_map(tms[i], d.dt, scs[i], 4, dims, st, box))) { delete pl; return rc; }
    }
    {
        const uint64_t Ktot = (uint64_t)(p.ntaps * p.kb_main + p.kb_s1 + p.kb_s2) * 64;
        uint64_t dims[2] = {Ktot, (uint64_t)cout_rows * p.phases};
        uint64_t st[1] = {Ktot * 2};
        uint32_t box[2] = {64, (uint32_t)pl->block_n};
        if ((rc = tc_encode_map(&p.tmB, d.dt, d.wmat, 2, dims, st, box))) { delete pl; return rc; }
    }
    if (d.mode == TC_MODE_STD) {
        const uint64_t Co = d.Cout;
        if (!d.upsample) {
            uint64_t dims[4] = {Co, (uint64_t)d.Wo, (uint64_t)d.Ho, N};
            uint64_t st[3] = {Co * 2, d.Wo * Co * 2, (uint64_t)d.Ho * d.Wo * Co * 2};
            uint32_t box[4] = {64, HL_WT, 4, 1};
            rc = tc_encode_map(&p.tmOut, d.dt, d.out, 4, dims, st, box);
        } else {
            // output pixel (2i+a, 2j+b): view {C, b:2, j:W, a:2, (n,i): N*H}; (n,i) merge because image pitch = H * (4*W*C)
            uint64_t dims[5] = {Co, 2, W, 2, N * H};
            uint64_t st[4] = {Co * 2, 2 * Co * 2, 2 * W * Co * 2, 4 * W * Co * 2};
            uint32_t box[5] = {64, 1, HL_WT, 1, 4};
            rc = tc_encode_map(&p.tmOut, d.dt, d.out, 5, dims, st, box);
        }
        if (rc) { delete pl; return rc; }
    }
    p.tile_begin = 0; p.tile_end = p.m_tiles * p.phases * p.n_tiles;
    pl->grid = std::min(p.tile_end, tc_num_sms());
    *out = pl;
    return 0;
}

void conv_halo_plan_destroy(ConvHaloPlan* p) { delete p; }

template <int BLOCK_N, typename T, int MODE, int CW, int MT, bool GN>
static int launch_halo(const ConvHaloPlan* pl, const HaloParams& p, cudaStream_t s) {
    static size_t attr_smem_dev[PD_MAX_DEVICES] = {0};
    size_t& attr_smem = attr_smem_dev[pd_cur_dev()];
    if (pl->smem > attr_smem) {
        PD_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, T, MODE, CW, MT, GN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)pl->smem));
        attr_smem = pl->smem;
    }
    if (GN) {
        // Programmatic dependent launch: the kernel ahead of a GN-variant conv is its gn_coef_kernel, which releases its
        // dependents at once (griddepcontrol.launch_dependents).  This grid may therefore start while the coefficients are
        // still being computed: barrier init, TMEM allocation, the first weight and halo-tile loads all run under gn_coef;
        // only the transform warps wait (griddepcontrol.wait) before they read the table.  PHENDIFF_B200_PDL=0 disables.
        static const int pdl = [] { const char* e = getenv("PHENDIFF_B200_PDL"); return e ? atoi(e) : 1; }();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(std::min(pl->grid, p.tile_end - p.tile_begin)); cfg.blockDim = dim3(HALO_THREADS_GN); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        PD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BLOCK_N, T, MODE, CW, MT, GN>, p));
        return 0;
    }
    conv_halo_kernel<BLOCK_N, T, MODE, CW, MT, GN><<<std::min(pl->grid, p.tile_end - p.tile_begin), HALO_THREADS, pl->smem, s>>>(p);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BLOCK_N, typename T, bool GN>
static int launch_halo_std(const ConvHaloPlan* pl, cudaStream_t s) {
    const bool cw2 = pl->p.epi.stats != nullptr && pl->p.epi.stats_cw == 2;
    if (BLOCK_N <= 128 && pl->p.mt == 2)
        return cw2 ? launch_halo<(BLOCK_N <= 128 ? BLOCK_N : 128), T, TC_MODE_STD, 2, 2, GN>(pl, pl->p, s)
                   : launch_halo<(BLOCK_N <= 128 ? BLOCK_N : 128), T, TC_MODE_STD, 4, 2, GN>(pl, pl->p, s);
    return cw2 ? launch_halo<BLOCK_N, T, TC_MODE_STD, 2, 1, GN>(pl, pl->p, s) : launch_halo<BLOCK_N, T, TC_MODE_STD, 4, 1, GN>(pl, pl->p, s);
}

int conv_halo_launch(const ConvHaloPlan* pl, cudaStream_t s, const ConvTcLaunch* extra) {
    if (pl->mode == TC_MODE_DDIM) {
        PD_REQUIRE(extra != nullptr, "conv_out plan needs per-launch outputs");
        HaloParams p = pl->p;
        p.epi.model_out = extra->model_out;
        p.epi.x_t = extra->x_t;
        p.epi.cfg = extra->cfg;
        if (extra->img_count > 0) {
            // conv_out tiles are (image, tile-in-image) with one N tile and one phase: an image range is a tile range
            const int per_img = p.tilesW * p.tilesH;
            PD_REQUIRE(p.n_tiles == 1 && p.phases == 1 && extra->img_begin >= 0 && (extra->img_begin + extra->img_count) * per_img <= p.tile_end,
                       "conv_out image range outside the plan");
            p.tile_begin = extra->img_begin * per_img;
            p.tile_end = (extra->img_begin + extra->img_count) * per_img;
            p.epi.img_off = extra->img_begin;
        }
        if (extra->x_t) {
            PD_REQUIRE(extra->step != nullptr, "x_t update needs step coefficients");
            PD_REQUIRE(extra->step->sigma == 0.f, "fused conv_out update requires eta == 0");
            p.epi.step = *extra->step;
        }
        PD_DISPATCH_HALF(pl->dt, T, {
            if (pl->gn) return p.mt == 2 ? launch_halo<16, T, TC_MODE_DDIM, 4, 2, true>(pl, p, s) : launch_halo<16, T, TC_MODE_DDIM, 4, 1, true>(pl, p, s);
            return p.mt == 2 ? launch_halo<16, T, TC_MODE_DDIM, 4, 2, false>(pl, p, s) : launch_halo<16, T, TC_MODE_DDIM, 4, 1, false>(pl, p, s);
        });
    }
    if (pl->gn) {
        PD_DISPATCH_HALF(pl->dt, T, {
            switch (pl->block_n) {
                case 256: return launch_halo_std<256, T, true>(pl, s);
                case 128: return launch_halo_std<128, T, true>(pl, s);
            }
        });
        set_error("conv_halo: fused GroupNorm needs a 128- or 256-channel output tile");
        return 1;
    }
    PD_DISPATCH_HALF(pl->dt, T, {
        switch (pl->block_n) {
            case 256: return launch_halo_std<256, T, false>(pl, s);
            case 128: return launch_halo_std<128, T, false>(pl, s);
            case 64: return launch_halo_std<64, T, false>(pl, s);
        }
    });
    set_error("conv_halo: bad block_n");
    return 1;
}

}  // namespace pd
