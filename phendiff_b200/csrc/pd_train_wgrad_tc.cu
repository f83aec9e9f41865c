// phendiff_b200 — weight gradient of a stride-1 'same' convolution on the tcgen05 tensor cores (training step, SURVEY §8 row f2;
// the reference gets it from torch.autograd: accelerator.backward(loss), src/utils_training.py:436).
//
//   stage[tap][co][ci] += sum over pixels p of dY[p, co] * X[p shifted by tap, ci]
//
// is a GEMM whose CONTRACTION runs over pixels: M = 128 output channels, N = 128 input channels, K = pixels.  In NHWC memory the
// channel is the contiguous index, so both operands are "MN-major" for the tensor core: a TMA box of [pixel rows] x [64 channels]
// (128-byte rows, SWIZZLE_128B) IS the canonical MN-major SW128 operand tile, pixels running along K in 8-row swizzle atoms
// (SBO = 1024 B), the second 64-channel box LBO bytes further.
//
// 3x3: one CTA owns one kernel ROW r of the 3x3 stencil for its (co tile, ci tile): three accumulators (taps s = 0, 1, 2) of
// 128 x 128 fp32 = 384 TMEM columns.  A K block is R image rows (R * W = 128 pixels) loaded with a ROW PITCH OF W + 2: the dY box
// starts at column 0 (columns W, W + 1 are out of bounds -> zero filled), the X box at column -1 and image row y + r - 1 (zero
// filled borders = the convolution's padding).  Tap s is then the SAME X tile addressed s rows (s * 128 bytes) further: dY pixel
// (j, w) meets X(j + r - 1, w + s - 1); the products that would wrap around a row end hit a zero dY row.  The K extent is rounded
// up to 144 rows; the tail rows of every ring slot are zeroed once and never written by TMA.
// 1x1: one accumulator, K blocks of 128 flat pixels.
// 3x3 stride 2 (Downsample2D, padding 1): dY pixel (oh, ow) meets X(2 oh + r - 1, 2 ow + s - 1).  X is viewed as (2C, W/2, 2, H/2, N): the
// column parity rides in the channel index, the row parity is its own dimension.  Kernel row r fixes the row parity and a row offset
// (r = 0: parity 1, row oh - 1; r = 1: parity 0, row oh; r = 2: parity 1, row oh); the three taps of the row need BOTH column-parity
// tiles, each loaded from column -1: s = 0 -> parity-1 tile + 0 rows, s = 1 -> parity-0 tile + 1 row, s = 2 -> parity-1 tile + 1 row
// (6 boxes per ring slot, 2 slots).
//
// Split-K over pixel blocks fills the machine (a 128 -> 128 layer has 3 tiles): every CTA adds its partial tile to the staging
// buffer with 16-byte vector reductions; launch_wgrad_unstage folds it into the OIHW gradient.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (whole warp in uniform code, one elected lane per instruction — see
// profiles/r4_attention_tc3.md for why), warps 2-5 epilogue (TMEM lane quarter = warp % 4).  3-slot ring, one CTA per SM.
#include <cuda.h>

#include <algorithm>
#include <cstring>
#include <memory>

#include "pd_tc_common.cuh"
#include "pd_train.h"

namespace pd {

namespace {

constexpr int WG_ROWS = 144;                       // K rows per ring slot (9 MMA K steps)
constexpr int WG_BOX_BYTES = WG_ROWS * 128;        // one 64-channel box: 18 KB (a multiple of 1024)
constexpr int WG_SLOT_BYTES = 4 * WG_BOX_BYTES;    // dY co 0-63, 64-127, X ci 0-63, 64-127
constexpr int WG_SLOTS = 3;
constexpr int WG_SLACK = 512;                      // tap-shifted reads run up to 2 rows past the last box
constexpr int WG_SMEM = WG_SLOTS * WG_SLOT_BYTES + WG_SLACK + 1024 /* alignment */ + 128 /* barriers */;
constexpr int WG_THREADS = 192;

struct WgParams {
    int taps;            // 3 (one row of a 3x3 stencil per CTA) or 1
    int ksteps;          // MMA K steps per block: 9 (3x3) or 8 (1x1)
    int box_rows;        // rows the TMA boxes write per slot
    int C1, Ctot, Cout;
    int tiles_m, tiles_n, rgroups;   // rgroups: 3 for 3x3 (kernel rows), 1 for 1x1
    int ksplit, nkb;     // K blocks in total, split into ksplit contiguous ranges
    int blocks_per_img;  // 3x3: H / R
    int R;
    int stride2;         // 3x3 stride-2 (padding 1) convolution: X is addressed through its (row parity, column parity) phase view
    int nslots, slot_bytes;
    float* stage;        // (taps_total, Cout, Ctot)
};

__device__ __forceinline__ uint64_t mn_sw128_desc(uint32_t smem_addr) {
    // MN-major SWIZZLE_128B: LBO = byte distance between the two 64-channel boxes, SBO = 1024 B between 8-row K groups
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((WG_BOX_BYTES >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x1,
                const __grid_constant__ CUtensorMap tm_x2, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(sm + WG_SLOTS * WG_SLOT_BYTES + WG_SLACK);     // (3 slots x 4 boxes = 2 slots x 6 boxes)
    uint64_t* empty = full + WG_SLOTS;
    uint64_t* accum = empty + WG_SLOTS;
    uint32_t* tmem_slot = (uint32_t*)(accum + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item: (tile_m, tile_n, r) x K split
    int id = blockIdx.x;
    const int split = id % p.ksplit; id /= p.ksplit;
    const int rg = id % p.rgroups; id /= p.rgroups;
    const int tn = id % p.tiles_n;
    const int tm = id / p.tiles_n;
    const int kb0 = (int)(((long long)p.nkb * split) / p.ksplit), kb1 = (int)(((long long)p.nkb * (split + 1)) / p.ksplit);
    if (kb1 <= kb0) return;

    // zero the tail rows of every box (rows the TMA boxes never write: they stay zero for the whole kernel) and the slack behind the ring,
    // make it visible to the async proxy
    {
        const int tail16 = (WG_ROWS - p.box_rows) * 8;        // 16-byte units per box tail
        for (int i = threadIdx.x; i < 12 * tail16; i += WG_THREADS) {
            const int box = i / tail16, o = i - box * tail16;
            ((uint4*)(sm + (size_t)box * WG_BOX_BYTES + (size_t)p.box_rows * 128))[o] = make_uint4(0, 0, 0, 0);
        }
        for (int i = threadIdx.x; i < WG_SLACK / 16; i += WG_THREADS) ((uint4*)(sm + WG_SLOTS * WG_SLOT_BYTES))[i] = make_uint4(0, 0, 0, 0);
        // tap-shifted reads of a box run up to 2 rows into the box behind it: before the first TMA write lands there, those rows must
        // hold finite values (they meet zero dY rows, and 0 x NaN would poison the accumulator)
        for (int i = threadIdx.x; i < 12 * 16; i += WG_THREADS) ((uint4*)(sm + (size_t)(i >> 4) * WG_BOX_BYTES))[i & 15] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < WG_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(accum, 1);
        fence_barrier_init();
        prefetch_tmap(&tm_dy); prefetch_tmap(&tm_x1); prefetch_tmap(&tm_x2);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = (p.stride2 ? 6u : 4u) * (uint32_t)p.box_rows * 128u;
            const int co0 = tm * 128, ci0 = tn * 128;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, slot = it % p.nslots;
                if (it >= p.nslots) mbar_wait(&empty[slot], (uint32_t)((it / p.nslots) - 1) & 1u);
                uint8_t* base = sm + slot * p.slot_bytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[slot])), "r"(tx) : "memory");
                if (p.stride2) {
                    const int n = kb / p.blocks_per_img, y0 = (kb % p.blocks_per_img) * p.R;
                    const int hp = rg == 1 ? 0 : 1, hh0 = y0 + (rg == 0 ? -1 : 0);
                    for (int b = 0; b < 2; ++b) tma_load_4d(&tm_dy, &full[slot], base + b * WG_BOX_BYTES, co0 + 64 * b, 0, y0, n);
                    for (int wp = 0; wp < 2; ++wp)
                        for (int b = 0; b < 2; ++b)
                            tma_load_5d(&tm_x1, &full[slot], base + (2 + 2 * wp + b) * WG_BOX_BYTES, ci0 + 64 * b + wp * p.C1, -1, hp, hh0, n);
                } else if (p.taps == 3) {
                    const int n = kb / p.blocks_per_img, y0 = (kb % p.blocks_per_img) * p.R;
                    for (int b = 0; b < 2; ++b) tma_load_4d(&tm_dy, &full[slot], base + b * WG_BOX_BYTES, co0 + 64 * b, 0, y0, n);
                    for (int b = 0; b < 2; ++b) {
                        const int c = ci0 + 64 * b;
                        if (c < p.C1) tma_load_4d(&tm_x1, &full[slot], base + (2 + b) * WG_BOX_BYTES, c, -1, y0 + rg - 1, n);
                        else tma_load_4d(&tm_x2, &full[slot], base + (2 + b) * WG_BOX_BYTES, c - p.C1, -1, y0 + rg - 1, n);
                    }
                } else {
                    for (int b = 0; b < 2; ++b) tma_load_2d(&tm_dy, &full[slot], base + b * WG_BOX_BYTES, co0 + 64 * b, kb * 128);
                    for (int b = 0; b < 2; ++b) {
                        const int c = ci0 + 64 * b;
                        if (c < p.C1) tma_load_2d(&tm_x1, &full[slot], base + (2 + b) * WG_BOX_BYTES, c, kb * 128);
                        else tma_load_2d(&tm_x2, &full[slot], base + (2 + b) * WG_BOX_BYTES, c - p.C1, kb * 128);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D fp32, A / B 16-bit, both MN-major (bits 15 / 16), N = 128, M = 128
        constexpr uint32_t fmt = std::is_same<T, bf16>::value ? 1u : 0u;
        constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sm0 = smem_u32(sm);
        const uint64_t dproto = mn_sw128_desc(0);
        const uint32_t hi = (uint32_t)(dproto >> 32), lo_flags = (uint32_t)dproto;   // lo word: LBO field (bits 16..29), address added below
        const int ksteps = p.ksteps, taps = p.taps;
        for (int kb = kb0; kb < kb1; ++kb) {
            const int it = kb - kb0, slot = it % p.nslots;
            mbar_wait(&full[slot], (uint32_t)(it / p.nslots) & 1u);
            tc_fence_after();
            const uint32_t a0 = lo_flags + ((sm0 + (uint32_t)slot * (uint32_t)p.slot_bytes) >> 4);
            const uint32_t b0 = a0 + ((2u * WG_BOX_BYTES) >> 4);
            // B start of the three taps relative to b0 (16-byte units): stride 1: the same tile 0 / 1 / 2 rows further; stride 2: the
            // column-parity-1 tile (2 boxes further), the parity-0 tile + 1 row, the parity-1 tile + 1 row
            const uint32_t o0 = p.stride2 ? ((2u * WG_BOX_BYTES) >> 4) : 0u;
            const uint32_t o1 = 128u >> 4;
            const uint32_t o2 = p.stride2 ? ((2u * WG_BOX_BYTES + 128u) >> 4) : (256u >> 4);
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t acc = (it | ks) ? 1u : 0u;
                const uint32_t a_lo = a0 + (uint32_t)ks * (2048u >> 4), b_lo = b0 + (uint32_t)ks * (2048u >> 4);
                if (elect_one()) {
                    if (taps == 3) {
                        umma_f16kind(tmem, desc64(hi, a_lo), desc64(hi, b_lo + o0), idesc, acc);
                        umma_f16kind(tmem + 128u, desc64(hi, a_lo), desc64(hi, b_lo + o1), idesc, acc);
                        umma_f16kind(tmem + 256u, desc64(hi, a_lo), desc64(hi, b_lo + o2), idesc, acc);
                    } else {
                        umma_f16kind(tmem, desc64(hi, a_lo), desc64(hi, b_lo), idesc, acc);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(&empty[slot]);
                if (kb == kb1 - 1) umma_commit(accum);
            }
            __syncwarp();
        }
    } else {
        // epilogue: this warp's TMEM lane quarter = 32 output channels; lane = one co row, 32 ci columns per load
        mbar_wait(accum, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int co = tm * 128 + q * 32 + lane;
        const int ktaps = p.rgroups == 3 ? 9 : 1;
        (void)ktaps;
        for (int s = 0; s < p.taps; ++s) {
            const int tap = p.rgroups == 3 ? rg * 3 + s : 0;
            float* row = p.stage + ((size_t)tap * p.Cout + co) * p.Ctot + tn * 128;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 128 + c0), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    red_add_v4(row + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

struct WgradTcPlan {
    CUtensorMap tm_dy, tm_x1, tm_x2;
    WgParams p;
    int dt, grid;
};

bool wgrad_tc_supported(const WgradTcDesc& d, std::string* why) {
    auto no = [&](const char* m) { if (why) *why = m; return false; };
    if (d.dt != DT_BF16 && d.dt != DT_F16) return no("tensor-core wgrad takes 16-bit operands");
    if (d.ksize != 1 && d.ksize != 3) return no("kernel size must be 1 or 3");
    if (d.Cout % 128 != 0 || (d.C1 + d.C2) % 128 != 0 || d.C1 % 64 != 0 || d.C2 % 64 != 0) return no("channel counts must tile into 128 x 128 (sources in 64s)");
    if (d.stride == 2) {
        if (d.ksize != 3 || d.C2 != 0 || d.H % 2 || d.W % 2) return no("stride 2: 3x3 over one source with even extents");
        const int Wo = d.W / 2, Ho = d.H / 2;
        if (Wo > 128 || Wo < 16 || 128 % Wo != 0 || Ho % (128 / Wo) != 0) return no("stride 2: output row width must divide 128");
        return true;
    }
    if (d.ksize == 3) {
        if (d.W > 128 || d.W < 16 || 128 % d.W != 0) return no("3x3: row width must divide 128");
        const int R = 128 / d.W;
        if (d.H % R != 0) return no("3x3: image height must be a multiple of 128 / W");
    }
    return true;
}

int wgrad_tc_plan_create(const WgradTcDesc& d, WgradTcPlan** out) {
    std::string why;
    PD_REQUIRE(wgrad_tc_supported(d, &why), ("wgrad_tc: unsupported shape: " + why).c_str());
    auto pl = std::make_unique<WgradTcPlan>();
    WgParams& p = pl->p;
    memset(&p, 0, sizeof(p));
    pl->dt = d.dt;
    p.C1 = d.C1; p.Ctot = d.C1 + d.C2; p.Cout = d.Cout; p.stage = d.stage;
    p.tiles_m = d.Cout / 128; p.tiles_n = p.Ctot / 128;
    const uint64_t N = d.N, H = d.H, W = d.W;
    int rc = 0;
    auto map4 = [&](CUtensorMap* tm, const void* base, uint64_t C, int R) {
        uint64_t dims[4] = {C, W, H, N};
        uint64_t st[3] = {C * 2, W * C * 2, H * W * C * 2};
        uint32_t box[4] = {64, (uint32_t)(W + 2), (uint32_t)R, 1};
        return tc_encode_map(tm, d.dt, base, 4, dims, st, box, true);
    };
    auto map2 = [&](CUtensorMap* tm, const void* base, uint64_t C) {
        uint64_t dims[2] = {C, N * H * W};
        uint64_t st[1] = {C * 2};
        uint32_t box[2] = {64, 128};
        return tc_encode_map(tm, d.dt, base, 2, dims, st, box, true);
    };
    p.nslots = WG_SLOTS; p.slot_bytes = WG_SLOT_BYTES;
    if (d.stride == 2) {
        const uint64_t Wo = W / 2, Ho = H / 2, C = d.C1;
        const int R = 128 / (int)Wo;
        p.stride2 = 1; p.nslots = 2; p.slot_bytes = 6 * WG_BOX_BYTES;
        p.taps = 3; p.rgroups = 3; p.ksteps = 9; p.R = R; p.box_rows = R * ((int)Wo + 2); p.blocks_per_img = (int)Ho / R; p.nkb = d.N * p.blocks_per_img;
        {
            uint64_t dims[4] = {(uint64_t)d.Cout, Wo, Ho, N};
            uint64_t st[3] = {(uint64_t)d.Cout * 2, Wo * d.Cout * 2, Ho * Wo * d.Cout * 2};
            uint32_t box[4] = {64, (uint32_t)(Wo + 2), (uint32_t)R, 1};
            if ((rc = tc_encode_map(&pl->tm_dy, d.dt, d.dy, 4, dims, st, box, true))) return rc;
        }
        {
            uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, N};
            uint64_t st[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            uint32_t box[5] = {64, (uint32_t)(Wo + 2), 1, (uint32_t)R, 1};
            if ((rc = tc_encode_map(&pl->tm_x1, d.dt, d.x1, 5, dims, st, box, true))) return rc;
            pl->tm_x2 = pl->tm_x1;
        }
    } else if (d.ksize == 3) {
        const int R = 128 / d.W;
        p.taps = 3; p.rgroups = 3; p.ksteps = 9; p.R = R; p.box_rows = R * (d.W + 2); p.blocks_per_img = d.H / R; p.nkb = d.N * p.blocks_per_img;
        if ((rc = map4(&pl->tm_dy, d.dy, d.Cout, R))) return rc;
        if ((rc = map4(&pl->tm_x1, d.x1, d.C1, R))) return rc;
        if ((rc = map4(&pl->tm_x2, d.C2 ? d.x2 : d.x1, d.C2 ? d.C2 : d.C1, R))) return rc;
    } else {
        p.taps = 1; p.rgroups = 1; p.ksteps = 8; p.R = 0; p.box_rows = 128; p.blocks_per_img = 0; p.nkb = (int)((N * H * W + 127) / 128);
        if ((rc = map2(&pl->tm_dy, d.dy, d.Cout))) return rc;
        if ((rc = map2(&pl->tm_x1, d.x1, d.C1))) return rc;
        if ((rc = map2(&pl->tm_x2, d.C2 ? d.x2 : d.x1, d.C2 ? d.C2 : d.C1))) return rc;
    }
    const int tiles = p.tiles_m * p.tiles_n * p.rgroups;
    const int sms = tc_num_sms();
    // split K so that the grid fills two whole waves of one-CTA-per-SM (a third, partial wave cost 24 % on the 48-tile layers: ncu r7e,
    // grid 336 on 148 SMs), keeping at least 8 K blocks per CTA
    int ks = std::max(1, (2 * sms) / tiles);
    ks = std::min(ks, std::max(1, p.nkb / 8));
    p.ksplit = ks;
    pl->grid = tiles * ks;
    *out = pl.release();
    return 0;
}

void wgrad_tc_plan_destroy(WgradTcPlan* p) { delete p; }

int wgrad_tc_launch(const WgradTcPlan* pl, cudaStream_t s) {
    static bool attr_done[PD_MAX_DEVICES][2] = {};
    const int dev = pd_cur_dev();
    const int ti = pl->dt == DT_BF16 ? 0 : 1;
    if (!attr_done[dev][ti]) {
        if (pl->dt == DT_BF16) PD_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        else PD_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<f16>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        attr_done[dev][ti] = true;
    }
    if (pl->dt == DT_BF16) wgrad_tc_kernel<bf16><<<pl->grid, WG_THREADS, WG_SMEM, s>>>(pl->tm_dy, pl->tm_x1, pl->tm_x2, pl->p);
    else wgrad_tc_kernel<f16><<<pl->grid, WG_THREADS, WG_SMEM, s>>>(pl->tm_dy, pl->tm_x1, pl->tm_x2, pl->p);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pd
