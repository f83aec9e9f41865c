// phendiff_b200 — CUDA-core kernels: embeddings, GroupNorm(+SiLU), SIMT convolution (fp32 validation mode and the
// odd-shaped conv_in / conv_out of the bf16 path), fused conv_out + DDIM update, attention (SIMT), scheduler and
// pipeline elementwise work, weight re-layout.  All of these are HBM- or latency-bound side work; the dense
// contractions of the bf16 path live in pd_conv_tc.cu (tcgen05) and pd_attn_mma.cu.
#include "pd_kernels.h"
#include <algorithm>
#include <vector>

namespace pd {

// =====================================================================================================================
// time + class embedding  (reference: cond_unet_2d.py:289-309; diffusers get_timestep_embedding / TimestepEmbedding)
// =====================================================================================================================
__global__ void __launch_bounds__(256) embed_kernel(EmbedArgs a, const float* __restrict__ freqs) {
    extern __shared__ float sm[];
    float* e0 = sm;           // C0
    float* h1 = sm + a.C0;    // D
    // grid (rows, D/64): every block recomputes the (cheap) first layer and produces 64 outputs of the second, so the
    // two dependent mat-vecs of the 2-row DDIB case spread over 2 x D/64 SMs instead of running on 2
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int j_begin = blockIdx.y * 64, j_end = min(j_begin + 64, a.D);
    // dedupe: block b computes the row of class b at the (single) timestep; else the row of image b
    const float t = a.timesteps ? a.timesteps[a.dedupe ? 0 : b] : a.t_scalar;
    if (b == 0 && blockIdx.y == 0) {
        for (int i = tid; i < a.B; i += blockDim.x) {
            int32_t r = i;
            if (a.dedupe) r = (a.ncls == 0 || (a.cfg_pairs > 0 && i >= a.cfg_pairs)) ? a.ncls : (int32_t)a.labels[i];   // row ncls: unconditional
            a.row_idx[i] = r;
        }
    }
    const int half = a.C0 / 2;
    for (int k = tid; k < half; k += blockDim.x) {
        float arg = t * freqs[k];
        float s = sinf(arg), c = cosf(arg);
        if (a.flip) { e0[k] = c; e0[half + k] = s; } else { e0[k] = s; e0[half + k] = c; }
    }
    __syncthreads();
    for (int j = warp; j < a.D; j += nwarp) {
        const float* w = a.w1 + (size_t)j * a.C0;
        float acc = 0.f;
        for (int k = lane; k < a.C0; k += 32) acc += w[k] * e0[k];
        acc = warp_sum(acc);
        if (lane == 0) h1[j] = silu<true>(acc + a.b1[j]);
    }
    __syncthreads();
    for (int j = j_begin + warp; j < j_end; j += nwarp) {
        const float* w = a.w2 + (size_t)j * a.D;
        float acc = 0.f;
        for (int k = lane; k < a.D; k += 32) acc += w[k] * h1[k];
        acc = warp_sum(acc);
        if (lane == 0) {
            float e = acc + a.b2[j];
            if (a.dedupe) e += (b < a.ncls) ? a.class_table[(size_t)b * a.D + j] : 0.f;   // row ncls = class_emb of zeros (guidance pass)
            else if (a.class_emb) e += a.class_emb[(size_t)b * a.D + j];
            else if (a.labels && a.class_table) e += a.class_table[(size_t)a.labels[b] * a.D + j];
            a.emb_act[(size_t)b * a.D + j] = silu<true>(e);
        }
    }
}

struct FreqTable { float* dev = nullptr; int c0 = -1; float shift = 0.f; };
static FreqTable g_freq_tables[PD_MAX_DEVICES];   // per device, (<= 4096 entries); rebuilt when (C0, shift) change

int launch_embed(const EmbedArgs& a, cudaStream_t s) {
    const int half = a.C0 / 2;
    FreqTable& ft = g_freq_tables[pd_cur_dev()];
    float*& g_freqs = ft.dev;
    if (g_freqs == nullptr || ft.c0 != a.C0 || ft.shift != a.shift) {
        PD_REQUIRE(half <= 4096, "time embedding too wide");
        if (!g_freqs) PD_CHECK_CUDA(cudaMalloc(&g_freqs, 4096 * sizeof(float)));
        std::vector<float> f(half);
        for (int k = 0; k < half; ++k) {
            // fp32 arithmetic in the order diffusers uses: (-ln(10000) * k) / (half - shift), then exp
            float e = (-9.210340371976184f * (float)k) / ((float)half - a.shift);
            f[k] = expf(e);
        }
        PD_CHECK_CUDA(cudaMemcpyAsync(g_freqs, f.data(), half * sizeof(float), cudaMemcpyHostToDevice, s));
        PD_CHECK_CUDA(cudaStreamSynchronize(s));  // f is a local; one-time cost at first use
        ft.c0 = a.C0;
        ft.shift = a.shift;
    }
    PD_REQUIRE(a.row_idx != nullptr, "embed: row index buffer missing");
    PD_REQUIRE(!a.dedupe || (!a.class_emb && !a.timesteps && (a.ncls == 0 || (a.labels && a.class_table))), "embed: dedupe needs one scalar timestep and, for class-conditioned models, labels");
    PD_REQUIRE(a.cfg_pairs == 0 || (a.dedupe && 2 * a.cfg_pairs == a.B), "embed: the guidance pass needs labels, one scalar timestep and 2P images");
    size_t smem = (size_t)(a.C0 + a.D) * sizeof(float);
    embed_kernel<<<dim3(embed_rows(a), (a.D + 63) / 64), 256, smem, s>>>(a, g_freqs);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void __launch_bounds__(256) temb_proj_kernel(const float* __restrict__ emb_act, const float* __restrict__ wcat,
                                                         const float* __restrict__ bcat, int D, int J,
                                                         float* __restrict__ out) {
    extern __shared__ float e[];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < D; k += blockDim.x) e[k] = emb_act[(size_t)b * D + k];
    __syncthreads();
    const int j0 = blockIdx.x * 64;
    for (int jj = warp; jj < 64; jj += 8) {
        int j = j0 + jj;
        if (j >= J) break;
        const float* w = wcat + (size_t)j * D;
        float acc = 0.f;
        for (int k = lane; k < D; k += 32) acc += w[k] * e[k];
        acc = warp_sum(acc);
        if (lane == 0) out[(size_t)b * J + j] = acc + bcat[j];
    }
}

int launch_temb_proj(const float* emb_act, const float* wcat, const float* bcat, int B, int D, int J, float* out,
                     cudaStream_t s) {
    dim3 grid((J + 63) / 64, B);
    temb_proj_kernel<<<grid, 256, D * sizeof(float), s>>>(emb_act, wcat, bcat, D, J, out);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// GroupNorm (+SiLU), NHWC, two concatenated sources.  HBM-bound: stats = 1 read, apply = 1 read + 1 write.
// Thread mapping: a thread owns one 8-channel vector column `cv` and walks pixel rows, so every global access is a
// 16/32-byte vector and a warp covers contiguous memory.
// =====================================================================================================================
// chunk statistics (N, C/cw, 2): sum and sum of squares over H*W of every cw-channel chunk (cw = 4, 2 or 1)
template <typename T>
__global__ void __launch_bounds__(512) gn_chunk_stats_kernel(const T* __restrict__ x, int HW, int C, int cw, int rows_per_block,
                                                              double* __restrict__ stats) {
    const int ncv = C / 8;
    const int n = blockIdx.y;
    const int cv = threadIdx.x % ncv, r0 = threadIdx.x / ncv, rstep = blockDim.x / ncv;
    const int row_begin = blockIdx.x * rows_per_block;
    const int row_end = min(row_begin + rows_per_block, HW);
    const T* src = x + (size_t)n * HW * C + cv * 8;
    // shifted sums: the first row this thread sees is its pivot, so the fp32 partials hold deviations (size ~ std), not values
    // (size ~ mean); they are turned back into plain sum / sum of squares in fp64: sum v = n p + s, sum v^2 = q + 2 p s + n p^2
    float s[8], q[8], pv[8];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; pv[i] = 0.f; }
    for (int r = row_begin + r0; r < row_end; r += rstep) {
        float v[8];
        load8(src + (size_t)r * C, v);
        if (cnt == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pv[i] = v[i];
        }
        ++cnt;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - pv[i]; s[i] += d; q[i] += d * d; }
    }
    if (cnt == 0) return;
    double S[8], Q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double pd_ = (double)pv[i], nn = (double)cnt;
        S[i] = nn * pd_ + (double)s[i];
        Q[i] = (double)q[i] + 2.0 * pd_ * (double)s[i] + nn * pd_ * pd_;
    }
    double* dst = stats + ((size_t)n * (C / cw) + (cv * 8) / cw) * 2;
    if (cw == 4) {
        atomicAdd(dst + 0, (S[0] + S[1]) + (S[2] + S[3])); atomicAdd(dst + 1, (Q[0] + Q[1]) + (Q[2] + Q[3]));
        atomicAdd(dst + 2, (S[4] + S[5]) + (S[6] + S[7])); atomicAdd(dst + 3, (Q[4] + Q[5]) + (Q[6] + Q[7]));
    } else if (cw == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(dst + 2 * j, S[2 * j] + S[2 * j + 1]); atomicAdd(dst + 2 * j + 1, Q[2 * j] + Q[2 * j + 1]); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(dst + 2 * j, S[j]); atomicAdd(dst + 2 * j + 1, Q[j]); }
    }
}

static int gn_block(int C) {      // one 8-channel column per thread; up to 512 columns (the SD-2.1-width config concatenates to 2560 channels)
    int ncv = C / 8;
    int rpp = 256 / ncv;
    if (rpp < 1) rpp = 1;
    return ncv * rpp;
}

int launch_gn_chunk_stats(int dt, const void* x, int N, int HW, int C, int cw, double* stats, cudaStream_t s) {
    PD_REQUIRE(C % 8 == 0 && C / 8 <= 512, "GroupNorm channel count must be a multiple of 8 and <= 4096");
    PD_REQUIRE(cw == 4 || cw == 2 || cw == 1, "statistics chunk width must be 4, 2 or 1");
    const int block = gn_block(C);
    const int rstep = block / (C / 8);
    int rows = rstep * 16;   // 16 vector loads per thread, then a handful of atomics
    if (rows > HW) rows = HW;
    dim3 grid((HW + rows - 1) / rows, N);
    PD_DISPATCH_DT(dt, T, (gn_chunk_stats_kernel<T><<<grid, block, 0, s>>>((const T*)x, HW, C, cw, rows, stats)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// normalise + affine (+SiLU) of concat(x1, x2): group statistics are finalised from the sources' chunk statistics in the
// block prologue (per-channel scale / shift in shared memory), then every thread streams 8-channel vectors.

// first-iteration prefetch of gn_apply: the raw 8 elements of a row (16-bit types: one 16-byte load; fp32: handled as two
// loads folded into one uint4 pair by the fp32 specialisation below)
template <typename T> __device__ __forceinline__ uint4 ld_raw8(const T* p) { return *reinterpret_cast<const uint4*>(p); }
template <typename T> __device__ __forceinline__ void unpack_raw8(const uint4& r, float v[8]) { unpack8<T>(r, v); }
template <> __device__ __forceinline__ void unpack_raw8<float>(const uint4&, float v[8]) { for (int i = 0; i < 8; ++i) v[i] = 0.f; }   // fp32 rows are not prefetched
template <typename T, bool kPrecise>
__global__ void __launch_bounds__(512) gn_apply_kernel(GNArgs a, int rows_per_block) {
    extern __shared__ float sm[];   // scale[C], shift[C], mean[groups], rstd[groups]
    const int C = a.C1 + a.C2, ncv = C / 8;
    float* s_sc = sm;
    float* s_sh = sm + C;
    float* s_mean = sm + 2 * C;
    float* s_rstd = s_mean + a.groups;
    const int n = blockIdx.y;
    const int cpg = C / a.groups, cw = a.stats_cw;
    // this thread's column of 8 channels and its rows; the first four 16-byte loads are issued BEFORE the statistics are
    // finalised (they do not depend on them), so the block's prologue overlaps its first memory round trip
    const int cv = threadIdx.x % ncv, r0 = threadIdx.x / ncv, rstep = blockDim.x / ncv;
    const bool worker = r0 < rstep;
    const int row_begin = blockIdx.x * rows_per_block;
    const int row_end = min(row_begin + rows_per_block, a.HW);
    const int c = cv * 8;
    const T* src;
    int pitch, coff;
    if (c < a.C1) { src = (const T*)a.x1; pitch = a.C1; coff = c; } else { src = (const T*)a.x2; pitch = a.C2; coff = c - a.C1; }
    src += (size_t)n * a.HW * pitch + coff;
    T* dst = (T*)a.out + (size_t)n * a.HW * C + c;
    int r = row_begin + r0;
    uint4 pre[4];
    const bool have_pre = sizeof(T) == 2 && worker && (r + 3 * rstep < row_end);
    if (have_pre) {
#pragma unroll
        for (int u = 0; u < 4; ++u) pre[u] = ld_raw8(src + (size_t)(r + u * rstep) * pitch);
    }
    const double inv_cnt = 1.0 / ((double)cpg * (double)a.HW);
    for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
        double sum = 0.0, sq = 0.0;
        for (int cc = g * cpg; cc < (g + 1) * cpg; cc += cw) {
            const double* st = (cc < a.C1) ? a.stats1 + ((size_t)n * (a.C1 / cw) + cc / cw) * 2
                                           : a.stats2 + ((size_t)n * (a.C2 / cw) + (cc - a.C1) / cw) * 2;
            sum += st[0]; sq += st[1];
        }
        const double mean = sum * inv_cnt;
        const float var = (float)fmax(sq * inv_cnt - mean * mean, 0.0);   // the subtraction that cancels is done in fp64
        s_mean[g] = (float)mean;
        s_rstd[g] = kPrecise ? 1.0f / sqrtf(var + a.eps) : rsqrtf(var + a.eps);
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
        const int g = cc / cpg;
        const float scv = s_rstd[g] * a.gamma[cc];
        s_sc[cc] = scv;
        s_sh[cc] = a.beta[cc] - s_mean[g] * scv;
    }
    __syncthreads();
    if (!worker) return;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = s_sc[c + i]; sh[i] = s_sh[c + i]; }
    if (have_pre) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[8];
            unpack_raw8<T>(pre[u], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float y = v[i] * sc[i] + sh[i];
                v[i] = a.silu ? silu<kPrecise>(y) : y;
            }
            store8(dst + (size_t)(r + u * rstep) * C, v);
        }
        r += 4 * rstep;
    }
    for (; r + 3 * rstep < row_end; r += 4 * rstep) {   // 4 independent 16-byte loads in flight per thread
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(src + (size_t)(r + u * rstep) * pitch, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float y = v[u][i] * sc[i] + sh[i];
                v[u][i] = a.silu ? silu<kPrecise>(y) : y;
            }
            store8(dst + (size_t)(r + u * rstep) * C, v[u]);
        }
    }
    for (; r < row_end; r += rstep) {
        float v[8];
        load8(src + (size_t)r * pitch, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float y = v[i] * sc[i] + sh[i];
            v[i] = a.silu ? silu<kPrecise>(y) : y;
        }
        store8(dst + (size_t)r * C, v);
    }
}

int launch_gn_apply(int dt, bool precise, const GNArgs& a, cudaStream_t s) {
    const int C = a.C1 + a.C2;
    PD_REQUIRE(C % 8 == 0 && a.C1 % 8 == 0 && C / 8 <= 512, "GroupNorm channel count must be a multiple of 8 and <= 4096");
    PD_REQUIRE(C % a.groups == 0, "channels not divisible by groups");
    const int cw = a.stats_cw;
    PD_REQUIRE((cw == 4 || cw == 2 || cw == 1) && (C / a.groups) % cw == 0 && a.C1 % cw == 0, "statistics chunk width must divide the group width");
    PD_REQUIRE(a.stats1 && (a.C2 == 0 || a.stats2), "GroupNorm sources need chunk statistics");
    const int block = gn_block(C);
    const int rstep = block / (C / 8);
    int rows = rstep * (a.HW >= 4096 ? 32 : 16);   // 8 (4) iterations of 4 rows per thread: the prologue is amortised over >= 64 KB
    if (rows > a.HW) rows = a.HW;
    dim3 grid((a.HW + rows - 1) / rows, a.N);
    const size_t smem = (size_t)(2 * C + 2 * a.groups) * sizeof(float);
    if (precise) PD_DISPATCH_DT(dt, T, (gn_apply_kernel<T, true><<<grid, block, smem, s>>>(a, rows)));
    else PD_DISPATCH_DT(dt, T, (gn_apply_kernel<T, false><<<grid, block, smem, s>>>(a, rows)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// GroupNorm folded to per-image per-channel coefficients for the convolution that applies it to its own input tiles
// (pd_conv_halo.cu, GN variant): coef[n, c] = (rstd[n,g] * gamma[c], beta[c] - mean[n,g] * rstd[n,g] * gamma[c]).  Same
// finalisation of the producers' chunk statistics as gn_apply_kernel's prologue (groups may straddle the two sources).
__global__ void __launch_bounds__(256) gn_coef_kernel(GNArgs a, float2* coef) {
    // let the consuming convolution (launched with programmatic stream serialization) start its prologue and first tile
    // loads now; its transform warps wait for this grid to finish before they read `coef`
    asm volatile("griddepcontrol.launch_dependents;");
    extern __shared__ float sm[];   // mean[groups], rstd[groups]
    float* s_mean = sm;
    float* s_rstd = sm + a.groups;
    const int C = a.C1 + a.C2, n = blockIdx.x;
    const int cpg = C / a.groups, cw = a.stats_cw;
    const double inv_cnt = 1.0 / ((double)cpg * (double)a.HW);
    for (int g = threadIdx.x; g < a.groups; g += blockDim.x) {
        double sum = 0.0, sq = 0.0;
        for (int cc = g * cpg; cc < (g + 1) * cpg; cc += cw) {
            const double* st = (cc < a.C1) ? a.stats1 + ((size_t)n * (a.C1 / cw) + cc / cw) * 2
                                           : a.stats2 + ((size_t)n * (a.C2 / cw) + (cc - a.C1) / cw) * 2;
            sum += st[0]; sq += st[1];
        }
        const double mean = sum * inv_cnt;
        const float var = (float)fmax(sq * inv_cnt - mean * mean, 0.0);   // the subtraction that cancels is done in fp64
        s_mean[g] = (float)mean;
        s_rstd[g] = rsqrtf(var + a.eps);
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
        const int g = cc / cpg;
        const float scv = s_rstd[g] * a.gamma[cc];
        coef[(size_t)n * C + cc] = make_float2(scv, a.beta[cc] - s_mean[g] * scv);
    }
}

int launch_gn_coef(const GNArgs& a, float2* coef, cudaStream_t s) {
    const int C = a.C1 + a.C2;
    PD_REQUIRE(C % a.groups == 0, "channels not divisible by groups");
    const int cw = a.stats_cw;
    PD_REQUIRE((cw == 4 || cw == 2 || cw == 1) && (C / a.groups) % cw == 0 && a.C1 % cw == 0, "statistics chunk width must divide the group width");
    PD_REQUIRE(a.stats1 && (a.C2 == 0 || a.stats2), "GroupNorm sources need chunk statistics");
    gn_coef_kernel<<<a.N, 256, 2 * a.groups * sizeof(float), s>>>(a, coef);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// SIMT implicit-GEMM convolution, fp32 FMA.  M = N*Ho*Wo output pixels, Ncol = Cout, K = k*k*(C1+C2).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile.  This is the fp32 validation path (<= 1e-4 vs the oracle)
// and the in-CUDA fallback for shapes the tcgen05 kernel does not take.
// =====================================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int M = a.N * a.Ho * a.Wo, Ct = a.C1 + a.C2;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    // A-load role: pixel pi, 4 channels at chunk
    const int pi = tid >> 2, chunk = (tid & 3) * 4;
    const int m = m0 + pi;
    int img = 0, ho = 0, wo = 0;
    const bool mvalid = m < M;
    if (mvalid) { img = m / (a.Ho * a.Wo); int rem = m - img * a.Ho * a.Wo; ho = rem / a.Wo; wo = rem - ho * a.Wo; }
    // B-load role
    const int brow = tid >> 4, bcol = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int ntap = a.ksize * a.ksize;
    for (int tap = 0; tap < ntap; ++tap) {
        const int r = tap / a.ksize, sx = tap - r * a.ksize;
        const int ih = ho * a.stride - a.pad + r, iw = wo * a.stride - a.pad + sx;
        const bool pvalid = mvalid && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
        const size_t pix = ((size_t)img * a.H + ih) * a.W + iw;
        for (int ci0 = 0; ci0 < Ct; ci0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            if (pvalid) {
                int ci = ci0 + chunk;
                const T* p = (ci < a.C1) ? (const T*)a.x1 + pix * a.C1 + ci : (const T*)a.x2 + pix * a.C2 + (ci - a.C1);
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = to_f(p[i]);
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + bcol < a.Cout)
                bv = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * Ct + ci0 + brow) * a.Cout + n0 + bcol);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) As[chunk + i][pi] = av[i];
            *reinterpret_cast<float4*>(&Bs[brow][bcol]) = bv;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 af = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float4 bf = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                float aa[4] = {af.x, af.y, af.z, af.w}, bb[4] = {bf.x, bf.y, bf.z, bf.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += aa[i] * bb[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
        const int im = mm / (a.Ho * a.Wo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= a.Cout) continue;
            float v = acc[i][j];
            if (a.bias) v += a.bias[c];
            if (a.addvec) v += a.addvec[(size_t)(a.addvec_row ? a.addvec_row[im] : im) * a.addvec_stride + c];
            if (a.residual) v += to_f(((const T*)a.residual)[(size_t)mm * a.Cout + c]);
            v *= a.out_scale;
            ((T*)a.out)[(size_t)mm * a.Cout + c] = from_f<T>(v);
        }
    }
}

int launch_conv_simt(int dt, const ConvArgs& a, cudaStream_t s) {
    const int Ct = a.C1 + a.C2;
    PD_REQUIRE(Ct % 16 == 0 && a.C1 % 4 == 0 && a.Cout % 4 == 0, "conv_simt needs Cin % 16 == 0 and Cout % 4 == 0");
    const int M = a.N * a.Ho * a.Wo;
    dim3 grid((M + 63) / 64, (a.Cout + 63) / 64);
    PD_DISPATCH_DT(dt, T, (conv_simt_kernel<T><<<grid, 256, 0, s>>>(a)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// conv_in: (N,Cin,H,W) fp32 NCHW -> (N,H,W,Cout) NHWC, 3x3 pad 1.  K = 9*Cin = 27: bandwidth-bound, weights in smem,
// a warp covers 32 consecutive pixels of a row (coalesced NCHW reads, broadcast weight reads).
// =====================================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, int N, int Cin, int H, int W,
                                                       int Cout, T* __restrict__ out, int n_src) {
    extern __shared__ float sw[];  // (9*Cin, Cout) + bias
    const int K = 9 * Cin;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) sw[i] = w[i];
    float* sb = sw + K * Cout;
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const size_t HW = (size_t)H * W, total = (size_t)N * HW;
    const size_t p = (size_t)blockIdx.x * 32 + lane;
    if (p >= total) return;
    const int n = p / HW;
    const int hw = p - (size_t)n * HW;
    const int h = hw / W, ww = hw - h * W;
    const int ns = n % n_src;   // source image (the guidance pass reads every sample twice)
    // gather the 9*Cin input values once (Cin <= 4 supported in registers)
    float in[36];
    for (int tap = 0; tap < 9; ++tap) {
        int ih = h + tap / 3 - 1, iw = ww + tap % 3 - 1;
        bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
        for (int ci = 0; ci < Cin; ++ci)
            in[tap * Cin + ci] = ok ? x[((size_t)ns * Cin + ci) * HW + (size_t)ih * W + iw] : 0.f;
    }
    for (int cg = warp; cg * 16 < Cout; cg += nwarp) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sb[cg * 16 + j];
        for (int k = 0; k < K; ++k) {
            const float v = in[k];
            const float4* wr = reinterpret_cast<const float4*>(sw + (size_t)k * Cout + cg * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 w4 = wr[q];
                acc[4 * q + 0] += v * w4.x; acc[4 * q + 1] += v * w4.y;
                acc[4 * q + 2] += v * w4.z; acc[4 * q + 3] += v * w4.w;
            }
        }
        T* o = out + p * Cout + cg * 16;
        store8(o, acc);
        store8(o + 8, acc + 8);
    }
}

int launch_conv_in(int dt, const float* x, const float* w, const float* bias, int N, int Cin, int H, int W, int Cout,
                   void* out, cudaStream_t s, int n_src) {
    if (n_src <= 0) n_src = N;
    PD_REQUIRE(Cin <= 4 && Cout % 16 == 0, "conv_in supports in_channels <= 4 and block_out_channels[0] % 16 == 0");
    size_t smem = ((size_t)9 * Cin * Cout + Cout) * sizeof(float);
    PD_REQUIRE(smem <= 200 * 1024, "conv_in weights do not fit shared memory");
    size_t total = (size_t)N * H * W;
    int grid = (int)((total + 31) / 32);
    PD_DISPATCH_DT(dt, T, {
        PD_CHECK_CUDA(cudaFuncSetAttribute(conv_in_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_in_kernel<T><<<grid, 256, smem, s>>>(x, w, bias, N, Cin, H, W, Cout, (T*)out, n_src);
    });
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// scheduler update, shared by the standalone step kernel and the conv_out epilogue
// =====================================================================================================================
__global__ void ddim_step_kernel(pd_step_coeffs_t c, const float* __restrict__ x, const float* __restrict__ m,
                                 const float* __restrict__ noise, float* __restrict__ x_out, float* __restrict__ x0_out,
                                 int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float x0;
        float o = ddim_update(c, x[i], m[i], noise ? noise[i] : 0.f, &x0);
        if (x_out) x_out[i] = o;
        if (x0_out) x0_out[i] = x0;
    }
}

int launch_ddim_step(const pd_step_coeffs_t& c, const float* x, const float* m, const float* noise, float* x_out,
                     float* x0_out, int64_t n, cudaStream_t s) {
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    ddim_step_kernel<<<grid, 256, 0, s>>>(c, x, m, noise, x_out, x0_out, n);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// conv_out (+ fused DDIM update).  Cout <= 4, Cin = 32*CPL.  Each lane keeps its CPL input channels' 9*CPL*Cout
// weights in registers; a warp reduces one pixel at a time and handles runs of 32 consecutive pixels so that the
// NCHW fp32 writes (model output and x_t update) are coalesced.  x_t is read once and written once per step.
// =====================================================================================================================
__device__ __forceinline__ void load4(const float* p, float v[]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float v[]) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
    float2 f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}
__device__ __forceinline__ void load4(const f16* p, float v[]) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&r.x));
    float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&r.y));
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}

template <typename T, int CPL>
__global__ void __launch_bounds__(256) conv_out_kernel(ConvOutArgs a, pd_step_coeffs_t st, int has_step) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float wr[9][CPL][3];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float4 w4 = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * a.Cin + lane * CPL + j) * 4);
            wr[tap][j][0] = w4.x; wr[tap][j][1] = w4.y; wr[tap][j][2] = w4.z;
        }
    const size_t HW = (size_t)a.H * a.W;
    const size_t first = (size_t)a.img_begin * HW, total = first + (size_t)(a.img_count > 0 ? a.img_count : a.N) * HW;
    const size_t nruns = (total - first + 31) / 32;
    for (size_t run = gwarp; run < nruns; run += nwarps) {
        float keep[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < 32; ++i) {
            const size_t p = first + run * 32 + i;
            if (p >= total) break;
            const int n = p / HW;
            const int hw = p - (size_t)n * HW;
            const int h = hw / a.W, w = hw - h * a.W;
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
                if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) continue;
                const T* src = (const T*)a.act + (((size_t)n * a.H + ih) * a.W + iw) * a.Cin + lane * CPL;
                float v[CPL];
                if (CPL == 8) load8(src, v);
                else if (CPL == 4) load4(src, v);
                else {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) v[j] = to_f(src[j]);
                }
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    acc[0] += v[j] * wr[tap][j][0]; acc[1] += v[j] * wr[tap][j][1]; acc[2] += v[j] * wr[tap][j][2];
                }
            }
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                float v = warp_sum(acc[co]);
                if (lane == i) keep[co] = v;
            }
        }
        const size_t p = first + run * 32 + lane;
        if (p < total) {
            const int n = p / HW;
            const size_t hw = p - (size_t)n * HW;
            const int no = n - a.img_begin;   // outputs are indexed from the first image of the launch
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                if (co >= a.Cout) break;
                float m = keep[co] + a.bias[co];
                const size_t idx = ((size_t)no * a.Cout + co) * HW + hw;
                if (a.cfg.uncond) {
                    const float u = a.cfg.uncond[idx];
                    m = (a.cfg.eqn == 0 ? u : m) + a.cfg.w[no] * (m - u);
                }
                if (a.model_out) a.model_out[idx] = m;
                if (has_step) a.x[idx] = ddim_update(st, a.x[idx], m, 0.f, nullptr);
            }
        }
    }
}

// any channel count (the register-tiled kernel above takes 32 / 64 / 128 / 256): one warp per output pixel, lanes stride over the input
// channels, weights through L1.  Only the fp32 validation mode of unusual widths (e.g. the 320-channel SD-2.1-width config) lands here.
template <typename T>
__global__ void __launch_bounds__(256) conv_out_generic_kernel(ConvOutArgs a, pd_step_coeffs_t st, int has_step) {
    const int lane = threadIdx.x & 31;
    const size_t gwarp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t HW = (size_t)a.H * a.W;
    const size_t first = (size_t)a.img_begin * HW, total = first + (size_t)(a.img_count > 0 ? a.img_count : a.N) * HW;
    for (size_t p = first + gwarp; p < total; p += nwarps) {
        const int n = p / HW;
        const size_t hw = p - (size_t)n * HW;
        const int h = hw / a.W, w = hw - (size_t)h * a.W;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int tap = 0; tap < 9; ++tap) {
            const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
            if (ih < 0 || ih >= a.H || iw < 0 || iw >= a.W) continue;
            const T* src = (const T*)a.act + (((size_t)n * a.H + ih) * a.W + iw) * a.Cin;
            for (int c = lane; c < a.Cin; c += 32) {
                const float v = to_f(src[c]);
                const float4 w4 = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * a.Cin + c) * 4);
                acc[0] += v * w4.x; acc[1] += v * w4.y; acc[2] += v * w4.z;
            }
        }
#pragma unroll
        for (int co = 0; co < 3; ++co) acc[co] = warp_sum(acc[co]);
        if (lane == 0) {
            const int no = n - a.img_begin;
            for (int co = 0; co < 3 && co < a.Cout; ++co) {
                float m = acc[co] + a.bias[co];
                const size_t idx = ((size_t)no * a.Cout + co) * HW + hw;
                if (a.cfg.uncond) {
                    const float u = a.cfg.uncond[idx];
                    m = (a.cfg.eqn == 0 ? u : m) + a.cfg.w[no] * (m - u);
                }
                if (a.model_out) a.model_out[idx] = m;
                if (has_step) a.x[idx] = ddim_update(st, a.x[idx], m, 0.f, nullptr);
            }
        }
    }
}

int launch_conv_out(int dt, const ConvOutArgs& a, cudaStream_t s) {
    PD_REQUIRE(a.Cout <= 3, "conv_out supports out_channels <= 3");
    const int cpl = a.Cin % 32 == 0 ? a.Cin / 32 : 0;
    pd_step_coeffs_t st{};
    int has = 0;
    if (a.step) { st = *a.step; has = 1; PD_REQUIRE(st.sigma == 0.f, "fused conv_out update requires eta == 0"); }
    PD_REQUIRE(a.img_begin >= 0 && a.img_begin + a.img_count <= a.N, "conv_out image range outside the tensor");
    size_t total = (size_t)(a.img_count > 0 ? a.img_count : a.N) * a.H * a.W;
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 8);
    if (grid < 1) grid = 1;
    if (!(cpl == 1 || cpl == 2 || cpl == 4 || cpl == 8)) {
        PD_DISPATCH_DT(dt, T, (conv_out_generic_kernel<T><<<grid, 256, 0, s>>>(a, st, has)));
        PD_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    PD_DISPATCH_DT(dt, T, {
        if (cpl == 1) conv_out_kernel<T, 1><<<grid, 256, 0, s>>>(a, st, has);
        else if (cpl == 2) conv_out_kernel<T, 2><<<grid, 256, 0, s>>>(a, st, has);
        else if (cpl == 4) conv_out_kernel<T, 4><<<grid, 256, 0, s>>>(a, st, has);
        else conv_out_kernel<T, 8><<<grid, 256, 0, s>>>(a, st, has);
    });
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// nearest-neighbour 2x upsample, NHWC (Upsample2D's F.interpolate; the conv that follows runs on the result)
// =====================================================================================================================
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, int N, int H, int W, int C, T* __restrict__ out) {
    const int cv = C / 8;
    const size_t total = (size_t)N * 2 * H * 2 * W * cv;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int c = i % cv;
        size_t p = i / cv;
        int wo = p % (2 * W); p /= (2 * W);
        int ho = p % (2 * H);
        int n = p / (2 * H);
        const T* src = x + (((size_t)n * H + ho / 2) * W + wo / 2) * C + c * 8;
        T* dst = out + (((size_t)n * 2 * H + ho) * 2 * W + wo) * C + c * 8;
        if (sizeof(T) == 2) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        else { float v[8]; load8(src, v); store8(dst, v); }
    }
}

int launch_upsample2x(int dt, const void* x, int N, int H, int W, int C, void* out, cudaStream_t s) {
    PD_REQUIRE(C % 8 == 0, "upsample needs C % 8 == 0");
    size_t total = (size_t)N * 4 * H * W * (C / 8);
    int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)148 * 16);
    PD_DISPATCH_DT(dt, T, (upsample2x_kernel<T><<<grid, 256, 0, s>>>((const T*)x, N, H, W, C, (T*)out)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// conv_in on the tensor cores: im2col of the NCHW fp32 sample into (N,H,W,64) 16-bit rows, k = tap*Cin + ci (zero padded)
// =====================================================================================================================
template <typename T, int CIN>
__global__ void __launch_bounds__(256) im2col_in_kernel(const float* __restrict__ x, int N, int H, int W, T* __restrict__ out, int n_src) {
    const size_t HW = (size_t)H * W, total = (size_t)N * HW;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int n = (int)(p / HW) % n_src;   // source image (the guidance pass reads every sample twice)
    const int hw = p % HW;
    const int h = hw / W, w = hw - h * W;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int ih = h + tap / 3 - 1, iw = w + tap % 3 - 1;
        const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
            v[tap * CIN + ci] = ok ? __ldg(x + ((size_t)n * CIN + ci) * HW + (size_t)ih * W + iw) : 0.f;
    }
    T* o = out + p * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) store8(o + j * 8, v + j * 8);
}

int launch_im2col_in(int dt, const float* x, int N, int Cin, int H, int W, void* out, cudaStream_t s, int n_src) {
    PD_REQUIRE(Cin >= 1 && Cin <= 4, "im2col conv_in supports 1..4 input channels");
    if (n_src <= 0) n_src = N;
    const size_t total = (size_t)N * H * W;
    const int grid = (int)((total + 255) / 256);
    PD_DISPATCH_HALF(dt, T, {
        switch (Cin) {
            case 1: im2col_in_kernel<T, 1><<<grid, 256, 0, s>>>(x, N, H, W, (T*)out, n_src); break;
            case 2: im2col_in_kernel<T, 2><<<grid, 256, 0, s>>>(x, N, H, W, (T*)out, n_src); break;
            case 3: im2col_in_kernel<T, 3><<<grid, 256, 0, s>>>(x, N, H, W, (T*)out, n_src); break;
            default: im2col_in_kernel<T, 4><<<grid, 256, 0, s>>>(x, N, H, W, (T*)out, n_src); break;
        }
    });
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// attention core, SIMT (fp32 validation path).  One thread per query, K/V tiles of 128 keys in shared memory,
// online softmax in chunks of 16 keys.  head_dim 8.
// =====================================================================================================================
template <typename T, bool kPrecise>
__global__ void __launch_bounds__(128) attention_simt_kernel(const T* __restrict__ qkv, int S, int C, T* __restrict__ out) {
    constexpr int D = 8, TK = 128;
    __shared__ float Ks[TK][D];
    __shared__ float Vs[TK][D];
    const int n = blockIdx.z, head = blockIdx.y;
    const int qi = blockIdx.x * 128 + threadIdx.x;
    const size_t rowp = (size_t)3 * C;
    const T* base = qkv + (size_t)n * S * rowp;
    float q[D];
    const float scale = 0.35355339059327373f;  // 1/sqrt(8)
    if (qi < S) {
        float v[8];
        load8(base + (size_t)qi * rowp + head * D, v);
#pragma unroll
        for (int i = 0; i < D; ++i) q[i] = v[i] * scale;
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) q[i] = 0.f;
    }
    float mx = -INFINITY, l = 0.f, o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    for (int k0 = 0; k0 < S; k0 += TK) {
        __syncthreads();
        {
            const int kj = k0 + threadIdx.x;
            float kv[8], vv[8];
            if (kj < S) {
                load8(base + (size_t)kj * rowp + C + head * D, kv);
                load8(base + (size_t)kj * rowp + 2 * C + head * D, vv);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) { kv[i] = 0.f; vv[i] = 0.f; }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) { Ks[threadIdx.x][i] = kv[i]; Vs[threadIdx.x][i] = vv[i]; }
        }
        __syncthreads();
        const int kmax = min(TK, S - k0);
        for (int c0 = 0; c0 < kmax; c0 += 16) {
            float sc[16];
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float s = -INFINITY;
                if (c0 + j < kmax) {
                    s = 0.f;
#pragma unroll
                    for (int i = 0; i < D; ++i) s += q[i] * Ks[c0 + j][i];
                }
                sc[j] = s;
                cm = fmaxf(cm, s);
            }
            const float nm = fmaxf(mx, cm);
            const float corr = kPrecise ? expf(mx - nm) : __expf(mx - nm);
            l *= corr;
#pragma unroll
            for (int i = 0; i < D; ++i) o[i] *= corr;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (c0 + j < kmax) {
                    const float p = kPrecise ? expf(sc[j] - nm) : __expf(sc[j] - nm);
                    l += p;
#pragma unroll
                    for (int i = 0; i < D; ++i) o[i] += p * Vs[c0 + j][i];
                }
            }
            mx = nm;
        }
    }
    if (qi < S) {
        const float inv = 1.0f / l;
        float v[8];
#pragma unroll
        for (int i = 0; i < D; ++i) v[i] = o[i] * inv;
        store8(out + ((size_t)n * S + qi) * C + head * D, v);
    }
}

// Any head dimension (attention_head_dim: null = ONE head of dim C, cond_unet_2d.py:176-178,192-194,222-224; the shipped
// orig_google_ddpm_model_denoiser.json): a warp owns two queries, its lanes split the head dimension (element e of a row lives in
// lane e % 32), a score is a 32-lane shuffle reduction, softmax is online (running max and sum per query), fp32 throughout.
// K and V rows are read straight from L2 (one coalesced row per key): these layers are 0.1 % of the FLOPs of the models that
// use them, so the kernel is written for generality, not for the roofline.
template <typename T, bool kPrecise, int DPL>
__global__ void __launch_bounds__(256) attention_generic_kernel(const T* __restrict__ qkv, int S, int C, int d, float scale,
                                                                 T* __restrict__ out) {
    const int n = blockIdx.z, head = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q0 = (blockIdx.x * 8 + warp) * 2;
    if (q0 >= S) return;
    const bool two = q0 + 1 < S;
    const size_t rowp = (size_t)3 * C;
    const T* base = qkv + (size_t)n * S * rowp + (size_t)head * d;
    float qa[DPL], qb[DPL], oa[DPL], ob[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        const int e = lane + 32 * i;
        qa[i] = e < d ? to_f(base[(size_t)q0 * rowp + e]) * scale : 0.f;
        qb[i] = (e < d && two) ? to_f(base[(size_t)(q0 + 1) * rowp + e]) * scale : 0.f;
        oa[i] = 0.f; ob[i] = 0.f;
    }
    float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
    for (int k = 0; k < S; ++k) {
        const T* kr = base + (size_t)k * rowp + C;
        const T* vr = kr + C;
        float sa = 0.f, sb = 0.f, vv[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            const int e = lane + 32 * i;
            const float kv = e < d ? to_f(kr[e]) : 0.f;
            vv[i] = e < d ? to_f(vr[e]) : 0.f;
            sa = fmaf(qa[i], kv, sa);
            sb = fmaf(qb[i], kv, sb);
        }
        sa = warp_sum(sa); sb = warp_sum(sb);
        const float na = fmaxf(ma, sa), nb = fmaxf(mb, sb);
        const float ca = kPrecise ? expf(ma - na) : __expf(ma - na), cb = kPrecise ? expf(mb - nb) : __expf(mb - nb);
        const float pa = kPrecise ? expf(sa - na) : __expf(sa - na), pb = kPrecise ? expf(sb - nb) : __expf(sb - nb);
        la = la * ca + pa; lb = lb * cb + pb;
#pragma unroll
        for (int i = 0; i < DPL; ++i) { oa[i] = fmaf(oa[i], ca, pa * vv[i]); ob[i] = fmaf(ob[i], cb, pb * vv[i]); }
        ma = na; mb = nb;
    }
    const float ia = 1.0f / la, ib = 1.0f / lb;
    T* o = out + ((size_t)n * S + q0) * C + (size_t)head * d;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        const int e = lane + 32 * i;
        if (e < d) {
            o[e] = from_f<T>(oa[i] * ia);
            if (two) o[C + e] = from_f<T>(ob[i] * ib);
        }
    }
}

template <typename T, int DPL>
static void launch_attn_generic_t(bool precise, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s) {
    dim3 grid((S + 15) / 16, C / d, N);
    const float scale = 1.0f / sqrtf((float)d);
    if (precise) attention_generic_kernel<T, true, DPL><<<grid, 256, 0, s>>>((const T*)qkv, S, C, d, scale, (T*)out);
    else attention_generic_kernel<T, false, DPL><<<grid, 256, 0, s>>>((const T*)qkv, S, C, d, scale, (T*)out);
}

int launch_attention_generic(int dt, bool precise, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s) {
    PD_REQUIRE(d > 0 && C % d == 0, "attention channels must be a multiple of the head dimension");
    PD_REQUIRE(d <= 1024, "attention head dimension above 1024 is not implemented");
    PD_DISPATCH_DT(dt, T, {
        if (d <= 32) launch_attn_generic_t<T, 1>(precise, qkv, N, S, C, d, out, s);
        else if (d <= 64) launch_attn_generic_t<T, 2>(precise, qkv, N, S, C, d, out, s);
        else if (d <= 128) launch_attn_generic_t<T, 4>(precise, qkv, N, S, C, d, out, s);
        else if (d <= 256) launch_attn_generic_t<T, 8>(precise, qkv, N, S, C, d, out, s);
        else if (d <= 512) launch_attn_generic_t<T, 16>(precise, qkv, N, S, C, d, out, s);
        else launch_attn_generic_t<T, 32>(precise, qkv, N, S, C, d, out, s);
    });
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_attention_simt(int dt, bool precise, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s) {
    if (d != 8) return launch_attention_generic(dt, precise, qkv, N, S, C, d, out, s);
    PD_REQUIRE(C % 8 == 0, "attention channels must be a multiple of 8");
    dim3 grid((S + 127) / 128, C / d, N);
    if (precise) PD_DISPATCH_DT(dt, T, (attention_simt_kernel<T, true><<<grid, 128, 0, s>>>((const T*)qkv, S, C, (T*)out)));
    else PD_DISPATCH_DT(dt, T, (attention_simt_kernel<T, false><<<grid, 128, 0, s>>>((const T*)qkv, S, C, (T*)out)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// scheduler.add_noise / get_velocity, classifier-free guidance combine, pipeline de-normalisation
// =====================================================================================================================
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ ca,
                             const float* __restrict__ cb, float* __restrict__ out, int64_t per, int64_t total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int n = i / per;
        out[i] = ca[n] * a[i] + cb[n] * b[i];
    }
}
int launch_axpby(const float* a, const float* b, const float* ca, const float* cb, float* out, int B, int64_t per,
                 cudaStream_t s) {
    int64_t total = (int64_t)B * per;
    int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    axpby_kernel<<<grid, 256, 0, s>>>(a, b, ca, cb, out, per, total);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void cfg_kernel(const float* __restrict__ cond, const float* __restrict__ uncond, const float* __restrict__ w,
                           int eqn, float* __restrict__ out, int64_t per, int64_t total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int n = i / per;
        const float c = cond[i], u = uncond[i];
        out[i] = (eqn == 0 ? u : c) + w[n] * (c - u);
    }
}
int launch_cfg(const float* cond, const float* uncond, const float* w, int eqn, float* out, int B, int64_t per,
               cudaStream_t s) {
    int64_t total = (int64_t)B * per;
    int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    cfg_kernel<<<grid, 256, 0, s>>>(cond, uncond, w, eqn, out, per, total);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void denorm_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int HW, int64_t total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into NHWC output
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int c = i % C;
        const int64_t p = i / C;
        const int64_t hw = p % HW, n = p / HW;
        float v = x[(n * C + c) * HW + hw] / 2.f + 0.5f;
        v = (v < 0.f) ? 0.f : ((v > 1.f) ? 1.f : v);   // NaN propagates like torch.clamp
        out[i] = v;
    }
}
int launch_denorm(const float* x, float* out, int B, int C, int H, int W, cudaStream_t s) {
    int64_t total = (int64_t)B * C * H * W;
    int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    denorm_kernel<<<grid, 256, 0, s>>>(x, out, C, H * W, total);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =====================================================================================================================
// weight re-layout (once, at pd_unet_finalize)
// =====================================================================================================================
__global__ void relayout_simt_kernel(const float* __restrict__ w, int O, int I, int k, float* __restrict__ out) {
    const size_t total = (size_t)O * I * k * k;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {   // i indexes the output (tap, ci, o)
        const int o = i % O;
        size_t r = i / O;
        const int ci = r % I;
        const int tap = r / I;
        out[i] = w[((size_t)o * I + ci) * k * k + tap];
    }
}
int launch_relayout_simt(const float* w, int O, int I, int k, float* out, cudaStream_t s) {
    size_t total = (size_t)O * I * k * k;
    int grid = (int)std::min<size_t>((total + 255) / 256, 4096);
    relayout_simt_kernel<<<grid, 256, 0, s>>>(w, O, I, k, out);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
__global__ void relayout_tc_kernel(const float* __restrict__ w, int O, int I, int k, T* __restrict__ out, int ktot,
                                   int koff) {
    const size_t total = (size_t)O * I * k * k;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {   // i indexes (o, tap, ci)
        const int ci = i % I;
        size_t r = i / I;
        const int tap = r % (k * k);
        const int o = r / (k * k);
        out[(size_t)o * ktot + koff + (size_t)tap * I + ci] = from_f<T>(w[((size_t)o * I + ci) * k * k + tap]);
    }
}
int launch_relayout_tc(int dt, const float* w, int O, int I, int k, void* out, int ktot, int koff, cudaStream_t s) {
    size_t total = (size_t)O * I * k * k;
    int grid = (int)std::min<size_t>((total + 255) / 256, 4096);
    PD_DISPATCH_HALF(dt, T, (relayout_tc_kernel<T><<<grid, 256, 0, s>>>(w, O, I, k, (T*)out, ktot, koff)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void relayout_convout_kernel(const float* __restrict__ w, int O, int I, float* __restrict__ out) {
    const int total = 9 * I * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int o = i % 4;
        const int ci = (i / 4) % I;
        const int tap = i / (4 * I);
        out[i] = (o < O) ? w[((size_t)o * I + ci) * 9 + tap] : 0.f;
    }
}
int launch_relayout_convout(const float* w, int O, int I, float* out, cudaStream_t s) {
    relayout_convout_kernel<<<16, 256, 0, s>>>(w, O, I, out);
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Upsample2D (nearest 2x, then 3x3 conv) as four 2x2 sub-pixel phase convolutions: phase (a,b) tap (dr,dc) sums the 3x3
// weights whose upsampled source pixel falls on low-res pixel (i-1+a+dr, j-1+b+dc):  a=0: dr0<-{r0} dr1<-{r1,r2};
// a=1: dr0<-{r0,r1} dr1<-{r2}; same for columns.  out (4*O, 4*I): row = (a*2+b)*O + o, k = (dr*2+dc)*I + i.
template <typename T>
__global__ void relayout_upsample_kernel(const float* __restrict__ w, int O, int I, T* __restrict__ out) {
    const size_t total = (size_t)16 * O * I;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        const int i = idx % I;
        size_t r = idx / I;
        const int tap = r % 4; r /= 4;
        const int o = r % O;
        const int ph = r / O;
        const int a = ph >> 1, b = ph & 1, dr = tap >> 1, dc = tap & 1;
        const int r_lo = (a == 0) ? (dr == 0 ? 0 : 1) : (dr == 0 ? 0 : 2), r_hi = (a == 0) ? (dr == 0 ? 0 : 2) : (dr == 0 ? 1 : 2);
        const int s_lo = (b == 0) ? (dc == 0 ? 0 : 1) : (dc == 0 ? 0 : 2), s_hi = (b == 0) ? (dc == 0 ? 0 : 2) : (dc == 0 ? 1 : 2);
        float acc = 0.f;
        for (int rr = r_lo; rr <= r_hi; ++rr)
            for (int ss = s_lo; ss <= s_hi; ++ss) acc += w[((size_t)o * I + i) * 9 + rr * 3 + ss];
        out[((size_t)ph * O + o) * (4 * (size_t)I) + (size_t)tap * I + i] = from_f<T>(acc);
    }
}
int launch_relayout_upsample(int dt, const float* w, int O, int I, void* out, cudaStream_t s) {
    size_t total = (size_t)16 * O * I;
    int grid = (int)std::min<size_t>((total + 255) / 256, 4096);
    PD_DISPATCH_HALF(dt, T, (relayout_upsample_kernel<T><<<grid, 256, 0, s>>>(w, O, I, (T*)out)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
__global__ void cast_half_kernel(const float* __restrict__ x, T* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = from_f<T>(x[i]);
}
int launch_cast_half(int dt, const float* x, void* out, int64_t n, cudaStream_t s) {
    int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    if (grid < 1) grid = 1;
    PD_DISPATCH_HALF(dt, T, (cast_half_kernel<T><<<grid, 256, 0, s>>>(x, (T*)out, n)));
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pd
