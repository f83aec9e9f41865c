// phendiff_b200 — internal launch interface between the executor (pd_api.cu) and the kernel files.
#pragma once
#include "pd_common.cuh"
#include "../../include/phendiff_b200.h"

namespace pd {

// ---- embeddings (cond_unet_2d.py:289-309; diffusers Timesteps/TimestepEmbedding) -------------------------------
struct EmbedArgs {
    const float* timesteps;       // (B), or null: every sample uses t_scalar
    float t_scalar;
    const int64_t* labels;        // (B) or null
    const float* class_emb;       // (B,D) or null
    const float *w1, *b1;         // (D,C0),(D)
    const float *w2, *b2;         // (D,D),(D)
    const float* class_table;     // (ncls,D) or null
    int B, C0, D, ncls, flip;
    float shift;
    float* emb_act;               // (rows,D) = SiLU(time_embedding + class embedding)
    // The embedding depends only on (t, class): with one scalar t and integer labels (the DDIB path) only `ncls`
    // distinct rows exist.  dedupe != 0: rows = ncls, row_idx[b] = labels[b]; else rows = B, row_idx[b] = b.
    int dedupe;
    int32_t* row_idx;             // (B) out: row of emb_act / of the projected table used by image b
    // classifier-free guidance in ONE pass (pipeline_conditionial_ddim.py:308-317): cfg_pairs = P > 0 (needs dedupe):
    // the pass holds 2P images, image P + i is the UNCONDITIONAL copy of image i (class embedding = zeros): one extra row
    // `ncls` without class embedding, row_idx[i] = labels[i] for i < P and ncls for i >= P.
    int cfg_pairs;
};
// (an unconditional model — no class table, ncls = 0 — has ONE row under dedupe: every image reads row 0)
inline int embed_rows(const EmbedArgs& a) { return a.dedupe ? a.ncls + ((a.cfg_pairs > 0 || a.ncls == 0) ? 1 : 0) : a.B; }

// guidance combine applied by a conv_out epilogue to its conditional output m_c before the scheduler update
// (pipeline_conditionial_ddim.py:323-332): m = (eqn == 0 ? u : m_c) + w[img] * (m_c - u), u = uncond[(img, c, hw)]
struct CfgEpi {
    const float* uncond;          // (P, Cout, H, W) fp32 unconditional model output, or null: no guidance
    const float* w;               // (P) guidance scale per image
    int eqn;                      // 0 "imagen", 1 "CFG"
};
int launch_embed(const EmbedArgs& a, cudaStream_t s);
// all ResnetBlock2D.time_emb_proj at once: out(B,J) = emb_act(B,D) @ wcat(J,D)^T + bcat(J)
int launch_temb_proj(const float* emb_act, const float* wcat, const float* bcat, int B, int D, int J, float* out,
                     cudaStream_t s);

// ---- GroupNorm (+SiLU) over NHWC with two channel-concatenated sources ------------------------------------------
struct GNArgs {
    const void* x1; const void* x2;   // (N,HW,C1) , (N,HW,C2) or null
    int C1, C2, N, HW, groups;
    float eps;
    const float *gamma, *beta;        // (C1+C2)
    int silu;
    // per-tensor "chunk statistics": (N, C/4, 2) fp64 = sum and sum of squares over H*W of every 4-channel chunk
    // (fp32 partial sums over <= 128 elements inside a warp / thread, accumulated across warps and finalised in fp64:
    // a single-pass fp32 E[x^2] - E[x]^2 cancels catastrophically when |mean| >> std).
    // Written by the producing conv's epilogue (tensor-core path) or by launch_gn_chunk_stats; the chunk
    // width cw (4, 2 or 1; model-wide) divides every group width, so any GroupNorm over any concat is finalised from them.
    int stats_cw;
    const double* stats1; const double* stats2;
    void* out;                        // (N,HW,C1+C2)
};
// standalone producer of chunk statistics for one NHWC tensor (stats zero on entry)
int launch_gn_chunk_stats(int dt, const void* x, int N, int HW, int C, int cw, double* stats, cudaStream_t s);
int launch_gn_apply(int dt, bool precise, const GNArgs& a, cudaStream_t s);
// GroupNorm as (scale, shift) per (image, channel) for a consumer that normalises its own input tiles (x1/x2/out/silu unused)
int launch_gn_coef(const GNArgs& a, float2* coef, cudaStream_t s);

// ---- generic SIMT convolution (fp32 validation path; odd shapes of the bf16 path) -------------------------------
struct ConvArgs {
    const void* x1; const void* x2;   // NHWC sources, channel-concatenated
    int C1, C2, N, H, W, Cout, ksize, stride, pad, Ho, Wo;
    const float* w;                   // (k*k*(C1+C2), Cout) fp32, tap-major then input channel
    const float* bias;                // (Cout) or null
    const float* addvec;              // (rows, addvec_stride) or null: per-image per-channel add (time embedding)
    const int32_t* addvec_row;        // (N) row of addvec used by image n, or null (row = n)
    int addvec_stride;
    const void* residual;             // (N,Ho,Wo,Cout) or null
    float out_scale;                  // multiplies the final sum (1/output_scale_factor)
    void* out;                        // (N,Ho,Wo,Cout)
};
int launch_conv_simt(int dt, const ConvArgs& a, cudaStream_t s);

// conv_in: NCHW fp32 sample -> NHWC activations (cond_unet_2d.py:313)
// n_src: image n reads x[n % n_src] (the 2P-image guidance pass feeds every sample twice); 0 = N
int launch_conv_in(int dt, const float* x, const float* w /*(9*Cin,Cout)*/, const float* bias, int N, int Cin, int H,
                   int W, int Cout, void* out, cudaStream_t s, int n_src = 0);
// conv_out (+ optional fused DDIM update): NHWC activations -> NCHW fp32 (cond_unet_2d.py:348, A.5)
struct ConvOutArgs {
    const void* act;                  // (N,H,W,Cin) = SiLU(GN(sample))
    const float* w;                   // (9, Cin, 4) fp32 (out channel padded to 4)
    const float* bias;                // (Cout)
    int N, H, W, Cin, Cout;
    float* model_out;                 // (N,Cout,H,W) or null
    float* x;                         // (N,Cout,H,W) updated in place when `step` != null
    const pd_step_coeffs_t* step;     // host pointer (copied by value into the launch) or null
    // image range [img_begin, img_begin + img_count) of `act` handled by this launch (img_count 0: all N); model_out / x /
    // cfg are indexed by (image - img_begin)
    int img_begin, img_count;
    CfgEpi cfg;
};
int launch_conv_out(int dt, const ConvOutArgs& a, cudaStream_t s);

int launch_upsample2x(int dt, const void* x, int N, int H, int W, int C, void* out, cudaStream_t s);
// conv_in on the tensor cores: gather the 3x3 neighbourhood of the NCHW fp32 sample into (N,H,W,64) 16-bit rows
// (k = tap*Cin + ci, zero padded), which a 1x1 tcgen05 GEMM with the (Cout, 64) re-laid-out weights consumes
int launch_im2col_in(int dt, const float* x, int N, int Cin, int H, int W, void* out, cudaStream_t s, int n_src = 0);

// ---- attention core: softmax(q k^T / sqrt(d)) v on packed qkv (N,S,3C) ------------------------------------------
int launch_attention_simt(int dt, bool precise, const void* qkv, int N, int S, int C, int d, void* out,
                          cudaStream_t s);
// any head dimension (attention_head_dim: null -> one head of dim C); launch_attention_simt forwards d != 8 here
int launch_attention_generic(int dt, bool precise, const void* qkv, int N, int S, int C, int d, void* out, cudaStream_t s);
// bf16 / fp16 tensor-core path.  qfold = the factor already folded into q by the caller: 1 for raw q, PD_ATTN_QFOLD when
// the q rows of the fused qkv weight were pre-multiplied at finalize (scores then leave the MMA in log2 units)
constexpr float PD_ATTN_QFOLD = 0.35355339059327373f * 1.4426950408889634f;   // log2(e) / sqrt(8)
// force_variant: 0 = PHENDIFF_B200_ATTN_KERNEL / default, 2 = chunked warp-level, 3 = head-resident warp-level, 4 / 5 / 6 = tcgen05
// kernels tc / tc2 / tc3
int launch_attention_mma(int dt, const void* qkv, int N, int S, int C, int d, float qfold, void* out, cudaStream_t s,
                         int force_variant = 0);
// tcgen05 / TMEM kernel (pd_attn_tc.cu): q is multiplied by qmul while staging; flags[(n, head, 128-query tile, warp)] = 1 where
// the 16-bit P overflowed and the rows must be recomputed by the exact warp-level pass
size_t attention_tc_smem_bytes(int S);
int attention_mma_launches(int S);
int launch_attention_tc(int dt, const void* qkv, int N, int S, int C, float qmul, void* out, uint8_t* flags, int poly_pairs,
                        cudaStream_t s);
// experimental variant (pd_attn_tc2.cu): same contract
int launch_attention_tc2(int dt, const void* qkv, int N, int S, int C, float qmul, void* out, uint8_t* flags, int poly_pairs,
                        cudaStream_t s);

// third tcgen05 design (pd_attn_tc3.cu): persistent, TMA-staged, three independent softmax warpgroup streams.  q must carry the
// finalize-time fold (scores in log2 units); same flags contract as the other two
bool attention_tc3_supported(int S, int C);
int launch_attention_tc3(int dt, const void* qkv, int N, int S, int C, void* out, uint8_t* flags, int poly_pairs, cudaStream_t s);

// ---- scheduler / pipeline elementwise ---------------------------------------------------------------------------
#ifdef __CUDACC__
// one scheduler update, shared by the standalone step kernel and the conv_out epilogues
__device__ __forceinline__ float ddim_update(const pd_step_coeffs_t& c, float x, float m, float noise, float* x0_out) {
    float x0, e;
    if (c.pred_type == PD_PRED_EPSILON) {
        x0 = (x - c.sqrt_beta * m) / c.sqrt_alpha;   // IEEE: alpha = 0 gives +-inf / NaN exactly as the reference
        e = m;
    } else if (c.pred_type == PD_PRED_SAMPLE) {
        x0 = m;
        e = (x - c.sqrt_alpha * x0) / c.sqrt_beta;
    } else {
        x0 = c.sqrt_alpha * x - c.sqrt_beta * m;
        e = c.sqrt_alpha * m + c.sqrt_beta * x;
    }
    if (c.clip) x0 = (x0 < -c.clip_range) ? -c.clip_range : ((x0 > c.clip_range) ? c.clip_range : x0);  // NaN stays NaN
    if (c.use_clipped_model_output) e = (x - c.sqrt_alpha * x0) / c.sqrt_beta;
    float out = c.sqrt_alpha_next * x0 + c.dir_coef * e;
    if (c.sigma != 0.f) out += c.sigma * noise;
    if (x0_out) *x0_out = x0;
    return out;
}
#endif
int launch_ddim_step(const pd_step_coeffs_t& c, const float* x, const float* m, const float* noise, float* x_out,
                     float* x0_out, int64_t n, cudaStream_t s);
int launch_axpby(const float* a, const float* b, const float* ca, const float* cb, float* out, int B, int64_t per,
                 cudaStream_t s);
int launch_cfg(const float* cond, const float* uncond, const float* w, int eqn, float* out, int B, int64_t per,
               cudaStream_t s);
int launch_denorm(const float* x, float* out, int B, int C, int H, int W, cudaStream_t s);

// ---- weight re-layout ---------------------------------------------------------------------------------------------
// OIHW fp32 (O,I,k,k) -> (k*k*I, O) fp32
int launch_relayout_simt(const float* w, int O, int I, int k, float* out, cudaStream_t s);
// OIHW fp32 -> bf16/fp16 (O, ktot) at column offset koff, K index = tap*I + i
int launch_relayout_tc(int dt, const float* w, int O, int I, int k, void* out, int ktot, int koff, cudaStream_t s);
// conv_out OIHW (O,I,3,3) -> (9, I, 4) fp32
int launch_relayout_convout(const float* w, int O, int I, float* out, cudaStream_t s);
int launch_cast_half(int dt, const float* x, void* out, int64_t n, cudaStream_t s);

// ---- tcgen05 implicit-GEMM convolution (pd_conv_tc.cu: per-tap TMA tiles; pd_conv_halo.cu: halo tiles) -------------
struct ConvTcPlan;  // opaque: tensor maps + launch geometry for one layer at one (N,H,W)
enum { TC_MODE_STD = 0, TC_MODE_DDIM = 1 };
struct ConvTcDesc {
    // main segment: ksize x ksize conv over `x` (N,H,W,C) bf16/fp16 NHWC (already normalised / concatenated)
    int dt;                           // DT_BF16 or DT_F16
    const void* x; int C;             // C = channels of the main segment (both sources together)
    // halo kernel only: the last C2 of the C channels come from x2 (channel concat that is never materialised)
    const void* x2; int C2;
    // halo kernel, 3x3 stride-1 only: the conv input is silu(v * gn_coef[n,c].x + gn_coef[n,c].y) of the raw main-segment
    // values v (GroupNorm + SiLU folded into per-image per-channel coefficients, launch_gn_coef); applied to the halo
    // tiles in shared memory, the normalised tensor is never written to HBM.  (N, C) float2 or null.
    const float2* gn_coef;
    int N, H, W, ksize, stride, pad, Ho, Wo, Cout;
    // upsample != 0: nearest-2x upsample followed by the 3x3 conv (Upsample2D), executed as four 2x2 sub-pixel phase
    // convolutions on the LOW-resolution input (H,W); Ho = 2H, Wo = 2W; wmat = (4*Cout, 4*C) phase weights
    int upsample;
    // optional 1x1 shortcut segment over up to two concatenated sources at the OUTPUT resolution
    const void* sc1; int Csc1;
    const void* sc2; int Csc2;
    const void* wmat;                 // (Cout, Ktot) bf16/fp16, Ktot = k*k*C + Csc1 + Csc2
    const float* bias;                // (Cout) or null
    const float* addvec; int addvec_stride;
    const int32_t* addvec_row;        // (N) or null
    const void* residual;             // (N,Ho,Wo,Cout) or null
    float out_scale;
    void* out;                        // (N,Ho,Wo,Cout)
    double* stats_out;                // (N, Cout/stats_cw, 2) fp64 chunk statistics of the stored output (atomic adds); or null
    int stats_cw;                     // 4 or 2
    int mode;                         // TC_MODE_STD, or TC_MODE_DDIM: conv_out (Cout <= 16, wmat rows padded to 16) whose
                                      // epilogue writes NCHW fp32 model output and/or updates x_t in place
};
struct ConvTcLaunch {                 // per-launch values of a TC_MODE_DDIM plan
    float* model_out;                 // (N,Cout,Ho,Wo) fp32 or null
    float* x_t;                       // (N,Cout,Ho,Wo) fp32 updated in place, or null
    const pd_step_coeffs_t* step;     // host pointer, copied by value; required when x_t != null
    int img_begin = 0, img_count = 0; // image range of the plan's N handled by this launch (0: all); outputs indexed from img_begin
    CfgEpi cfg = {nullptr, nullptr, 0};
};
bool conv_tc_supported(const ConvTcDesc& d, std::string* why);      // per-tap kernel (v1)
bool conv_halo_supported(const ConvTcDesc& d, std::string* why);    // halo kernel (v2)
bool conv_tc_can_emit_stats(const ConvTcDesc& d);                   // v1 only: false when a warp's 32 rows span images
int conv_tc_plan_create(const ConvTcDesc& d, ConvTcPlan** out);     // picks the halo kernel when it supports the shape
void conv_tc_plan_destroy(ConvTcPlan* p);
int conv_tc_launch(const ConvTcPlan* p, cudaStream_t s, const ConvTcLaunch* extra = nullptr);
long long conv_pair_launch_count();   // launches of the 2-CTA (cta_group::2) 1x1 kernel since the library was loaded (tests)
// OIHW fp32 (O,I,3,3) -> (4*O, 4*I) 16-bit sub-pixel phase weights (row = phase*O + o, k = (dr*2+dc)*I + i)
int launch_relayout_upsample(int dt, const float* w, int O, int I, void* out, cudaStream_t s);

}  // namespace pd
