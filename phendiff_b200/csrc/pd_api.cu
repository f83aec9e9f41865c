// phendiff_b200 — C ABI (include/phendiff_b200.h), model graph, static buffer planner and step executor.
//
// The layer graph mirrors what CustomCondUNet2DModel.__init__ builds (reference: src/cond_unet_2d/cond_unet_2d.py:126-242,
// blocks per SURVEY Appendix A.1/A.2); forward order mirrors cond_unet_2d.py:244-362.  Parameter names are the diffusers
// checkpoint keys (Appendix A.7).  The executor records the whole forward once per (micro-batch, H, W) as a flat list of
// kernel launches over a statically planned workspace (no allocation, no host sync on the step path).
#include "pd_model.h"

namespace pd {

static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }

}  // namespace pd

using namespace pd;


namespace pd {

static Param* add_param(pd_unet* m, const std::string& name, std::vector<int64_t> shape) {
    auto p = std::make_unique<Param>();
    p->name = name;
    p->shape = shape;
    p->numel = 1;
    for (auto s : shape) p->numel *= (size_t)s;
    Param* r = p.get();
    m->by_name[name] = r;
    m->params.push_back(std::move(p));
    return r;
}
static void make_conv(pd_unet* m, ConvL& c, const std::string& pfx, int cin, int cout, int k, int stride, int pad) {
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.pad = pad;
    c.w = add_param(m, pfx + ".weight", {cout, cin, k, k});
    c.b = add_param(m, pfx + ".bias", {cout});
}
static void make_gn(pd_unet* m, GNL& g, const std::string& pfx, int C) {
    g.C = C;
    g.g = add_param(m, pfx + ".weight", {C});
    g.b = add_param(m, pfx + ".bias", {C});
}
static void make_res(pd_unet* m, ResL& r, const std::string& pfx, int cin, int cout, float scale) {
    r.cin = cin; r.cout = cout; r.scale = scale;
    make_gn(m, r.n1, pfx + ".norm1", cin);
    make_conv(m, r.c1, pfx + ".conv1", cin, cout, 3, 1, 1);
    r.tw = add_param(m, pfx + ".time_emb_proj.weight", {cout, m->D});
    r.tb = add_param(m, pfx + ".time_emb_proj.bias", {cout});
    make_gn(m, r.n2, pfx + ".norm2", cout);
    make_conv(m, r.c2, pfx + ".conv2", cout, cout, 3, 1, 1);
    r.has_sc = cin != cout;
    if (r.has_sc) make_conv(m, r.sc, pfx + ".conv_shortcut", cin, cout, 1, 1, 0);
    r.temb_off = m->J;
    m->J += cout;
}
static void make_attn(pd_unet* m, AttnL& a, const std::string& pfx, int C, float rescale) {
    a.C = C; a.rescale = rescale;
    make_gn(m, a.gn, pfx + ".group_norm", C);
    const char* names[4] = {"to_q", "to_k", "to_v", "to_out.0"};
    const char* old[4] = {"query", "key", "value", "proj_attn"};   // A.7: deprecated spellings accepted on load
    Param** ws[4] = {&a.qw, &a.kw, &a.vw, &a.ow};
    Param** bs[4] = {&a.qb, &a.kb, &a.vb, &a.ob};
    for (int i = 0; i < 4; ++i) {
        *ws[i] = add_param(m, pfx + "." + names[i] + ".weight", {C, C});
        *bs[i] = add_param(m, pfx + "." + names[i] + ".bias", {C});
        m->alias[pfx + "." + old[i] + ".weight"] = pfx + "." + names[i] + ".weight";
        m->alias[pfx + "." + old[i] + ".bias"] = pfx + "." + names[i] + ".bias";
    }
}

static int build_graph(pd_unet* m) {
    const pd_unet_config_t& c = m->cfg;
    const int nb = c.n_blocks;
    const int* boc = c.block_out_channels;
    m->D = boc[0] * 4;   // cond_unet_2d.py:111
    m->J = 0;
    make_conv(m, m->conv_in, "conv_in", c.in_channels, boc[0], 3, 1, 1);
    m->te_w1 = add_param(m, "time_embedding.linear_1.weight", {m->D, boc[0]});
    m->te_b1 = add_param(m, "time_embedding.linear_1.bias", {m->D});
    m->te_w2 = add_param(m, "time_embedding.linear_2.weight", {m->D, m->D});
    m->te_b2 = add_param(m, "time_embedding.linear_2.bias", {m->D});
    if (c.num_class_embeds > 0) m->cls = add_param(m, "class_embedding.weight", {c.num_class_embeds, m->D});
    // down (cond_unet_2d.py:159-182)
    int out = boc[0];
    m->down.resize(nb);
    for (int i = 0; i < nb; ++i) {
        int cin = out;
        out = boc[i];
        DownB& d = m->down[i];
        d.has_attn = c.down_attn[i] != 0;
        d.has_down = i != nb - 1;
        d.res.resize(c.layers_per_block);
        if (d.has_attn) d.attn.resize(c.layers_per_block);
        std::string pfx = "down_blocks." + std::to_string(i);
        for (int j = 0; j < c.layers_per_block; ++j) {
            make_res(m, d.res[j], pfx + ".resnets." + std::to_string(j), j == 0 ? cin : out, out, 1.f);
            if (d.has_attn) make_attn(m, d.attn[j], pfx + ".attentions." + std::to_string(j), out, 1.f);
        }
        if (d.has_down) make_conv(m, d.down, pfx + ".downsamplers.0.conv", out, out, 3, 2, c.downsample_padding);
    }
    // mid (cond_unet_2d.py:184-197)
    m->mid_has_attn = c.add_attention != 0;
    make_res(m, m->mid_r0, "mid_block.resnets.0", boc[nb - 1], boc[nb - 1], c.mid_block_scale_factor);
    if (m->mid_has_attn) make_attn(m, m->mid_attn, "mid_block.attentions.0", boc[nb - 1], c.mid_block_scale_factor);
    make_res(m, m->mid_r1, "mid_block.resnets.1", boc[nb - 1], boc[nb - 1], c.mid_block_scale_factor);
    // up (cond_unet_2d.py:199-228)
    m->up.resize(nb);
    out = boc[nb - 1];
    for (int i = 0; i < nb; ++i) {
        int prev = out;
        out = boc[nb - 1 - i];
        int cin = boc[nb - 1 - std::min(i + 1, nb - 1)];
        UpB& u = m->up[i];
        u.has_attn = c.up_attn[i] != 0;
        u.has_up = i != nb - 1;
        const int L = c.layers_per_block + 1;
        u.res.resize(L);
        if (u.has_attn) u.attn.resize(L);
        std::string pfx = "up_blocks." + std::to_string(i);
        for (int j = 0; j < L; ++j) {
            int skip = (j == L - 1) ? cin : out;
            int rin = (j == 0) ? prev : out;
            make_res(m, u.res[j], pfx + ".resnets." + std::to_string(j), rin + skip, out, 1.f);
            if (u.has_attn) make_attn(m, u.attn[j], pfx + ".attentions." + std::to_string(j), out, 1.f);
        }
        if (u.has_up) make_conv(m, u.up, pfx + ".upsamplers.0.conv", out, out, 3, 1, 1);
    }
    make_gn(m, m->norm_out, "conv_norm_out", boc[0]);
    make_conv(m, m->conv_out, "conv_out", boc[0], c.out_channels, 3, 1, 1);
    return 0;
}

template <typename T> static int dev_alloc(pd_unet* m, T** p, size_t n) {
    PD_CHECK_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    m->owned.push_back((void*)*p);
    return 0;
}

static int dev_alloc_bytes(pd_unet* m, void** p, size_t bytes) {
    PD_CHECK_CUDA(cudaMalloc(p, bytes));
    m->owned.push_back(*p);
    return 0;
}

static int finalize_conv(pd_unet* m, ConvL& c, cudaStream_t s, bool tc) {
    int rc;
    if ((rc = dev_alloc(m, &c.w_simt, c.w->numel))) return rc;
    if ((rc = launch_relayout_simt(c.w->dev, c.cout, c.cin, c.k, c.w_simt, s))) return rc;
    if (tc && c.cin % 64 == 0 && c.cout % 64 == 0) {
        const int ktot = c.k * c.k * c.cin;
        if ((rc = dev_alloc_bytes(m, &c.w_tc, (size_t)c.cout * ktot * 2))) return rc;
        if ((rc = launch_relayout_tc(m->dt, c.w->dev, c.cout, c.cin, c.k, c.w_tc, ktot, 0, s))) return rc;
    }
    return 0;
}

__global__ void scale_vec_kernel(float* a, float f, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] *= f;
}
__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

static int finalize_res(pd_unet* m, ResL& r, cudaStream_t s) {
    int rc;
    if ((rc = finalize_conv(m, r.c1, s, m->half))) return rc;
    if ((rc = finalize_conv(m, r.c2, s, m->half))) return rc;
    if (r.has_sc && (rc = finalize_conv(m, r.sc, s, m->half))) return rc;
    // time_emb_proj rows of the concatenated projection; conv1.bias folded into the projected vector
    PD_CHECK_CUDA(cudaMemcpyAsync(m->wcat + (size_t)r.temb_off * m->D, r.tw->dev, r.tw->numel * sizeof(float),
                                  cudaMemcpyDeviceToDevice, s));
    add_vec_kernel<<<(r.cout + 255) / 256, 256, 0, s>>>(r.tb->dev, r.c1.b->dev, m->bcat + r.temb_off, r.cout);
    if (r.has_sc) {
        if ((rc = dev_alloc(m, &r.b2sc, (size_t)r.cout))) return rc;
        add_vec_kernel<<<(r.cout + 255) / 256, 256, 0, s>>>(r.c2.b->dev, r.sc.b->dev, r.b2sc, r.cout);
        if (m->half && r.cin % 64 == 0 && r.cout % 64 == 0) {
            const int ktot = 9 * r.cout + r.cin;
            if ((rc = dev_alloc_bytes(m, &r.w2sc_tc, (size_t)r.cout * ktot * 2))) return rc;
            if ((rc = launch_relayout_tc(m->dt, r.c2.w->dev, r.cout, r.cout, 3, r.w2sc_tc, ktot, 0, s))) return rc;
            if ((rc = launch_relayout_tc(m->dt, r.sc.w->dev, r.cout, r.cin, 1, r.w2sc_tc, ktot, 9 * r.cout, s))) return rc;
        }
    }
    PD_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int finalize_attn(pd_unet* m, AttnL& a, cudaStream_t s) {
    int rc;
    const size_t CC = (size_t)a.C * a.C;
    if ((rc = dev_alloc(m, &a.wqkv_raw, 3 * CC))) return rc;
    if ((rc = dev_alloc(m, &a.bqkv, (size_t)3 * a.C))) return rc;
    Param* ws[3] = {a.qw, a.kw, a.vw};
    Param* bs[3] = {a.qb, a.kb, a.vb};
    for (int i = 0; i < 3; ++i) {
        PD_CHECK_CUDA(cudaMemcpyAsync(a.wqkv_raw + i * CC, ws[i]->dev, CC * sizeof(float), cudaMemcpyDeviceToDevice, s));
        PD_CHECK_CUDA(cudaMemcpyAsync(a.bqkv + (size_t)i * a.C, bs[i]->dev, a.C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if ((rc = dev_alloc(m, &a.wqkv_simt, 3 * CC))) return rc;
    if ((rc = launch_relayout_simt(a.wqkv_raw, 3 * a.C, a.C, 1, a.wqkv_simt, s))) return rc;
    if ((rc = dev_alloc(m, &a.wo_simt, CC))) return rc;
    if ((rc = launch_relayout_simt(a.ow->dev, a.C, a.C, 1, a.wo_simt, s))) return rc;
    if (m->half && a.C % 64 == 0) {
        if ((rc = dev_alloc_bytes(m, &a.wqkv_tc, 3 * CC * 2))) return rc;
        if ((rc = launch_cast_half(m->dt, a.wqkv_raw, a.wqkv_tc, (int64_t)(3 * CC), s))) return rc;
        // head_dim-8 tensor-core attention takes q in log2 units: a second copy of the fused qkv projection has
        // log2(e)/sqrt(d) folded into its q rows (weight and bias, in fp32, before the one rounding to 16 bits), so the
        // softmax needs no scale multiply.  wqkv_simt was taken from wqkv_raw above and stays unscaled.
        const int hd = m->cfg.attention_head_dim > 0 ? m->cfg.attention_head_dim : a.C;
        if (hd == 8) {
            if ((rc = dev_alloc(m, &a.bqkv_fold, (size_t)3 * a.C))) return rc;
            PD_CHECK_CUDA(cudaMemcpyAsync(a.bqkv_fold, a.bqkv, (size_t)3 * a.C * sizeof(float), cudaMemcpyDeviceToDevice, s));
            scale_vec_kernel<<<(a.C + 255) / 256, 256, 0, s>>>(a.bqkv_fold, PD_ATTN_QFOLD, (int64_t)a.C);
            scale_vec_kernel<<<(int)((CC + 255) / 256), 256, 0, s>>>(a.wqkv_raw, PD_ATTN_QFOLD, (int64_t)CC);
            if ((rc = dev_alloc_bytes(m, &a.wqkv_tc_fold, 3 * CC * 2))) return rc;
            if ((rc = launch_cast_half(m->dt, a.wqkv_raw, a.wqkv_tc_fold, (int64_t)(3 * CC), s))) return rc;
        }
        if ((rc = dev_alloc_bytes(m, &a.wo_tc, CC * 2))) return rc;
        if ((rc = launch_cast_half(m->dt, a.ow->dev, a.wo_tc, (int64_t)CC, s))) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// recording the forward program
// ---------------------------------------------------------------------------------------------------------------------
struct ConvOpt {
    const float* w_simt = nullptr;   // (k*k*Cin, Cout) fp32
    const void* w_tc = nullptr;      // (Cout, Ktot) 16-bit, or null: no tensor-core weights for this layer
    const float* bias = nullptr;     // bias of the main conv
    const float* bias_fused = nullptr;   // main + shortcut bias (used when the shortcut rides in the same GEMM)
    const float* bias_tc = nullptr;      // bias to use instead of `bias` when the tensor-core path runs (pre-scaled q rows)
    const float* addvec = nullptr;   // projected time-embedding table (rows, J) at this layer's column offset
    int addvec_stride = 0;
    Tensor* residual = nullptr;
    float out_scale = 1.f;
    Tensor *s1 = nullptr, *s2 = nullptr;   // fused 1x1 shortcut sources
    const ConvL* scL = nullptr;
    bool want_stats = true;          // the output feeds a GroupNorm
    // fused GroupNorm + SiLU of concat(x, x2) applied by the conv to its own input tiles (tcgen05 halo kernel, GN variant):
    // the normalised tensor is never written.  Only set after gn_fusable() said yes.
    const GNL* gn = nullptr;
    Tensor* x2 = nullptr;
};

// non-null stand-in for ConvTcDesc::gn_coef while only the shape is being checked (never dereferenced)
static const float2* const kGnCoefProbe = reinterpret_cast<const float2*>(uintptr_t(64));

struct Rec {
    pd_unet* m;
    bool dry;
    int mb;
    size_t esz;   // bytes per activation element
    int rc = 0;
    size_t stats_bump = 0;
    bool last_used_tc = false;   // whether the most recent conv() took the tensor-core path

    Tensor* alloc(int C, int H, int W) {
        auto t = std::make_unique<Tensor>();
        t->C = C; t->H = H; t->W = W; t->refs = 1;
        t->bytes = (size_t)mb * H * W * C * esz;
        t->off = m->arena.alloc(t->bytes);
        Tensor* r = t.get();
        m->tensors.push_back(std::move(t));
        return r;
    }
    void retain(Tensor* t) { t->refs++; }
    void release(Tensor* t) {
        if (--t->refs == 0) m->arena.release(t->off, t->bytes);
    }
    void* ptr(const Tensor* t) const { return (void*)(m->arena.base + t->off); }
    void* raw(size_t off) const { return (void*)(m->arena.base + off); }
    void push(OpFn op, int nlaunch, int cls, double flops = 0.0, const std::string& name = std::string()) {
        if (dry) return;
        m->ops.push_back(Op{op, cls, flops, nlaunch, name});
    }
    // chunk-statistics slot of a tensor: (mb, C/cw, 2) fp64 inside the per-forward zeroed statistics region
    void stats_alloc(Tensor* t) {
        t->stats_off = stats_bump;
        stats_bump += (((size_t)mb * (t->C / m->stats_cw) * 2 * sizeof(double)) + 255) & ~(size_t)255;
    }
    double* stats_ptr(const Tensor* t) const { return (double*)((uint8_t*)raw(m->stats_off) + t->stats_off); }
    // standalone producer of chunk statistics (SIMT-produced tensors, shapes whose epilogue cannot emit them)
    void stats_kernel(Tensor* t) {
        stats_alloc(t);
        if (dry) return;
        const void* x = ptr(t); double* st = stats_ptr(t);
        const int N = mb, HW = t->H * t->W, C = t->C, cw = m->stats_cw, dt = m->dt;
        push([=](const Ctx&, cudaStream_t s) { return launch_gn_chunk_stats(dt, x, N, HW, C, cw, st, s); }, 1, CLS_GN);
    }

    // GroupNorm(+SiLU) of concat(a, b) -> new tensor; group statistics come from the sources' chunk statistics
    Tensor* gn(const GNL& g, Tensor* a, Tensor* b, bool do_silu) {
        const int C = a->C + (b ? b->C : 0);
        Tensor* o = alloc(C, a->H, a->W);
        if (a->stats_off == (size_t)-1 || (b && b->stats_off == (size_t)-1)) { rc = 1; set_error("internal: GroupNorm source without statistics"); return o; }
        GNArgs ga{};
        ga.C1 = a->C; ga.C2 = b ? b->C : 0; ga.N = mb; ga.HW = a->H * a->W; ga.groups = m->cfg.norm_num_groups;
        ga.eps = m->cfg.norm_eps; ga.gamma = g.g->dev; ga.beta = g.b->dev; ga.silu = do_silu; ga.stats_cw = m->stats_cw;
        if (!dry) {
            ga.x1 = ptr(a); ga.x2 = b ? ptr(b) : nullptr; ga.out = ptr(o);
            ga.stats1 = stats_ptr(a); ga.stats2 = b ? stats_ptr(b) : nullptr;
            const int dt = m->dt;
            const bool precise = !m->half;
            push([ga, dt, precise](const Ctx&, cudaStream_t s) { return launch_gn_apply(dt, precise, ga, s); }, 1, CLS_GN, 0.0,
                 "gn_apply C=" + std::to_string(C) + " @" + std::to_string(a->H) + "x" + std::to_string(a->W));
        }
        return o;
    }

    // generic conv of `x` (+ optional fused 1x1 shortcut over (s1|s2)) -> new tensor
    Tensor* conv(const ConvL& L, Tensor* x, const ConvOpt& o_, bool upsample = false) {
        const ConvOpt& c = o_;
        const int Ho = upsample ? 2 * x->H : ((L.stride == 2) ? x->H / 2 : x->H);
        const int Wo = upsample ? 2 * x->W : ((L.stride == 2) ? x->W / 2 : x->W);
        Tensor* o = alloc(L.cout, Ho, Wo);
        const int Cmain = x->C + (c.x2 ? c.x2->C : 0);
        ConvTcDesc d{};
        d.dt = m->dt; d.C = Cmain; d.C2 = c.x2 ? c.x2->C : 0; d.N = mb; d.H = x->H; d.W = x->W; d.ksize = L.k; d.stride = L.stride; d.pad = L.pad;
        d.Ho = Ho; d.Wo = Wo; d.Cout = L.cout; d.upsample = upsample ? 1 : 0; d.mode = TC_MODE_STD;
        d.Csc1 = c.s1 ? c.s1->C : 0; d.Csc2 = c.s2 ? c.s2->C : 0;
        d.stats_cw = m->stats_cw;
        d.gn_coef = c.gn ? kGnCoefProbe : nullptr;   // shape checks only look at null / non-null; the real table is bound below
        const bool want_tc = m->half && m->cfg.conv_impl == 0 && c.w_tc != nullptr;
        bool use_tc = want_tc && (conv_halo_supported(d, nullptr) || (!c.gn && !c.x2 && conv_tc_supported(d, nullptr)));
        last_used_tc = use_tc;
        if (upsample && !use_tc) { rc = 1; set_error("internal: fused upsample conv requested for an unsupported shape"); return o; }
        if ((c.gn || c.x2) && !use_tc) { rc = 1; set_error("internal: fused GroupNorm conv requested for an unsupported shape"); return o; }
        // per-image per-channel (scale, shift) of the fused GroupNorm: one tiny kernel ahead of the conv
        size_t coef_off = 0, coef_bytes = 0;
        if (c.gn) {
            if (x->stats_off == (size_t)-1 || (c.x2 && c.x2->stats_off == (size_t)-1)) { rc = 1; set_error("internal: GroupNorm source without statistics"); return o; }
            coef_bytes = (size_t)mb * Cmain * sizeof(float2);
            coef_off = m->arena.alloc(coef_bytes);
            if (!dry) {
                GNArgs ga{};
                ga.C1 = x->C; ga.C2 = c.x2 ? c.x2->C : 0; ga.N = mb; ga.HW = x->H * x->W; ga.groups = m->cfg.norm_num_groups;
                ga.eps = m->cfg.norm_eps; ga.gamma = c.gn->g->dev; ga.beta = c.gn->b->dev; ga.silu = 1; ga.stats_cw = m->stats_cw;
                ga.stats1 = stats_ptr(x); ga.stats2 = c.x2 ? stats_ptr(c.x2) : nullptr;
                float2* cf = (float2*)raw(coef_off);
                push([ga, cf](const Ctx&, cudaStream_t s) { return launch_gn_coef(ga, cf, s); }, 1, CLS_GN, 0.0,
                     "gn_coef C=" + std::to_string(Cmain) + " @" + std::to_string(x->H) + "x" + std::to_string(x->W));
            }
        }
        if (use_tc) {
            const bool fused_stats = c.want_stats && m->stats_cw >= 2 && (conv_halo_supported(d, nullptr) || conv_tc_can_emit_stats(d));
            if (fused_stats) stats_alloc(o);
            m->tc_layers += dry ? 0 : 1;
            if (!dry) {
                d.x = ptr(x); d.x2 = c.x2 ? ptr(c.x2) : nullptr; d.gn_coef = c.gn ? (const float2*)raw(coef_off) : nullptr;
                d.sc1 = c.s1 ? ptr(c.s1) : nullptr; d.sc2 = c.s2 ? ptr(c.s2) : nullptr;
                d.wmat = c.w_tc; d.bias = c.s1 ? c.bias_fused : (c.bias_tc ? c.bias_tc : c.bias); d.addvec = c.addvec; d.addvec_stride = c.addvec_stride;
                d.addvec_row = c.addvec ? (const int32_t*)raw(m->rowidx_off) : nullptr;
                d.residual = c.residual ? ptr(c.residual) : nullptr; d.out_scale = c.out_scale; d.out = ptr(o);
                d.stats_out = fused_stats ? stats_ptr(o) : nullptr;
                ConvTcPlan* pl = nullptr;
                int r = conv_tc_plan_create(d, &pl);
                if (r) { rc = r; return o; }
                m->tc_plans.push_back(pl);
                const double ktot = upsample ? 4.0 * x->C : (double)(L.k * L.k * Cmain + d.Csc1 + d.Csc2);
                const std::string nm = std::string(c.gn ? "gn+" : "") + std::string(upsample ? "up+conv" : "conv") + std::to_string(L.k) + "x" + std::to_string(L.k) + (L.stride == 2 ? "s2 " : " ") +
                                       std::to_string(Cmain) + (d.Csc1 + d.Csc2 ? "+sc" + std::to_string(d.Csc1 + d.Csc2) : std::string()) + "->" +
                                       std::to_string(L.cout) + " @" + std::to_string(Ho) + "x" + std::to_string(Wo) +
                                       (conv_halo_supported(d, nullptr) ? " halo" : " tap");
                push([pl](const Ctx&, cudaStream_t s) { return conv_tc_launch(pl, s); }, 1, CLS_CONV_TC,
                     2.0 * mb * Ho * Wo * (double)L.cout * ktot, nm);
            }
            if (c.want_stats && !fused_stats) stats_kernel(o);
            if (c.gn) m->arena.release(coef_off, coef_bytes);
            return o;
        }
        // SIMT path; a fused shortcut request is split into (1x1 conv -> tmp) + (conv with residual = tmp)
        m->simt_layers += dry ? 0 : 1;
        const int32_t* rowidx = (!dry && c.addvec) ? (const int32_t*)raw(m->rowidx_off) : nullptr;
        Tensor* tmp = nullptr;
        if (c.s1) {
            tmp = alloc(L.cout, Ho, Wo);
            ConvArgs ca{};
            ca.C1 = c.s1->C; ca.C2 = c.s2 ? c.s2->C : 0; ca.N = mb; ca.H = Ho; ca.W = Wo; ca.Cout = L.cout; ca.ksize = 1; ca.stride = 1;
            ca.pad = 0; ca.Ho = Ho; ca.Wo = Wo; ca.w = c.scL->w_simt; ca.bias = c.scL->b->dev; ca.out_scale = 1.f;
            if (!dry) {
                ca.x1 = ptr(c.s1); ca.x2 = c.s2 ? ptr(c.s2) : nullptr; ca.out = ptr(tmp);
                const int dt = m->dt;
                push([ca, dt](const Ctx&, cudaStream_t s) { return launch_conv_simt(dt, ca, s); }, 1, CLS_CONV_SIMT,
                     2.0 * ca.N * ca.Ho * ca.Wo * (double)ca.Cout * (double)(ca.ksize * ca.ksize * (ca.C1 + ca.C2)));
            }
        }
        ConvArgs ca{};
        ca.C1 = x->C; ca.C2 = 0; ca.N = mb; ca.H = x->H; ca.W = x->W; ca.Cout = L.cout; ca.ksize = L.k; ca.stride = L.stride;
        ca.pad = L.pad; ca.Ho = Ho; ca.Wo = Wo; ca.w = c.w_simt; ca.bias = c.bias; ca.addvec = c.addvec; ca.addvec_stride = c.addvec_stride;
        ca.addvec_row = rowidx; ca.out_scale = c.out_scale;
        if (!dry) {
            ca.x1 = ptr(x); ca.x2 = nullptr; ca.out = ptr(o);
            ca.residual = tmp ? ptr(tmp) : (c.residual ? ptr(c.residual) : nullptr);
            const int dt = m->dt;
            push([ca, dt](const Ctx&, cudaStream_t s) { return launch_conv_simt(dt, ca, s); }, 1, CLS_CONV_SIMT,
                     2.0 * ca.N * ca.Ho * ca.Wo * (double)ca.Cout * (double)(ca.ksize * ca.ksize * (ca.C1 + ca.C2)));
        }
        if (tmp) release(tmp);
        if (c.want_stats) stats_kernel(o);
        return o;
    }

    // ResnetBlock2D (A.1) on concat(a, b); returns the block output (refs = 1)
    // can GroupNorm + SiLU of concat(a, b) ride inside the conv that consumes it (tcgen05 halo kernel, GN variant)?
    bool gn_fusable(const ConvL& L, const void* w_tc, const Tensor* a, const Tensor* b, const Tensor* s1, const Tensor* s2) const {
        static const int on = [] { const char* e = getenv("PHENDIFF_B200_GNFUSE"); return e ? atoi(e) : 1; }();
        if (!on || !m->half || m->cfg.conv_impl != 0 || !w_tc || L.stride != 1) return false;
        ConvTcDesc d{};
        d.dt = m->dt; d.C = a->C + (b ? b->C : 0); d.C2 = b ? b->C : 0; d.N = mb; d.H = a->H; d.W = a->W; d.ksize = L.k; d.stride = 1;
        d.pad = L.pad; d.Ho = a->H; d.Wo = a->W; d.Cout = L.cout; d.mode = TC_MODE_STD; d.stats_cw = m->stats_cw;
        d.Csc1 = s1 ? s1->C : 0; d.Csc2 = s2 ? s2->C : 0; d.gn_coef = kGnCoefProbe;
        return conv_halo_supported(d, nullptr);
    }

    Tensor* resnet(ResL& R, Tensor* a, Tensor* b) {
        // conv1 (+ time embedding row; conv1.bias is folded into that row at finalize)
        ConvOpt o1;
        o1.w_simt = R.c1.w_simt; o1.w_tc = R.c1.w_tc;
        o1.addvec = dry ? nullptr : (const float*)raw(m->temb_off) + R.temb_off; o1.addvec_stride = m->J;
        Tensor* h1;
        if (gn_fusable(R.c1, R.c1.w_tc, a, b, nullptr, nullptr)) {
            o1.gn = &R.n1; o1.x2 = b;
            h1 = conv(R.c1, a, o1);
        } else {
            Tensor* hn = gn(R.n1, a, b, true);
            h1 = conv(R.c1, hn, o1);
            release(hn);
        }
        const void* w2 = R.has_sc ? R.w2sc_tc : R.c2.w_tc;
        const bool fuse2 = gn_fusable(R.c2, w2, h1, nullptr, R.has_sc ? a : nullptr, R.has_sc ? b : nullptr);
        Tensor* h1n = fuse2 ? h1 : gn(R.n2, h1, nullptr, true);
        if (!fuse2) release(h1);
        ConvOpt o2;
        if (fuse2) o2.gn = &R.n2;
        o2.w_simt = R.c2.w_simt; o2.bias = R.c2.b->dev; o2.out_scale = 1.0f / R.scale;
        if (R.has_sc) {
            // out = (conv2(h) + conv_shortcut(x)) / scale: one K-concatenated GEMM on the tensor-core path
            o2.w_tc = R.w2sc_tc; o2.s1 = a; o2.s2 = b; o2.scL = &R.sc; o2.bias_fused = R.b2sc;
        } else {
            o2.w_tc = R.c2.w_tc; o2.residual = a;
        }
        Tensor* o = conv(R.c2, h1n, o2);
        release(h1n);
        return o;
    }

    // Attention (A.2): GN -> fused qkv projection -> softmax(qk^T/sqrt(d))v -> out projection + residual
    Tensor* attention(AttnL& A, Tensor* x) {
        Tensor* xn = gn(A.gn, x, nullptr, false);
        ConvL lq; lq.cin = A.C; lq.cout = 3 * A.C; lq.k = 1; lq.stride = 1; lq.pad = 0;
        ConvOpt oq;
        const int S = x->H * x->W, C = A.C, d = m->cfg.attention_head_dim > 0 ? m->cfg.attention_head_dim : A.C;
        const bool use_mma = m->half && m->cfg.attn_impl == 0 && (S % 64 == 0) && d == 8;
        const bool fold = use_mma && A.wqkv_tc_fold != nullptr;
        oq.w_simt = A.wqkv_simt; oq.w_tc = fold ? A.wqkv_tc_fold : A.wqkv_tc; oq.bias = A.bqkv; oq.want_stats = false;
        if (fold) oq.bias_tc = A.bqkv_fold;
        Tensor* qkv = conv(lq, xn, oq);
        const float qfold = (fold && last_used_tc) ? PD_ATTN_QFOLD : 1.0f;   // the SIMT projection uses the raw weights
        release(xn);
        Tensor* ao = alloc(A.C, x->H, x->W);
        {
            if (!dry) {
                const void* qp = ptr(qkv);
                void* op = ptr(ao);
                const int N = mb;
                const int dt = m->dt;
                const bool precise = !m->half;
                if (use_mma) push([=](const Ctx&, cudaStream_t s) { return launch_attention_mma(dt, qp, N, S, C, d, qfold, op, s); }, attention_mma_launches(S), CLS_ATTN, 4.0 * N * (double)S * S * C, "attention S=" + std::to_string(S) + " C=" + std::to_string(C));
                else push([=](const Ctx&, cudaStream_t s) { return launch_attention_simt(dt, precise, qp, N, S, C, d, op, s); }, 1, CLS_ATTN, 4.0 * N * (double)S * S * C);
            }
        }
        release(qkv);
        ConvL lo; lo.cin = A.C; lo.cout = A.C; lo.k = 1; lo.stride = 1; lo.pad = 0;
        ConvOpt oo;
        oo.w_simt = A.wo_simt; oo.w_tc = A.wo_tc; oo.bias = A.ob->dev; oo.residual = x; oo.out_scale = 1.0f / A.rescale;
        Tensor* o = conv(lo, ao, oo);
        release(ao);
        return o;
    }

    int record() {
        pd_unet* M = m;
        const pd_unet_config_t& c = M->cfg;
        const int H = M->H, W = M->W, C0 = c.block_out_channels[0];
        const int rows_max = std::max(mb, c.num_class_embeds + 1);   // + the unconditional row of the guidance pass
        // fixed small buffers: statistics region (zeroed once per forward), embedding rows, projected table, row index
        M->stats_off = M->arena.alloc(std::max<size_t>(M->stats_bytes, 1024));
        M->emb_off = M->arena.alloc((size_t)rows_max * M->D * sizeof(float));
        M->temb_off = M->arena.alloc((size_t)rows_max * M->J * sizeof(float));
        M->rowidx_off = M->arena.alloc((size_t)mb * sizeof(int32_t));
        // unconditional model output of the guidance pass (P = mb / 2 samples, NCHW fp32): written by conv_out's first launch,
        // read by the epilogue of its second
        M->cfg_u_off = M->arena.alloc((size_t)std::max(1, mb / 2) * c.out_channels * H * W * sizeof(float));
        if (M->tail) {
            // ragged batch: the last micro-batch runs on scratch copies of its inputs, padded to mb images with its first image
            const size_t hw = (size_t)H * W;
            M->tail_x_off = M->arena.alloc((size_t)mb * c.in_channels * hw * sizeof(float));
            M->tail_out_off = M->arena.alloc((size_t)mb * c.out_channels * hw * sizeof(float));
            M->tail_t_off = M->arena.alloc((size_t)mb * sizeof(float));
            M->tail_lab_off = M->arena.alloc((size_t)2 * mb * sizeof(int64_t));
            M->tail_emb_off = M->arena.alloc((size_t)mb * M->D * sizeof(float));
        }
        if (!dry) {
            void* stats = raw(M->stats_off);
            const size_t sb = M->stats_bytes;
            EmbedArgs ea{};
            ea.w1 = M->te_w1->dev; ea.b1 = M->te_b1->dev; ea.w2 = M->te_w2->dev; ea.b2 = M->te_b2->dev;
            ea.class_table = M->cls ? M->cls->dev : nullptr; ea.B = mb; ea.C0 = C0; ea.D = M->D; ea.ncls = c.num_class_embeds;
            ea.flip = c.flip_sin_to_cos; ea.shift = c.freq_shift; ea.emb_act = (float*)raw(M->emb_off);
            ea.row_idx = (int32_t*)raw(M->rowidx_off);
            float* temb = (float*)raw(M->temb_off);
            const float* wcat = M->wcat; const float* bcat = M->bcat;
            const int D = M->D, J = M->J;
            push([=](const Ctx& cx, cudaStream_t s) {
                if (sb) { PD_CHECK_CUDA(cudaMemsetAsync(stats, 0, sb, s)); }
                EmbedArgs e = ea;
                e.timesteps = cx.timesteps; e.t_scalar = cx.t_scalar; e.labels = cx.labels; e.class_emb = cx.class_emb;
                e.cfg_pairs = cx.cfg_pairs;
                // one scalar timestep + integer labels (the DDIB path): only ncls distinct embedding rows exist
                // (or, for an unconditional model, exactly one)
                e.dedupe = (!cx.timesteps && !cx.class_emb && ((cx.labels && e.class_table && e.ncls > 0 && e.ncls <= e.B) || !e.class_table)) ? 1 : 0;
                if (!e.class_table) e.ncls = 0;
                int r = launch_embed(e, s);
                if (r) return r;
                return launch_temb_proj(e.emb_act, wcat, bcat, embed_rows(e), D, J, temb, s);
            }, 3, CLS_EMBED);
        }
        // conv_in (cond_unet_2d.py:313)
        Tensor* x = nullptr;
        bool in_tc = false;
        if (M->half && c.conv_impl == 0 && M->w_in_tc && c.in_channels <= 4) {
            ConvTcDesc d{};
            d.dt = M->dt; d.C = 64; d.N = mb; d.H = H; d.W = W; d.ksize = 1; d.stride = 1; d.pad = 0; d.Ho = H; d.Wo = W; d.Cout = C0;
            d.stats_cw = M->stats_cw;
            in_tc = conv_halo_supported(d, nullptr) || conv_tc_supported(d, nullptr);
        }
        if (in_tc) {
            // im2col (NCHW fp32 -> (N,H,W,64) 16-bit, k = tap*Cin + ci) + one 1x1 tcgen05 GEMM with K = 64
            Tensor* col = alloc(64, H, W);
            if (!dry) {
                void* o = ptr(col);
                const int N = mb, Cin = c.in_channels, dt = M->dt;
                push([=](const Ctx& cx, cudaStream_t s) { return launch_im2col_in(dt, cx.x, N, Cin, H, W, o, s, cx.cfg_pairs); }, 1, CLS_CONV_IN);
            }
            ConvL l1; l1.cin = 64; l1.cout = C0; l1.k = 1; l1.stride = 1; l1.pad = 0;
            ConvOpt oi;
            oi.w_tc = M->w_in_tc; oi.bias = M->conv_in.b->dev;
            x = conv(l1, col, oi);
            release(col);
        } else {
            x = alloc(C0, H, W);
            if (!dry) {
                void* o = ptr(x);
                const float* w = M->w_in; const float* b = M->conv_in.b->dev;
                const int N = mb, Cin = c.in_channels, dt = M->dt;
                push([=](const Ctx& cx, cudaStream_t s) { return launch_conv_in(dt, cx.x, w, b, N, Cin, H, W, C0, o, s, cx.cfg_pairs); }, 1, CLS_CONV_IN,
                     2.0 * N * H * W * 9.0 * Cin * C0);
            }
            stats_kernel(x);
        }
        if (rc) return rc;
        std::vector<Tensor*> skips;
        retain(x); skips.push_back(x);
        // down (cond_unet_2d.py:316-325)
        for (auto& d : M->down) {
            for (size_t j = 0; j < d.res.size(); ++j) {
                Tensor* y = resnet(d.res[j], x, nullptr);
                release(x); x = y;
                if (d.has_attn) { Tensor* z = attention(d.attn[j], x); release(x); x = z; }
                retain(x); skips.push_back(x);
            }
            if (d.has_down) {
                ConvOpt od;
                od.w_simt = d.down.w_simt; od.w_tc = d.down.w_tc; od.bias = d.down.b->dev;
                Tensor* y = conv(d.down, x, od);
                release(x); x = y;
                retain(x); skips.push_back(x);
            }
            if (rc) return rc;
        }
        // mid (cond_unet_2d.py:328)
        { Tensor* y = resnet(M->mid_r0, x, nullptr); release(x); x = y; }
        if (M->mid_has_attn) { Tensor* y = attention(M->mid_attn, x); release(x); x = y; }
        { Tensor* y = resnet(M->mid_r1, x, nullptr); release(x); x = y; }
        // up (cond_unet_2d.py:332-343)
        for (auto& u : M->up) {
            for (size_t j = 0; j < u.res.size(); ++j) {
                Tensor* sk = skips.back(); skips.pop_back();
                Tensor* y = resnet(u.res[j], x, sk);
                release(sk); release(x); x = y;
                if (u.has_attn) { Tensor* z = attention(u.attn[j], x); release(x); x = z; }
            }
            if (u.has_up) {
                ConvOpt ou;
                ou.w_simt = u.up.w_simt; ou.bias = u.up.b->dev;
                ConvTcDesc d{};
                d.dt = M->dt; d.C = x->C; d.N = mb; d.H = x->H; d.W = x->W; d.ksize = 3; d.stride = 1; d.pad = 1;
                d.Ho = 2 * x->H; d.Wo = 2 * x->W; d.Cout = u.up.cout; d.upsample = 1; d.stats_cw = M->stats_cw;
                if (M->half && c.conv_impl == 0 && u.w_up_tc && conv_halo_supported(d, nullptr)) {
                    // Upsample2D as four sub-pixel phase convs on the low-res tensor (no 4x intermediate)
                    ou.w_tc = u.w_up_tc;
                    Tensor* y = conv(u.up, x, ou, true);
                    release(x); x = y;
                } else {
                    Tensor* big = alloc(x->C, x->H * 2, x->W * 2);
                    if (!dry) {
                        const void* ip = ptr(x); void* op = ptr(big);
                        const int N = mb, h = x->H, w = x->W, C = x->C, dt = M->dt;
                        push([=](const Ctx&, cudaStream_t s) { return launch_upsample2x(dt, ip, N, h, w, C, op, s); }, 1, CLS_UPSAMPLE);
                    }
                    release(x);
                    ou.w_tc = u.up.w_tc;
                    Tensor* y = conv(u.up, big, ou);
                    release(big); x = y;
                }
            }
            if (rc) return rc;
        }
        // out (cond_unet_2d.py:346-348) + optional fused scheduler update (A.5)
        bool out_tc = false;
        ConvTcDesc od{};
        od.dt = M->dt; od.C = C0; od.N = mb; od.H = H; od.W = W; od.ksize = 3; od.stride = 1; od.pad = 1; od.Ho = H; od.Wo = W;
        od.Cout = c.out_channels; od.mode = TC_MODE_DDIM;
        if (M->half && c.conv_impl == 0 && M->w_out_tc) out_tc = conv_halo_supported(od, nullptr);
        // conv_norm_out + SiLU inside conv_out's own halo tiles when the tensor-core path runs (same GN variant as the ResNet convs)
        static const int gnfuse = [] { const char* e = getenv("PHENDIFF_B200_GNFUSE"); return e ? atoi(e) : 1; }();
        bool out_gn = false;
        if (out_tc && gnfuse && x->stats_off != (size_t)-1) {
            ConvTcDesc pd_ = od;
            pd_.gn_coef = kGnCoefProbe;
            out_gn = conv_halo_supported(pd_, nullptr);
        }
        Tensor* xn = out_gn ? x : gn(M->norm_out, x, nullptr, true);
        if (!out_gn) release(x);
        size_t coef_off = 0, coef_bytes = 0;
        if (out_gn) {
            coef_bytes = (size_t)mb * C0 * sizeof(float2);
            coef_off = M->arena.alloc(coef_bytes);
            if (!dry) {
                GNArgs ga{};
                ga.C1 = C0; ga.C2 = 0; ga.N = mb; ga.HW = H * W; ga.groups = c.norm_num_groups; ga.eps = c.norm_eps;
                ga.gamma = M->norm_out.g->dev; ga.beta = M->norm_out.b->dev; ga.silu = 1; ga.stats_cw = M->stats_cw; ga.stats1 = stats_ptr(x);
                float2* cf = (float2*)raw(coef_off);
                push([ga, cf](const Ctx&, cudaStream_t s) { return launch_gn_coef(ga, cf, s); }, 1, CLS_GN, 0.0,
                     "gn_coef C=" + std::to_string(C0) + " @" + std::to_string(H) + "x" + std::to_string(W));
            }
        }
        if (out_tc) {
            if (!dry) {
                od.x = ptr(xn); od.wmat = M->w_out_tc; od.bias = M->conv_out.b->dev; od.out_scale = 1.f;
                od.gn_coef = out_gn ? (const float2*)raw(coef_off) : nullptr;
                ConvTcPlan* pl = nullptr;
                int r = conv_tc_plan_create(od, &pl);
                if (r) return r;
                M->tc_plans.push_back(pl);
                M->tc_layers += 1;
                float* cfg_u = (float*)raw(M->cfg_u_off);
                push([pl, cfg_u, M](const Ctx& cx, cudaStream_t s) {
                    ConvTcLaunch ex{cx.model_out, cx.x_update, cx.step};
                    if (cx.cfg_pairs > 0) {
                        // guidance pass: (1) the unconditional half writes its plain output, (2) the conditional half combines
                        // m = base + w (m_c - m_u) in its epilogue and applies the scheduler update to x_t — the guided score
                        // never exists in HBM
                        ConvTcLaunch un{cfg_u, nullptr, nullptr};
                        un.img_begin = cx.cfg_pairs; un.img_count = cx.cfg_pairs;
                        int r = conv_tc_launch(pl, s, &un);
                        if (r) return r;
                        M->launches += 1;
                        ex.img_begin = 0; ex.img_count = cx.cfg_pairs;
                        ex.cfg = CfgEpi{cfg_u, cx.cfg_w, cx.cfg_eqn};
                    }
                    return conv_tc_launch(pl, s, &ex);
                }, 1, CLS_CONV_OUT, 2.0 * mb * H * W * 9.0 * C0 * c.out_channels, out_gn ? "gn+conv_out+ddim" : "conv_out+ddim");
            }
        } else if (!dry) {
            ConvOutArgs oa{};
            oa.act = ptr(xn); oa.w = M->w_out; oa.bias = M->conv_out.b->dev; oa.N = mb; oa.H = H; oa.W = W; oa.Cin = C0;
            oa.Cout = c.out_channels;
            const int dt = M->dt;
            float* cfg_u = (float*)raw(M->cfg_u_off);
            push([=](const Ctx& cx, cudaStream_t s) {
                ConvOutArgs a = oa;
                a.model_out = cx.model_out; a.x = cx.x_update; a.step = cx.step;
                if (cx.cfg_pairs > 0) {   // guidance pass, as on the tensor-core path: unconditional half first, then combine + update
                    ConvOutArgs un = oa;
                    un.model_out = cfg_u; un.img_begin = cx.cfg_pairs; un.img_count = cx.cfg_pairs;
                    int r = launch_conv_out(dt, un, s);
                    if (r) return r;
                    M->launches += 1;
                    a.img_begin = 0; a.img_count = cx.cfg_pairs;
                    a.cfg = CfgEpi{cfg_u, cx.cfg_w, cx.cfg_eqn};
                }
                return launch_conv_out(dt, a, s);
            }, 1, CLS_CONV_OUT, 2.0 * mb * H * W * 9.0 * C0 * c.out_channels);
        }
        if (out_gn) M->arena.release(coef_off, coef_bytes);
        release(xn);
        M->stats_needed = stats_bump;
        if (!dry && stats_bump > M->stats_bytes) { set_error("internal: statistics region too small"); return 1; }
        return rc;
    }
};

// widest statistics chunk (4, 2 or 1 channels) that divides every GroupNorm group width of the graph, concats included
static int pick_stats_cw(pd_unet* m) {
    const int G = m->cfg.norm_num_groups;
    int cw = 4;
    auto see = [&](int C) { while (cw > 1 && ((C / G) % cw != 0)) cw >>= 1; };
    auto res = [&](const ResL& r) { see(r.cin); see(r.cout); };
    for (auto& d : m->down) { for (auto& r : d.res) res(r); for (auto& a : d.attn) see(a.C); }
    res(m->mid_r0); res(m->mid_r1); if (m->mid_has_attn) see(m->mid_attn.C);
    for (auto& u : m->up) { for (auto& r : u.res) res(r); for (auto& a : u.attn) see(a.C); }
    see(m->norm_out.C);
    return cw;
}

static void clear_plan(pd_unet* m) {
    for (auto* p : m->tc_plans) conv_tc_plan_destroy(p);
    m->tc_plans.clear();
    m->ops.clear();
    m->tensors.clear();
    for (auto& chain : m->prof_events) for (auto& e : chain) cudaEventDestroy(e);
    m->prof_events.clear();
    m->prof_every = 0;
    m->bound = false;
    m->tc_layers = m->simt_layers = 0;
}

// a handle is bound to the device that was current at pd_unet_create: its weights, plans and tensor maps live there
static int check_device(const pd_unet* m) {
    int dev = -1;
    PD_CHECK_CUDA(cudaGetDevice(&dev));
    PD_REQUIRE(dev == m->device, "the current CUDA device is not the one this handle was created on");
    return 0;
}

static int run_program(pd_unet* m, const Ctx& c, cudaStream_t s) {
    std::vector<cudaEvent_t>* ev = nullptr;
    if (m->prof_every > 0) {
        if (m->prof_runs % m->prof_every == 0 && (int)m->prof_events.size() < m->prof_max) {
            m->prof_events.emplace_back(m->ops.size() + 1);
            ev = &m->prof_events.back();
            for (auto& e : *ev) PD_CHECK_CUDA(cudaEventCreate(&e));
            PD_CHECK_CUDA(cudaEventRecord((*ev)[0], s));
        }
        m->prof_runs++;
    }
    for (size_t i = 0; i < m->ops.size(); ++i) {
        Op& op = m->ops[i];
        m->launches += op.nlaunch;
        int r = op.fn(c, s);
        if (r) return r;
        if (ev) PD_CHECK_CUDA(cudaEventRecord((*ev)[i + 1], s));
    }
    return 0;
}

// Copies `n` items of `item` bytes from src into a scratch array of `mb` items and fills items n..mb-1 with item 0
// (the padding images of a ragged tail must be VALID inputs: real pixels, labels inside the class table).
static int copy_padded(void* dst, const void* src, size_t item, int n, int mb, cudaStream_t s) {
    PD_CHECK_CUDA(cudaMemcpyAsync(dst, src, item * n, cudaMemcpyDeviceToDevice, s));
    for (int j = n; j < mb; ++j)
        PD_CHECK_CUDA(cudaMemcpyAsync((uint8_t*)dst + item * j, src, item, cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // namespace pd

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

const char* pd_last_error(void) { return g_err.c_str(); }
/* launches of the opt-in 2-CTA 1x1 kernel (PHENDIFF_B200_LIN2CTA=1) since the library was loaded: lets a test prove it ran */
long long pd_debug_pair_kernel_launches(void) { return conv_pair_launch_count(); }

int pd_version(void) { return 100; }

int pd_unet_create(const pd_unet_config_t* cfg, pd_unet_t** out) {
    PD_REQUIRE(cfg && out, "null argument");
    PD_REQUIRE(cfg->n_blocks >= 1 && cfg->n_blocks <= PD_MAX_BLOCKS, "n_blocks out of range");
    PD_REQUIRE(cfg->precision == PD_PREC_FP32 || cfg->precision == PD_PREC_BF16 || cfg->precision == PD_PREC_FP16, "unknown precision");
    PD_REQUIRE(cfg->layers_per_block >= 1, "layers_per_block must be >= 1");
    PD_REQUIRE(cfg->norm_num_groups > 0, "norm_num_groups must be > 0");
    for (int i = 0; i < cfg->n_blocks; ++i) {
        PD_REQUIRE(cfg->block_out_channels[i] % cfg->norm_num_groups == 0, "block_out_channels must be divisible by norm_num_groups");
        PD_REQUIRE(cfg->block_out_channels[i] % 16 == 0, "block_out_channels must be multiples of 16");
    }
    int ndev = 0;
    PD_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    PD_REQUIRE(ndev > 0, "no CUDA device: phendiff_b200 has no CPU fallback");
    pd_unet* m = new pd_unet();
    m->cfg = *cfg;
    m->dt = cfg->precision;   // PD_PREC_* == DT_*
    m->half = m->dt != DT_F32;
    PD_CHECK_CUDA(cudaGetDevice(&m->device));
    cudaDeviceProp prop;
    PD_CHECK_CUDA(cudaGetDeviceProperties(&prop, m->device));
    if (prop.major != 10) {
        delete m;
        set_error("phendiff_b200 kernels are built for sm_100a only; found compute capability " +
                  std::to_string(prop.major) + "." + std::to_string(prop.minor));
        return 1;
    }
    build_graph(m);
    m->stats_cw = pick_stats_cw(m);
    *out = m;
    return 0;
}

int pd_unet_destroy(pd_unet_t* m) {
    if (!m) return 0;
    clear_plan(m);
    for (auto& p : m->params) if (p->dev) cudaFree(p->dev);
    for (void* p : m->owned) cudaFree(p);
    delete m;
    return 0;
}

int pd_unet_num_params(pd_unet_t* m, int32_t* n) {
    PD_REQUIRE(m && n, "null argument");
    *n = (int32_t)m->params.size();
    return 0;
}

int pd_unet_param_info(pd_unet_t* m, int32_t idx, const char** name, int32_t* ndim, int64_t shape[4]) {
    PD_REQUIRE(m && idx >= 0 && idx < (int)m->params.size(), "parameter index out of range");
    Param* p = m->params[idx].get();
    if (name) *name = p->name.c_str();
    if (ndim) *ndim = (int32_t)p->shape.size();
    if (shape) for (size_t i = 0; i < p->shape.size() && i < 4; ++i) shape[i] = p->shape[i];
    return 0;
}

int pd_unet_load_weight(pd_unet_t* m, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
    PD_REQUIRE(m && name && data, "null argument");
    std::string key(name);
    auto al = m->alias.find(key);
    if (al != m->alias.end()) key = al->second;
    auto it = m->by_name.find(key);
    PD_REQUIRE(it != m->by_name.end(), (std::string("unknown parameter name: ") + name).c_str());
    Param* p = it->second;
    size_t numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= (size_t)shape[i];
    // 1x1 conv weights may arrive as (O,I) linear weights and vice versa (deprecated attention blocks): sizes must match
    PD_REQUIRE(numel == p->numel, (std::string("shape mismatch for ") + name).c_str());
    if (!p->dev) PD_CHECK_CUDA(cudaMalloc((void**)&p->dev, p->numel * sizeof(float)));
    PD_CHECK_CUDA(cudaMemcpy(p->dev, data, p->numel * sizeof(float), cudaMemcpyDefault));
    p->loaded = true;
    m->finalized = false;
    return 0;
}

int pd_unet_finalize(pd_unet_t* m, pd_stream_t stream) {
    PD_REQUIRE(m, "null handle");
    cudaStream_t s = (cudaStream_t)stream;
    for (auto& p : m->params) PD_REQUIRE(p->loaded, (std::string("parameter never loaded: ") + p->name).c_str());
    clear_plan(m);
    for (void* p : m->owned) cudaFree(p);
    m->owned.clear();
    int rc;
    if ((rc = dev_alloc(m, &m->wcat, (size_t)m->J * m->D))) return rc;
    if ((rc = dev_alloc(m, &m->bcat, (size_t)m->J))) return rc;
    if ((rc = dev_alloc(m, &m->w_in, m->conv_in.w->numel))) return rc;
    if ((rc = launch_relayout_simt(m->conv_in.w->dev, m->conv_in.cout, m->conv_in.cin, 3, m->w_in, s))) return rc;
    if ((rc = dev_alloc(m, &m->w_out, (size_t)9 * m->conv_out.cin * 4))) return rc;
    if ((rc = launch_relayout_convout(m->conv_out.w->dev, m->conv_out.cout, m->conv_out.cin, m->w_out, s))) return rc;
    m->w_in_tc = m->w_out_tc = nullptr;
    if (m->half && m->conv_in.cin <= 4 && m->conv_in.cout % 64 == 0) {
        // conv_in as one K = 64 GEMM: (C0, 64) rows, k = tap*Cin + ci, zero beyond 9*Cin
        if ((rc = dev_alloc_bytes(m, &m->w_in_tc, (size_t)m->conv_in.cout * 64 * 2))) return rc;
        PD_CHECK_CUDA(cudaMemsetAsync(m->w_in_tc, 0, (size_t)m->conv_in.cout * 64 * 2, s));
        if ((rc = launch_relayout_tc(m->dt, m->conv_in.w->dev, m->conv_in.cout, m->conv_in.cin, 3, m->w_in_tc, 64, 0, s))) return rc;
    }
    if (m->half && m->conv_out.cout <= 16 && m->conv_out.cin % 64 == 0) {
        const int ktot = 9 * m->conv_out.cin;
        if ((rc = dev_alloc_bytes(m, &m->w_out_tc, (size_t)16 * ktot * 2))) return rc;
        PD_CHECK_CUDA(cudaMemsetAsync(m->w_out_tc, 0, (size_t)16 * ktot * 2, s));
        if ((rc = launch_relayout_tc(m->dt, m->conv_out.w->dev, m->conv_out.cout, m->conv_out.cin, 3, m->w_out_tc, ktot, 0, s))) return rc;
    }
    for (auto& d : m->down) {
        for (auto& r : d.res) if ((rc = finalize_res(m, r, s))) return rc;
        for (auto& a : d.attn) if ((rc = finalize_attn(m, a, s))) return rc;
        if (d.has_down && (rc = finalize_conv(m, d.down, s, m->half))) return rc;
    }
    if ((rc = finalize_res(m, m->mid_r0, s))) return rc;
    if ((rc = finalize_res(m, m->mid_r1, s))) return rc;
    if (m->mid_has_attn && (rc = finalize_attn(m, m->mid_attn, s))) return rc;
    for (auto& u : m->up) {
        for (auto& r : u.res) if ((rc = finalize_res(m, r, s))) return rc;
        for (auto& a : u.attn) if ((rc = finalize_attn(m, a, s))) return rc;
        if (u.has_up && (rc = finalize_conv(m, u.up, s, m->half))) return rc;
        u.w_up_tc = nullptr;
        if (u.has_up && m->half && u.up.cin % 64 == 0 && u.up.cout % 64 == 0) {
            if ((rc = dev_alloc_bytes(m, &u.w_up_tc, (size_t)16 * u.up.cout * u.up.cin * 2))) return rc;
            if ((rc = launch_relayout_upsample(m->dt, u.up.w->dev, u.up.cout, u.up.cin, u.w_up_tc, s))) return rc;
        }
    }
    PD_CHECK_CUDA(cudaStreamSynchronize(s));
    m->finalized = true;
    return 0;
}

int pd_unet_time_embed_dim(pd_unet_t* m, int32_t* dim) {
    PD_REQUIRE(m && dim, "null argument");
    *dim = m->D;
    return 0;
}

static int plan_impl(pd_unet_t* m, int32_t batch, int32_t height, int32_t width, size_t* workspace_bytes, bool pairs);

int pd_unet_plan(pd_unet_t* m, int32_t batch, int32_t height, int32_t width, size_t* workspace_bytes) {
    return plan_impl(m, batch, height, width, workspace_bytes, false);
}

int pd_unet_plan_guided(pd_unet_t* m, int32_t batch, int32_t height, int32_t width, size_t* workspace_bytes) {
    PD_REQUIRE(batch > 0, "bad shape");
    return plan_impl(m, 2 * batch, height, width, workspace_bytes, true);
}

static int plan_impl(pd_unet_t* m, int32_t batch, int32_t height, int32_t width, size_t* workspace_bytes, bool pairs) {
    PD_REQUIRE(m && workspace_bytes, "null argument");
    PD_REQUIRE(m->finalized, "pd_unet_finalize must be called before planning");
    PD_REQUIRE(batch > 0 && height > 0 && width > 0, "bad shape");
    const int ds = 1 << (m->cfg.n_blocks - 1);
    PD_REQUIRE(height % ds == 0 && width % ds == 0, "sample size must be a multiple of 2^(n_blocks-1)");
    clear_plan(m);
    m->B = batch; m->H = height; m->W = width;
    int cap = m->cfg.max_microbatch > 0 ? m->cfg.max_microbatch : 64;   // 64 images: >= 6.9 waves of 148 CTAs at every UNet level
    const int lim = std::min(cap, batch);
    int mb = 1;
    // guided plans (pairs): a pass holds P conditional samples and their P unconditional copies, so the micro-batch is even
    for (int d = 1; d <= lim; ++d) if (batch % d == 0 && (!pairs || d % 2 == 0)) mb = d;
    if (mb * 2 < lim || (pairs && mb % 2)) {
        // no divisor of the batch within a factor 2 of the cap (prime batches: 1): even micro-batches and a padded ragged tail
        const int k = (batch + lim - 1) / lim;
        mb = (batch + k - 1) / k;
        if (pairs && mb % 2) ++mb;
    }
    m->mb = mb;
    m->pairs = pairs;
    m->tail = batch % mb;
    // dry pass 1 sizes the statistics region, dry pass 2 gives the arena peak with that region in place
    m->stats_bytes = 0;
    int rc = 0;
    for (int pass = 0; pass < 2 && !rc; ++pass) {
        m->arena.reset(true, nullptr);
        m->tensors.clear();
        Rec r{m, true, mb, m->half ? (size_t)2 : sizeof(float)};
        rc = r.record();
        m->stats_bytes = m->stats_needed;
    }
    if (rc) return rc;
    m->ws_bytes = m->arena.peak + 1024;
    m->tensors.clear();
    *workspace_bytes = m->ws_bytes;
    return 0;
}

int pd_unet_bind_workspace(pd_unet_t* m, void* workspace, size_t bytes) {
    PD_REQUIRE(m && workspace, "null argument");
    PD_REQUIRE(m->B > 0, "pd_unet_plan must be called first");
    PD_REQUIRE(bytes >= m->ws_bytes, "workspace too small");
    PD_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
    for (auto* p : m->tc_plans) conv_tc_plan_destroy(p);
    m->tc_plans.clear(); m->ops.clear(); m->tensors.clear();
    m->tc_layers = m->simt_layers = 0;
    m->arena.reset(false, (uint8_t*)workspace);
    Rec r{m, false, m->mb, m->half ? (size_t)2 : sizeof(float)};
    int rc = r.record();
    if (rc) return rc;
    m->tensors.clear();
    m->bound = true;
    return 0;
}

int pd_unet_forward(pd_unet_t* m, const float* sample, const float* timesteps, const int64_t* class_labels,
                    const float* class_emb, float* out, pd_stream_t stream) {
    PD_REQUIRE(m && sample && timesteps && out, "null argument");
    PD_REQUIRE(m->bound, "pd_unet_plan + pd_unet_bind_workspace must be called first");
    PD_REQUIRE(!m->pairs, "the current plan is a guided plan (pd_unet_plan_guided): re-plan with pd_unet_plan");
    PD_REQUIRE(!(class_labels && class_emb), "Cannot specify both class_labels and class_emb");
    PD_REQUIRE(!(m->cls && !class_labels && !class_emb), "either class_labels or class_emb should be provided when doing class conditioning");
    { int rc = check_device(m); if (rc) return rc; }
    const size_t per_in = (size_t)m->cfg.in_channels * m->H * m->W, per_out = (size_t)m->cfg.out_channels * m->H * m->W;
    cudaStream_t s = (cudaStream_t)stream;
    for (int i = 0; i < m->B; i += m->mb) {
        Ctx c;
        c.x = sample + (size_t)i * per_in;
        c.timesteps = timesteps + i;
        c.labels = class_labels ? class_labels + i : nullptr;
        c.class_emb = class_emb ? class_emb + (size_t)i * m->D : nullptr;
        c.model_out = out + (size_t)i * per_out;
        const int n = std::min(m->mb, m->B - i);
        if (n < m->mb) {   // ragged tail on padded scratch copies
            uint8_t* base = m->arena.base;
            float* sx = (float*)(base + m->tail_x_off);
            float* so = (float*)(base + m->tail_out_off);
            int rc = copy_padded(sx, c.x, per_in * sizeof(float), n, m->mb, s);
            if (rc) return rc;
            if ((rc = copy_padded(base + m->tail_t_off, c.timesteps, sizeof(float), n, m->mb, s))) return rc;
            if (c.labels && (rc = copy_padded(base + m->tail_lab_off, c.labels, sizeof(int64_t), n, m->mb, s))) return rc;
            if (c.class_emb && (rc = copy_padded(base + m->tail_emb_off, c.class_emb, (size_t)m->D * sizeof(float), n, m->mb, s))) return rc;
            Ctx t = c;
            t.x = sx;
            t.timesteps = (const float*)(base + m->tail_t_off);
            if (c.labels) t.labels = (const int64_t*)(base + m->tail_lab_off);
            if (c.class_emb) t.class_emb = (const float*)(base + m->tail_emb_off);
            t.model_out = so;
            if ((rc = run_program(m, t, s))) return rc;
            PD_CHECK_CUDA(cudaMemcpyAsync(c.model_out, so, per_out * sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
            continue;
        }
        int rc = run_program(m, c, s);
        if (rc) return rc;
    }
    return 0;
}

int pd_ddim_step(const pd_step_coeffs_t* c, const float* x, const float* model_output, const float* noise, float* x_out,
                 float* x0_out, int64_t n, pd_stream_t stream) {
    PD_REQUIRE(c && x && model_output, "null argument");
    PD_REQUIRE(c->sigma == 0.f || noise, "eta > 0 needs a noise tensor");
    return launch_ddim_step(*c, x, model_output, noise, x_out, x0_out, n, (cudaStream_t)stream);
}

int pd_axpby_per_sample(const float* a, const float* b, const float* ca, const float* cb, float* out, int32_t batch,
                        int64_t per_sample, pd_stream_t stream) {
    PD_REQUIRE(a && b && ca && cb && out, "null argument");
    return launch_axpby(a, b, ca, cb, out, batch, per_sample, (cudaStream_t)stream);
}

int pd_cfg_combine(const float* cond, const float* uncond, const float* w, int32_t eqn, float* out, int32_t batch,
                   int64_t per_sample, pd_stream_t stream) {
    PD_REQUIRE(cond && uncond && w && out, "null argument");
    PD_REQUIRE(eqn == 0 || eqn == 1, "Unknown guidance equation; should be 'imagen' (0) or 'CFG' (1)");
    return launch_cfg(cond, uncond, w, eqn, out, batch, per_sample, (cudaStream_t)stream);
}

int pd_denorm_nhwc(const float* x, float* out, int32_t batch, int32_t channels, int32_t height, int32_t width,
                   pd_stream_t stream) {
    PD_REQUIRE(x && out, "null argument");
    return launch_denorm(x, out, batch, channels, height, width, (cudaStream_t)stream);
}

int pd_ddib_transfer(pd_unet_t* m, float* x, const int64_t* src_labels, const int64_t* tgt_labels,
                     const pd_step_coeffs_t* steps_host, int32_t n_inv, int32_t n_gen, pd_stream_t stream) {
    PD_REQUIRE(m && x && steps_host, "null argument");
    PD_REQUIRE(m->bound, "pd_unet_plan + pd_unet_bind_workspace must be called first");
    PD_REQUIRE(!m->pairs, "the current plan is a guided plan (pd_unet_plan_guided): re-plan with pd_unet_plan");
    PD_REQUIRE(m->cfg.in_channels == m->cfg.out_channels, "in/out channels must match for sampling");
    PD_REQUIRE(!m->cls || ((n_inv == 0 || src_labels) && (n_gen == 0 || tgt_labels)), "class-conditioned model needs source labels for inversion steps and target labels for generation steps");
    const size_t per = (size_t)m->cfg.in_channels * m->H * m->W;
    cudaStream_t s = (cudaStream_t)stream;
    { int rc = check_device(m); if (rc) return rc; }
    for (int i = 0; i < m->B; i += m->mb) {
        const int n = std::min(m->mb, m->B - i);
        float* xi = x + (size_t)i * per;
        const int64_t* src_i = src_labels ? src_labels + i : nullptr;
        const int64_t* tgt_i = tgt_labels ? tgt_labels + i : nullptr;
        if (n < m->mb) {   // ragged tail: the whole trajectory runs on a padded scratch copy, the n real images are copied back
            uint8_t* base = m->arena.base;
            int64_t* sl = (int64_t*)(base + m->tail_lab_off);
            int rc = copy_padded(base + m->tail_x_off, xi, per * sizeof(float), n, m->mb, s);
            if (rc) return rc;
            if (src_i && n_inv > 0) { if ((rc = copy_padded(sl, src_i, sizeof(int64_t), n, m->mb, s))) return rc; src_i = sl; }
            if (tgt_i && n_gen > 0) { if ((rc = copy_padded(sl + m->mb, tgt_i, sizeof(int64_t), n, m->mb, s))) return rc; tgt_i = sl + m->mb; }
            xi = (float*)(base + m->tail_x_off);
        }
        for (int sidx = 0; sidx < n_inv + n_gen; ++sidx) {
            Ctx c;
            c.x = xi;
            c.timesteps = nullptr;
            c.t_scalar = steps_host[sidx].timestep;
            c.labels = sidx < n_inv ? src_i : tgt_i;
            c.model_out = nullptr;
            c.x_update = xi;
            c.step = &steps_host[sidx];
            int rc = run_program(m, c, s);
            if (rc) return rc;
        }
        if (n < m->mb)
            PD_CHECK_CUDA(cudaMemcpyAsync(x + (size_t)i * per, xi, per * sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int pd_cfg_transfer(pd_unet_t* m, float* x, const int64_t* labels, const float* w, int32_t eqn,
                    const pd_step_coeffs_t* steps_host, int32_t n_steps, pd_stream_t stream) {
    PD_REQUIRE(m && x && labels && w && steps_host, "null argument");
    PD_REQUIRE(m->bound && m->pairs, "pd_unet_plan_guided + pd_unet_bind_workspace must be called first");
    PD_REQUIRE(m->cls, "classifier-free guidance needs a class-conditioned model");
    PD_REQUIRE(m->cfg.in_channels == m->cfg.out_channels, "in/out channels must match for sampling");
    PD_REQUIRE(eqn == 0 || eqn == 1, "Unknown guidance equation; should be 'imagen' (0) or 'CFG' (1)");
    { int rc = check_device(m); if (rc) return rc; }
    const size_t per = (size_t)m->cfg.in_channels * m->H * m->W;
    cudaStream_t s = (cudaStream_t)stream;
    const int P = m->mb / 2, samples = m->B / 2;
    for (int i = 0; i < samples; i += P) {
        const int n = std::min(P, samples - i);
        float* xi = x + (size_t)i * per;
        const int64_t* lab_i = labels + i;
        const float* w_i = w + i;
        if (n < P) {   // ragged tail: the trajectory runs on padded scratch copies, the n real samples are copied back
            uint8_t* base = m->arena.base;
            int rc = copy_padded(base + m->tail_x_off, xi, per * sizeof(float), n, P, s);
            if (rc) return rc;
            if ((rc = copy_padded(base + m->tail_lab_off, lab_i, sizeof(int64_t), n, P, s))) return rc;
            if ((rc = copy_padded(base + m->tail_t_off, w_i, sizeof(float), n, P, s))) return rc;
            xi = (float*)(base + m->tail_x_off);
            lab_i = (const int64_t*)(base + m->tail_lab_off);
            w_i = (const float*)(base + m->tail_t_off);
        }
        for (int sidx = 0; sidx < n_steps; ++sidx) {
            Ctx c;
            c.x = xi;
            c.timesteps = nullptr;
            c.t_scalar = steps_host[sidx].timestep;
            c.labels = lab_i;
            c.model_out = nullptr;
            c.x_update = xi;
            c.step = &steps_host[sidx];
            c.cfg_pairs = P; c.cfg_w = w_i; c.cfg_eqn = eqn;
            int rc = run_program(m, c, s);
            if (rc) return rc;
        }
        if (n < P)
            PD_CHECK_CUDA(cudaMemcpyAsync(x + (size_t)i * per, xi, per * sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int pd_unet_profile_begin(pd_unet_t* m, int32_t every_n, int32_t max_samples) {
    PD_REQUIRE(m && every_n > 0 && max_samples > 0, "bad argument");
    PD_REQUIRE(m->prof_events.empty(), "profile already running: call pd_unet_profile_end first");
    m->prof_every = every_n; m->prof_max = max_samples; m->prof_runs = 0;
    for (auto& op : m->ops) { op.ms = 0; op.samples = 0; }
    for (int i = 0; i < CLS_COUNT; ++i) { m->prof_ms[i] = 0; m->prof_flops[i] = 0; m->prof_launches[i] = 0; }
    return 0;
}

int pd_unet_profile_end(pd_unet_t* m, int32_t* samples) {
    PD_REQUIRE(m, "null handle");
    m->prof_every = 0;
    int rc = 0;
    for (auto& chain : m->prof_events) {
        if (chain.size() != m->ops.size() + 1) continue;
        cudaError_t e = cudaEventSynchronize(chain.back());
        if (e != cudaSuccess) { set_error(std::string("profile: ") + cudaGetErrorString(e)); rc = 2; break; }
        for (size_t i = 0; i < m->ops.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, chain[i], chain[i + 1]);
            Op& op = m->ops[i];
            op.ms += ms; op.samples += 1;
            m->prof_ms[op.cls] += ms; m->prof_flops[op.cls] += op.flops; m->prof_launches[op.cls] += op.nlaunch;
        }
    }
    if (samples) *samples = (int32_t)m->prof_events.size();
    for (auto& chain : m->prof_events) for (auto& e : chain) cudaEventDestroy(e);
    m->prof_events.clear();
    return rc;
}

int pd_unet_profile_query(pd_unet_t* m, int32_t cls, double* ms, int64_t* launches, double* flops) {
    PD_REQUIRE(m && cls >= 0 && cls < CLS_COUNT, "bad class id");
    if (ms) *ms = m->prof_ms[cls];
    if (launches) *launches = m->prof_launches[cls];
    if (flops) *flops = m->prof_flops[cls];
    return 0;
}

int pd_unet_profile_op(pd_unet_t* m, int32_t idx, const char** name, int32_t* cls, double* ms, int32_t* samples, double* flops) {
    PD_REQUIRE(m && idx >= 0 && idx < (int)m->ops.size(), "op index out of range");
    const Op& op = m->ops[idx];
    if (name) *name = op.name.c_str();
    if (cls) *cls = op.cls;
    if (ms) *ms = op.ms;
    if (samples) *samples = op.samples;
    if (flops) *flops = op.flops;
    return 0;
}

int pd_unet_plan_info(pd_unet_t* m, int32_t* microbatch, int32_t* tc_layers, int32_t* simt_layers, int32_t* ops) {
    PD_REQUIRE(m && m->bound, "no bound plan");
    if (microbatch) *microbatch = m->mb;
    if (tc_layers) *tc_layers = m->tc_layers;
    if (simt_layers) *simt_layers = m->simt_layers;
    if (ops) *ops = (int32_t)m->ops.size();
    return 0;
}

int pd_unet_launch_count(pd_unet_t* m, int64_t* n) {
    PD_REQUIRE(m && n, "null argument");
    *n = m->launches;
    return 0;
}

// ---- kernel-level test entry points ----------------------------------------------------------------------------------
int pd_test_conv_ex(const pd_test_conv_args_t* a, pd_stream_t stream) {
    PD_REQUIRE(a, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int dt = a->dtype, n = a->n, h = a->h, w = a->w, c1 = a->c1, c2 = a->c2, cout = a->cout, ksize = a->ksize;
    const int stride = a->stride, pad = a->pad, csc1 = a->csc1, csc2 = a->csc2;
    // stride-2 convs follow Downsample2D: output H/2 x W/2 (pad 1, or pad 0 with the implicit (0,1,0,1) zero pad)
    int ho = stride == 2 ? h / 2 : (h + 2 * pad - ksize) / stride + 1;
    int wo = stride == 2 ? w / 2 : (w + 2 * pad - ksize) / stride + 1;
    if (a->upsample) { ho = 2 * h; wo = 2 * w; }
    const int ct = c1 + c2;
    const int cw = a->stats_cw ? a->stats_cw : 4;
    int rc = 0;
    if (a->impl == 1 || a->impl == 2) {
        PD_REQUIRE(dt == DT_BF16 || dt == DT_F16, "tcgen05 path takes bf16 / fp16 activations");
        PD_REQUIRE(c2 == 0, "tcgen05 main segment takes one (already concatenated) source");
        const bool ddim = a->mode == TC_MODE_DDIM;
        const int rows = ddim ? 16 : (a->upsample ? 4 * cout : cout);
        const int ktot = a->upsample ? 4 * ct : ksize * ksize * ct + csc1 + csc2;
        void* wm = nullptr;
        PD_CHECK_CUDA(cudaMalloc(&wm, (size_t)rows * ktot * 2));
        PD_CHECK_CUDA(cudaMemsetAsync(wm, 0, (size_t)rows * ktot * 2, s));
        if (a->upsample) rc = launch_relayout_upsample(dt, a->weight, cout, ct, wm, s);
        else rc = launch_relayout_tc(dt, a->weight, cout, ct, ksize, wm, ktot, 0, s);
        if (!rc && (csc1 + csc2)) rc = launch_relayout_tc(dt, a->sc_w, cout, csc1 + csc2, 1, wm, ktot, ksize * ksize * ct, s);
        ConvTcDesc d{};
        d.dt = dt; d.x = a->x1; d.C = ct; d.N = n; d.H = h; d.W = w; d.ksize = ksize; d.stride = stride; d.pad = pad;
        d.Ho = ho; d.Wo = wo; d.Cout = cout; d.upsample = a->upsample; d.sc1 = a->sc1; d.Csc1 = csc1; d.sc2 = a->sc2; d.Csc2 = csc2;
        d.wmat = wm; d.bias = a->bias; d.addvec = a->addvec; d.addvec_stride = cout; d.addvec_row = a->addvec_row;
        d.residual = a->residual; d.out_scale = a->out_scale; d.out = a->out; d.stats_out = a->stats_out; d.stats_cw = cw;
        d.mode = a->mode;
        ConvTapPlan* tp = nullptr;
        ConvHaloPlan* hp = nullptr;
        if (!rc) rc = a->impl == 2 ? conv_halo_plan_create(d, &hp) : conv_tap_plan_create(d, &tp);
        ConvTcLaunch ex{a->model_out, a->x_t, a->step};
        if (!rc) rc = a->impl == 2 ? conv_halo_launch(hp, s, ddim ? &ex : nullptr) : conv_tap_launch(tp, s);
        cudaError_t e = cudaStreamSynchronize(s);
        if (tp) conv_tap_plan_destroy(tp);
        if (hp) conv_halo_plan_destroy(hp);
        cudaFree(wm);
        if (!rc && e != cudaSuccess) { set_error(std::string("conv_tc kernel failed: ") + cudaGetErrorString(e)); rc = 2; }
        return rc;
    }
    PD_REQUIRE(!a->upsample && a->mode == TC_MODE_STD, "the SIMT test path has no upsample / conv_out mode");
    float* wm = nullptr;
    PD_CHECK_CUDA(cudaMalloc((void**)&wm, (size_t)cout * ct * ksize * ksize * sizeof(float)));
    rc = launch_relayout_simt(a->weight, cout, ct, ksize, wm, s);
    void* tmp = nullptr;
    float* wsc = nullptr;
    const size_t esz = dt ? 2 : 4;
    if (!rc && (csc1 + csc2)) {
        PD_CHECK_CUDA(cudaMalloc(&tmp, (size_t)n * ho * wo * cout * esz));
        PD_CHECK_CUDA(cudaMalloc((void**)&wsc, (size_t)cout * (csc1 + csc2) * sizeof(float)));
        rc = launch_relayout_simt(a->sc_w, cout, csc1 + csc2, 1, wsc, s);
        ConvArgs ca{};
        ca.x1 = a->sc1; ca.x2 = a->sc2; ca.C1 = csc1; ca.C2 = csc2; ca.N = n; ca.H = ho; ca.W = wo; ca.Cout = cout; ca.ksize = 1;
        ca.stride = 1; ca.pad = 0; ca.Ho = ho; ca.Wo = wo; ca.w = wsc; ca.out_scale = 1.f; ca.out = tmp;
        if (!rc) rc = launch_conv_simt(dt, ca, s);
    }
    ConvArgs ca{};
    ca.x1 = a->x1; ca.x2 = a->x2; ca.C1 = c1; ca.C2 = c2; ca.N = n; ca.H = h; ca.W = w; ca.Cout = cout; ca.ksize = ksize;
    ca.stride = stride; ca.pad = pad; ca.Ho = ho; ca.Wo = wo; ca.w = wm; ca.bias = a->bias; ca.addvec = a->addvec;
    ca.addvec_row = a->addvec_row; ca.addvec_stride = cout; ca.residual = tmp ? tmp : a->residual; ca.out_scale = a->out_scale; ca.out = a->out;
    if (!rc) rc = launch_conv_simt(dt, ca, s);
    if (!rc && a->stats_out) rc = launch_gn_chunk_stats(dt, a->out, n, ho * wo, cout, cw, a->stats_out, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(wm);
    if (tmp) cudaFree(tmp);
    if (wsc) cudaFree(wsc);
    if (!rc && e != cudaSuccess) { set_error(std::string("conv_simt kernel failed: ") + cudaGetErrorString(e)); rc = 2; }
    return rc;
}

int pd_test_gn_conv(int32_t dt, int32_t n, int32_t h, int32_t w, int32_t c1, int32_t c2, int32_t cout, int32_t groups, float eps,
                    const void* x1, const void* x2, const float* gamma, const float* beta, const float* weight, const float* bias,
                    const float* addvec, const void* residual, const void* sc1, const void* sc2, int32_t csc1, int32_t csc2,
                    const float* sc_w, float out_scale, void* out, double* stats_out, pd_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    PD_REQUIRE(dt == DT_BF16 || dt == DT_F16, "fused GroupNorm convolution takes bf16 / fp16 activations");
    PD_REQUIRE(x1 && gamma && beta && weight && out, "null argument");
    const int ct = c1 + c2, hw = h * w;
    PD_REQUIRE(groups > 0 && ct % groups == 0, "channels not divisible by groups");
    int cw = 4;
    while (cw > 1 && ((ct / groups) % cw != 0 || c1 % cw != 0)) cw >>= 1;
    const int ktot = 9 * ct + csc1 + csc2;
    const size_t n1 = (size_t)n * (c1 / cw) * 2, n2 = (size_t)n * (c2 / cw) * 2;
    double* stats = nullptr;
    float2* coef = nullptr;
    void* wm = nullptr;
    PD_CHECK_CUDA(cudaMalloc((void**)&stats, (n1 + n2 + 2) * sizeof(double)));
    PD_CHECK_CUDA(cudaMalloc((void**)&coef, (size_t)n * ct * sizeof(float2)));
    PD_CHECK_CUDA(cudaMalloc(&wm, (size_t)cout * ktot * 2));
    PD_CHECK_CUDA(cudaMemsetAsync(stats, 0, (n1 + n2 + 2) * sizeof(double), s));
    PD_CHECK_CUDA(cudaMemsetAsync(wm, 0, (size_t)cout * ktot * 2, s));
    GNArgs ga{};
    ga.C1 = c1; ga.C2 = c2; ga.N = n; ga.HW = hw; ga.groups = groups; ga.eps = eps; ga.gamma = gamma; ga.beta = beta; ga.silu = 1;
    ga.stats_cw = cw; ga.stats1 = stats; ga.stats2 = c2 ? stats + n1 : nullptr;
    int rc = launch_gn_chunk_stats(dt, x1, n, hw, c1, cw, stats, s);
    if (!rc && c2) rc = launch_gn_chunk_stats(dt, x2, n, hw, c2, cw, stats + n1, s);
    if (!rc) rc = launch_gn_coef(ga, coef, s);
    if (!rc) rc = launch_relayout_tc(dt, weight, cout, ct, 3, wm, ktot, 0, s);
    if (!rc && (csc1 + csc2)) rc = launch_relayout_tc(dt, sc_w, cout, csc1 + csc2, 1, wm, ktot, 9 * ct, s);
    ConvTcDesc d{};
    d.dt = dt; d.x = x1; d.x2 = x2; d.C = ct; d.C2 = c2; d.gn_coef = coef; d.N = n; d.H = h; d.W = w; d.ksize = 3; d.stride = 1; d.pad = 1;
    d.Ho = h; d.Wo = w; d.Cout = cout; d.sc1 = sc1; d.Csc1 = csc1; d.sc2 = sc2; d.Csc2 = csc2; d.wmat = wm; d.bias = bias;
    d.addvec = addvec; d.addvec_stride = cout; d.residual = residual; d.out_scale = out_scale; d.out = out; d.stats_out = stats_out;
    d.stats_cw = 4; d.mode = TC_MODE_STD;
    ConvHaloPlan* hp = nullptr;
    if (!rc) rc = conv_halo_plan_create(d, &hp);
    if (!rc) rc = conv_halo_launch(hp, s, nullptr);
    cudaError_t e = cudaStreamSynchronize(s);
    if (hp) conv_halo_plan_destroy(hp);
    cudaFree(stats); cudaFree(coef); cudaFree(wm);
    if (!rc && e != cudaSuccess) { set_error(std::string("fused GroupNorm conv kernel failed: ") + cudaGetErrorString(e)); rc = 2; }
    return rc;
}

int pd_test_conv(int32_t use_tc, int32_t dt, int32_t n, int32_t h, int32_t w, int32_t c1, int32_t c2, int32_t cout,
                 int32_t ksize, int32_t stride, int32_t pad, const void* x1, const void* x2, const float* weight,
                 const float* bias, const float* addvec, const void* residual, const void* sc1, const void* sc2,
                 int32_t csc1, int32_t csc2, const float* sc_w, float out_scale, void* out, pd_stream_t stream) {
    pd_test_conv_args_t a;
    memset(&a, 0, sizeof(a));
    a.impl = use_tc; a.dtype = dt; a.n = n; a.h = h; a.w = w; a.c1 = c1; a.c2 = c2; a.cout = cout; a.ksize = ksize;
    a.stride = stride; a.pad = pad; a.x1 = x1; a.x2 = x2; a.weight = weight; a.bias = bias; a.addvec = addvec;
    a.residual = residual; a.sc1 = sc1; a.sc2 = sc2; a.csc1 = csc1; a.csc2 = csc2; a.sc_w = sc_w; a.out_scale = out_scale;
    a.out = out;
    return pd_test_conv_ex(&a, stream);
}

int pd_test_groupnorm(int32_t dt, int32_t n, int32_t hw, int32_t c1, int32_t c2, int32_t groups, float eps,
                      int32_t do_silu, const void* x1, const void* x2, const float* gamma, const float* beta, void* out,
                      pd_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    PD_REQUIRE(groups > 0 && (c1 + c2) % groups == 0, "channels not divisible by groups");
    int cw = 4;
    while (cw > 1 && (((c1 + c2) / groups) % cw != 0 || c1 % cw != 0)) cw >>= 1;
    double* stats = nullptr;
    const size_t n1 = (size_t)n * (c1 / cw) * 2, n2 = (size_t)n * (c2 / cw) * 2;
    PD_CHECK_CUDA(cudaMalloc((void**)&stats, (n1 + n2 + 2) * sizeof(double)));
    PD_CHECK_CUDA(cudaMemsetAsync(stats, 0, (n1 + n2 + 2) * sizeof(double), s));
    GNArgs ga{};
    ga.x1 = x1; ga.x2 = x2; ga.C1 = c1; ga.C2 = c2; ga.N = n; ga.HW = hw; ga.groups = groups; ga.eps = eps; ga.gamma = gamma;
    ga.beta = beta; ga.silu = do_silu; ga.stats1 = stats; ga.stats2 = c2 ? stats + n1 : nullptr; ga.stats_cw = cw; ga.out = out;
    int rc = launch_gn_chunk_stats(dt, x1, n, hw, c1, cw, stats, s);
    if (!rc && c2) rc = launch_gn_chunk_stats(dt, x2, n, hw, c2, cw, stats + n1, s);
    if (!rc) rc = launch_gn_apply(dt, dt == 0, ga, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(stats);
    if (!rc && e != cudaSuccess) { set_error(std::string("groupnorm kernel failed: ") + cudaGetErrorString(e)); rc = 2; }
    return rc;
}

int pd_test_attention(int32_t use_mma, int32_t dt, int32_t n, int32_t s_len, int32_t c, int32_t d, const void* qkv,
                      void* out, pd_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (use_mma) {
        rc = launch_attention_mma(dt, qkv, n, s_len, c, d, use_mma >= 2 ? PD_ATTN_QFOLD : 1.0f, out, s, use_mma >= 3 ? use_mma - 1 : 0);
    } else {
        rc = launch_attention_simt(dt, dt == 0, qkv, n, s_len, c, d, out, s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (!rc && e != cudaSuccess) { set_error(std::string("attention kernel failed: ") + cudaGetErrorString(e)); rc = 2; }
    return rc;
}

}  // extern "C"
