// phendiff_b200 — the model graph shared by the inference executor (pd_api.cu) and the training step (pd_train.cu):
// parameters in diffusers checkpoint naming, layer records mirroring what CustomCondUNet2DModel.__init__
// builds (reference: src/cond_unet_2d/cond_unet_2d.py:126-242), and the handle behind `pd_unet_t`.
#pragma once
#include "pd_kernels.h"
#include "pd_tc_common.cuh"

#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace pd {

// ---------------------------------------------------------------------------------------------------------------------
// parameters and layers
// ---------------------------------------------------------------------------------------------------------------------
struct Param {
    std::string name;
    std::vector<int64_t> shape;
    size_t numel = 0;
    float* dev = nullptr;
    bool loaded = false;
};

struct ConvL {
    Param *w = nullptr, *b = nullptr;
    int cin = 0, cout = 0, k = 3, stride = 1, pad = 1;
    float* w_simt = nullptr;  // (k*k*cin, cout) fp32
    void* w_tc = nullptr;     // (cout, k*k*cin) bf16/fp16
};
struct GNL { Param *g = nullptr, *b = nullptr; int C = 0; };
struct ResL {
    GNL n1, n2;
    ConvL c1, c2, sc;
    bool has_sc = false;
    Param *tw = nullptr, *tb = nullptr;
    int cin = 0, cout = 0, temb_off = 0;
    float scale = 1.f;
    void* w2sc_tc = nullptr;   // (cout, 9*cout + cin): conv2 and the 1x1 shortcut as ONE K-concatenated GEMM
    float* b2sc = nullptr;     // conv2.bias + conv_shortcut.bias
};
struct AttnL {
    GNL gn;
    Param *qw, *qb, *kw, *kb, *vw, *vb, *ow, *ob;
    int C = 0;
    float rescale = 1.f;
    float* wqkv_raw = nullptr;   // (3C, C) fp32, rows q|k|v
    float* bqkv = nullptr;       // (3C)
    float* wqkv_simt = nullptr;  // (C, 3C)
    void* wqkv_tc = nullptr;     // (3C, C)
    void* wqkv_tc_fold = nullptr;   // (3C, C) with the q rows pre-multiplied by PD_ATTN_QFOLD (head_dim-8 tensor-core attention)
    float* bqkv_fold = nullptr;  // (3C) bias matching wqkv_tc_fold
    float* wo_simt = nullptr;    // (C, C) transposed
    void* wo_tc = nullptr;       // (C, C)
};
struct DownB { std::vector<ResL> res; std::vector<AttnL> attn; bool has_attn = false, has_down = false; ConvL down; };
struct UpB { std::vector<ResL> res; std::vector<AttnL> attn; bool has_attn = false, has_up = false; ConvL up; void* w_up_tc = nullptr; };

struct Arena {
    bool dry = true;
    uint8_t* base = nullptr;
    size_t peak = 0;
    std::vector<std::pair<size_t, size_t>> freel;  // (offset, size), sorted by offset
    void reset(bool d, uint8_t* b) { dry = d; base = b; peak = 0; freel.clear(); freel.push_back({0, (size_t)1 << 60}); }
    size_t alloc(size_t bytes) {
        bytes = (bytes + 1023) & ~(size_t)1023;
        for (size_t i = 0; i < freel.size(); ++i) {
            if (freel[i].second >= bytes) {
                size_t off = freel[i].first;
                freel[i].first += bytes;
                freel[i].second -= bytes;
                if (freel[i].second == 0) freel.erase(freel.begin() + i);
                peak = std::max(peak, off + bytes);
                return off;
            }
        }
        return (size_t)-1;
    }
    void release(size_t off, size_t bytes) {
        bytes = (bytes + 1023) & ~(size_t)1023;
        auto it = std::lower_bound(freel.begin(), freel.end(), std::make_pair(off, (size_t)0));
        it = freel.insert(it, {off, bytes});
        if (it + 1 != freel.end() && it->first + it->second == (it + 1)->first) {
            it->second += (it + 1)->second;
            freel.erase(it + 1);
        }
        if (it != freel.begin() && (it - 1)->first + (it - 1)->second == it->first) {
            (it - 1)->second += it->second;
            freel.erase(it);
        }
    }
};

struct Tensor {   // NHWC activation of the current micro-batch
    size_t off = 0, bytes = 0;
    int C = 0, H = 0, W = 0, refs = 0;
    size_t stats_off = (size_t)-1;   // chunk statistics slot (bytes into the statistics region), or -1: none
};

struct Ctx {   // per-call inputs of the recorded program
    const float* x = nullptr;          // (mb, Cin, H, W) sample
    const float* timesteps = nullptr;  // (mb) or null -> t_scalar
    float t_scalar = 0.f;
    const int64_t* labels = nullptr;
    const float* class_emb = nullptr;
    float* model_out = nullptr;        // (mb, Cout, H, W) or null
    float* x_update = nullptr;         // x_t updated in place by the fused conv_out epilogue, or null
    const pd_step_coeffs_t* step = nullptr;
    // classifier-free guidance in one pass: the mb images are P = cfg_pairs conditional samples followed by their P
    // unconditional copies (x holds P images; image P + i reads x[i] and the embedding row without class embedding)
    int cfg_pairs = 0;
    const float* cfg_w = nullptr;      // (P) guidance scale per sample
    int cfg_eqn = 0;                   // 0 "imagen", 1 "CFG"
};

typedef std::function<int(const Ctx&, cudaStream_t)> OpFn;
enum { CLS_CONV_TC = 0, CLS_CONV_SIMT, CLS_GN, CLS_ATTN, CLS_EMBED, CLS_CONV_IN, CLS_CONV_OUT, CLS_UPSAMPLE, CLS_COUNT };
struct Op { OpFn fn; int cls; double flops; int nlaunch; std::string name; double ms = 0; int samples = 0; };

}  // namespace pd

using namespace pd;

struct pd_unet {
    pd_unet_config_t cfg;
    int device = 0;
    int dt = DT_F32;   // activation storage type
    bool half = false; // bf16 / fp16: tensor-core path available
    int D = 0;  // time_embed_dim
    int J = 0;  // total time_emb_proj outputs
    std::vector<std::unique_ptr<Param>> params;
    std::map<std::string, Param*> by_name;
    std::map<std::string, std::string> alias;
    // layers
    ConvL conv_in, conv_out;
    Param *te_w1, *te_b1, *te_w2, *te_b2, *cls = nullptr;
    std::vector<DownB> down;
    ResL mid_r0, mid_r1;
    AttnL mid_attn;
    bool mid_has_attn = true;
    std::vector<UpB> up;
    GNL norm_out;
    float* wcat = nullptr;      // (J, D)
    float* bcat = nullptr;      // (J)  time_emb_proj.bias + conv1.bias
    float* w_in = nullptr;      // (9*Cin, C0)
    float* w_out = nullptr;     // (9, C0, 4)
    void* w_in_tc = nullptr;    // (C0, 64) 16-bit: conv_in as a K = 64 GEMM over im2col rows (k = tap*Cin + ci)
    void* w_out_tc = nullptr;   // (16, 9*C0) 16-bit: conv_out rows zero-padded to 16
    int stats_cw = 4;           // channels per GroupNorm statistics chunk (divides every group width)
    size_t stats_needed = 0, rowidx_off = 0;
    bool finalized = false;
    std::vector<void*> owned;   // derived device buffers
    // plan
    int B = 0, H = 0, W = 0, mb = 0;
    bool pairs = false;         // plan made by pd_unet_plan_guided: B = 2 x samples, mb even
    int tail = 0;               // images in the last micro-batch when the batch is ragged (B % mb), else 0
    size_t tail_x_off = 0, tail_out_off = 0, tail_t_off = 0, tail_lab_off = 0, tail_emb_off = 0;   // padded scratch copies of the tail
    size_t ws_bytes = 0;
    Arena arena;
    std::vector<Op> ops;
    std::vector<ConvTcPlan*> tc_plans;
    std::vector<std::unique_ptr<Tensor>> tensors;
    bool bound = false;
    size_t stats_off = 0, stats_bytes = 0, emb_off = 0, temb_off = 0, cfg_u_off = 0;
    int64_t launches = 0;
    int tc_layers = 0, simt_layers = 0;
    // sampled per-op device timing (bench.py's roofline): every `prof_every`-th program run is bracketed with events
    int prof_every = 0, prof_max = 0;
    int64_t prof_runs = 0;
    std::vector<std::vector<cudaEvent_t>> prof_events;   // one chain of (ops + 1) events per sampled run
    double prof_ms[CLS_COUNT] = {0};
    double prof_flops[CLS_COUNT] = {0};
    int64_t prof_launches[CLS_COUNT] = {0};
};
