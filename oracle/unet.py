"""ORACLE (test infrastructure, never shipped on the product path).

Plain-PyTorch fp32 restatement of the reference's conditional UNet:
  * graph + forward: src/cond_unet_2d/cond_unet_2d.py:73-362 (CustomCondUNet2DModel)
  * the blocks it instantiates come from the un-vendored dependency `diffusers==0.18.2`
    (environment.yaml:80): get_down_block / get_up_block / UNetMidBlock2D -> ResnetBlock2D,
    Attention(AttnProcessor2_0), Downsample2D, Upsample2D, Timesteps, TimestepEmbedding.
    Their behaviour is restated from SURVEY.md Appendix A.1/A.2 with torch functional ops only.

PARITY PINNED PER BLOCK: the reference has no tests / golden tensors for this path and diffusers is not
importable here, but diffusers' own test-suite (tests/models/test_layers_utils.py) publishes known answers for exactly
these blocks under `torch.manual_seed(0)` + default initialisation: ResnetBlock2D (with and without the 1x1 shortcut),
the attention block (32 heads of dim 1; one head of dim 512), Upsample2D / Downsample2D with conv, and the sinusoid
embedding in the reference's (flip_sin_to_cos, freq_shift 0) setting.  Every block here reproduces its vector to the
printed precision (tests/test_oracle_published_kats.py).  The ASSEMBLY of the blocks into the UNet is pinned too: with the
weights re-drawn in diffusers' module-creation order (tests/util.py::diffusers_order_init) the whole graph + DDIM loop
reproduces the image corner published in tests/pipelines/ddim/test_ddim.py (DDIMPipelineFastTests.test_inference).  What
that vector cannot cover is the reference's own class conditioning (emb + class_embedding(labels), cond_unet_2d.py:297-309).  Parameter names follow the diffusers checkpoint layout (Appendix A.7) so a real PhenDiff
state_dict loads; parameter counts are pinned to SURVEY §8 (62 826 243 / 15 725 443).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .schedulers import _Config


class Timesteps(nn.Module):
    # diffusers embeddings.get_timestep_embedding (A.1): fp32 sinusoid, [cos | sin] when flip_sin_to_cos
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.shift)
        emb = torch.exp(exponent)
        emb = timesteps[:, None].float() * emb[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        if self.num_channels % 2 == 1:
            emb = F.pad(emb, (0, 1, 0, 0))
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, groups, eps, scale=1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0) if cin != cout else None
        self.scale = scale

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.scale


class Attention(nn.Module):
    # A.2: deprecated-attn-block flavoured Attention with AttnProcessor2_0 numerics
    def __init__(self, channels, head_dim, groups, eps, rescale=1.0):
        super().__init__()
        self.heads = channels // head_dim
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])
        self.rescale = rescale

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = x.view(b, c, h * w).transpose(1, 2)
        t = self.group_norm(t.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        d = c // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        # F.scaled_dot_product_attention(q, k, v) == softmax(q k^T / sqrt(d)) v, stated explicitly in fp32
        s = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * d)
        o = self.to_out[0](o)
        o = o.transpose(-1, -2).reshape(b, c, h, w)
        return (o + res) / self.rescale


class Downsample2D(nn.Module):
    def __init__(self, ch, padding):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, groups, eps, head_dim, attn, add_down, pad):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        if attn:
            self.attentions = nn.ModuleList([Attention(cout, head_dim, groups, eps) for _ in range(layers)])
        self.has_attn = attn
        self.downsamplers = nn.ModuleList([Downsample2D(cout, pad)]) if add_down else None

    def forward(self, x, temb):
        outs = ()
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb, groups, eps, head_dim, scale, add_attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps, scale),
                                      ResnetBlock2D(ch, ch, temb, groups, eps, scale)])
        self.attentions = nn.ModuleList([Attention(ch, head_dim, groups, eps, scale) if add_attention else None])

    def forward(self, x, temb):
        x = self.resnets[0](x, temb)
        if self.attentions[0] is not None:
            x = self.attentions[0](x)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, prev, cout, temb, layers, groups, eps, head_dim, attn, add_up):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        if attn:
            self.attentions = nn.ModuleList([Attention(cout, head_dim, groups, eps) for _ in range(layers)])
        self.has_attn = attn
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb):
        skips = list(skips)
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


_UNET_DEFAULTS = dict(
    sample_size=None, in_channels=3, out_channels=3, center_input_sample=False, time_embedding_type="positional",
    freq_shift=0, flip_sin_to_cos=True,
    down_block_types=("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
    up_block_types=("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
    block_out_channels=(224, 448, 672, 896), layers_per_block=2, mid_block_scale_factor=1, downsample_padding=1,
    act_fn="silu", attention_head_dim=8, norm_num_groups=32, norm_eps=1e-5, resnet_time_scale_shift="default",
    add_attention=True, class_embed_type=None, num_class_embeds=None,
)


class OracleCondUNet2D(nn.Module):
    """cond_unet_2d.py:73-362 with the shipped option set (positional time embedding, class_embed_type None)."""

    def __init__(self, **kwargs):
        super().__init__()
        cfg = dict(_UNET_DEFAULTS)
        for k, v in kwargs.items():
            if k.startswith("_") or k not in cfg:
                continue  # A.7: ignore "_class_name" etc. and unknown keys
            cfg[k] = v
        self.config = c = _Config(cfg)
        if c.time_embedding_type != "positional" or c.class_embed_type is not None or c.act_fn != "silu" \
                or c.resnet_time_scale_shift != "default":
            raise NotImplementedError("only the shipped option set is restated (SURVEY A.8)")
        boc = list(c.block_out_channels)
        if len(c.down_block_types) != len(c.up_block_types) or len(boc) != len(c.down_block_types):
            raise ValueError("block type / channel lists disagree")  # cond_unet_2d.py:116-124
        self.sample_size = c.sample_size
        self.time_embed_dim = ted = boc[0] * 4  # cond_unet_2d.py:111-113
        g, eps = c.norm_num_groups, c.norm_eps
        self.conv_in = nn.Conv2d(c.in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], c.flip_sin_to_cos, c.freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], ted)
        self.class_embedding = nn.Embedding(c.num_class_embeds, ted) if c.num_class_embeds is not None else None
        self.down_blocks = nn.ModuleList()
        out = boc[0]
        for i, t in enumerate(c.down_block_types):
            cin, out = out, boc[i]
            final = i == len(boc) - 1
            hd = c.attention_head_dim if c.attention_head_dim is not None else out
            self.down_blocks.append(DownBlock(cin, out, ted, c.layers_per_block, g, eps, hd,
                                              t == "AttnDownBlock2D", not final, c.downsample_padding))
        hd = c.attention_head_dim if c.attention_head_dim is not None else boc[-1]
        self.mid_block = MidBlock(boc[-1], ted, g, eps, hd, c.mid_block_scale_factor, c.add_attention)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out = rev[0]
        for i, t in enumerate(c.up_block_types):
            prev, out = out, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            final = i == len(boc) - 1
            hd = c.attention_head_dim if c.attention_head_dim is not None else out
            self.up_blocks.append(UpBlock(cin, prev, out, ted, c.layers_per_block + 1, g, eps, hd,
                                          t == "AttnUpBlock2D", not final))
        ng = g if g is not None else min(boc[0] // 4, 32)
        self.conv_norm_out = nn.GroupNorm(ng, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], c.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def embed(self, sample, timestep, class_labels=None, class_emb=None):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.long, device=sample.device)
        elif len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps * torch.ones(sample.shape[0], dtype=timesteps.dtype, device=timesteps.device)
        emb = self.time_embedding(self.time_proj(timesteps).to(self.dtype))
        if self.class_embedding is not None:
            if class_labels is None and class_emb is None:
                raise ValueError("either class_labels or class_emb should be provided when doing class conditioning")
            if class_emb is None:
                class_emb = self.class_embedding(class_labels).to(self.dtype)
            emb = emb + class_emb
        return emb

    def forward(self, sample, timestep, class_labels=None, class_emb=None, return_dict=True):
        if class_labels is not None and class_emb is not None:
            raise ValueError("Cannot specify both class_labels and class_emb")
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0
        emb = self.embed(sample, timestep, class_labels, class_emb)
        sample = self.conv_in(sample)
        res = (sample,)
        for blk in self.down_blocks:
            sample, r = blk(sample, emb)
            res += r
        sample = self.mid_block(sample, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            sample = blk(sample, r, emb)
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return SimpleNamespace(sample=sample)
