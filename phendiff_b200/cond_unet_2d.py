"""CustomCondUNet2DModel — drop-in for the reference's class-conditional UNet
(reference: src/cond_unet_2d/cond_unet_2d.py:29-362; constructor kwargs = the keys of models_configs/denoiser/*.json).

Same constructor keywords, `forward(sample, timestep, class_labels=None, class_emb=None, return_dict=True)`,
`.time_embed_dim`, `.config`, `.dtype`, `.device`, `.class_embedding`, `load_config` / `from_config`, and a
`state_dict()` in diffusers checkpoint naming — but the arithmetic runs in hand-written sm_100a
kernels behind the C ABI (include/phendiff_b200.h).  This class only owns the parameters and the plumbing.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib
from .config import ConfigMixin


@dataclass
class UNet2DOutput:
    sample: torch.Tensor


class _Holder(nn.Module):
    """Parameter container; the maths lives in the CUDA library."""

    def forward(self, *a, **k):  # pragma: no cover
        raise _lib.PhenDiffB200Error("sub-modules of CustomCondUNet2DModel only hold parameters; call the model itself")


_PRECISIONS = {"bf16": _lib.PD_PREC_BF16, "fp16": _lib.PD_PREC_FP16, "fp32": _lib.PD_PREC_FP32}


class CustomCondUNet2DModel(nn.Module, ConfigMixin):
    config_name = "config.json"

    def __init__(
        self,
        sample_size: Optional[Union[int, Tuple[int, int]]] = None,
        in_channels: int = 3,
        out_channels: int = 3,
        center_input_sample: bool = False,
        time_embedding_type: str = "positional",
        freq_shift: int = 0,
        flip_sin_to_cos: bool = True,
        down_block_types: Tuple[str] = ("DownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D", "AttnDownBlock2D"),
        up_block_types: Tuple[str] = ("AttnUpBlock2D", "AttnUpBlock2D", "AttnUpBlock2D", "UpBlock2D"),
        block_out_channels: Tuple[int] = (224, 448, 672, 896),
        layers_per_block: int = 2,
        mid_block_scale_factor: float = 1,
        downsample_padding: int = 1,
        act_fn: str = "silu",
        attention_head_dim: Optional[int] = 8,
        norm_num_groups: int = 32,
        norm_eps: float = 1e-5,
        resnet_time_scale_shift: str = "default",
        add_attention: bool = True,
        class_embed_type: Optional[str] = None,
        num_class_embeds: Optional[int] = None,
        precision: Optional[str] = None,
        max_microbatch: Optional[int] = None,
        conv_impl: Optional[str] = None,
        attn_impl: Optional[str] = None,
    ):
        super().__init__()
        self.register_to_config(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            center_input_sample=center_input_sample, time_embedding_type=time_embedding_type, freq_shift=freq_shift,
            flip_sin_to_cos=flip_sin_to_cos, down_block_types=tuple(down_block_types),
            up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
            layers_per_block=layers_per_block, mid_block_scale_factor=mid_block_scale_factor,
            downsample_padding=downsample_padding, act_fn=act_fn, attention_head_dim=attention_head_dim,
            norm_num_groups=norm_num_groups, norm_eps=norm_eps, resnet_time_scale_shift=resnet_time_scale_shift,
            add_attention=add_attention, class_embed_type=class_embed_type, num_class_embeds=num_class_embeds)
        self.sample_size = sample_size
        self.time_embed_dim = block_out_channels[0] * 4  # cond_unet_2d.py:111-113

        # the same input checks as the reference (cond_unet_2d.py:116-124)
        if len(down_block_types) != len(up_block_types):
            raise ValueError(
                f"Must provide the same number of `down_block_types` as `up_block_types`. `down_block_types`: {down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != len(down_block_types):
            raise ValueError(
                f"Must provide the same number of `block_out_channels` as `down_block_types`. `block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        # only the shipped option set has native kernels; refuse the rest rather than silently diverging
        if time_embedding_type != "positional":
            raise NotImplementedError("time_embedding_type 'fourier' is not used by any shipped config and is not implemented")
        if class_embed_type is not None:
            raise NotImplementedError("class_embed_type 'timestep'/'identity' are not used by any shipped config and are not implemented")
        if act_fn != "silu" or resnet_time_scale_shift != "default":
            raise NotImplementedError("only act_fn='silu' and resnet_time_scale_shift='default' are implemented")
        for t in down_block_types:
            if t not in ("DownBlock2D", "AttnDownBlock2D"):
                raise NotImplementedError(f"down block type {t} is not implemented")
        for t in up_block_types:
            if t not in ("UpBlock2D", "AttnUpBlock2D"):
                raise NotImplementedError(f"up block type {t} is not implemented")
        if len(block_out_channels) > _lib.PD_MAX_BLOCKS:
            raise NotImplementedError("too many blocks")

        self._precision = (precision or os.environ.get("PHENDIFF_B200_PRECISION", "fp16")).lower()
        if self._precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {list(_PRECISIONS)}")
        self._max_microbatch = int(max_microbatch if max_microbatch is not None else os.environ.get("PHENDIFF_B200_MICROBATCH", 0))
        self._conv_impl = (conv_impl or os.environ.get("PHENDIFF_B200_CONV", "tcgen05")).lower()
        self._attn_impl = (attn_impl or os.environ.get("PHENDIFF_B200_ATTN", "mma")).lower()

        self._handle = None
        self._handle_device = None
        self._synced_version = None
        self._synced_fingerprint = None
        self._dirty_epoch = 0
        self._plan_key = None
        self._workspace = None
        self._build_parameter_tree()

    # ------------------------------------------------------------------------------------------------------------
    # parameters (diffusers checkpoint names, Appendix A.7)
    # ------------------------------------------------------------------------------------------------------------
    def _param_table(self):
        """(name, shape) of every parameter, in the library's order.  Pure host logic (no GPU needed)."""
        c = self.config
        boc, L, D = list(c.block_out_channels), c.layers_per_block, self.time_embed_dim
        out = []

        def conv(p, ci, co, k):
            out.append((p + ".weight", (co, ci, k, k)))
            out.append((p + ".bias", (co,)))

        def norm(p, ch):
            out.append((p + ".weight", (ch,)))
            out.append((p + ".bias", (ch,)))

        def lin(p, ci, co):
            out.append((p + ".weight", (co, ci)))
            out.append((p + ".bias", (co,)))

        def res(p, ci, co):
            norm(p + ".norm1", ci); conv(p + ".conv1", ci, co, 3); lin(p + ".time_emb_proj", D, co)
            norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
            if ci != co:
                conv(p + ".conv_shortcut", ci, co, 1)

        def attn(p, ch):
            norm(p + ".group_norm", ch)
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(p + "." + n, ch, ch)

        conv("conv_in", c.in_channels, boc[0], 3)
        lin("time_embedding.linear_1", boc[0], D)
        lin("time_embedding.linear_2", D, D)
        if c.num_class_embeds is not None:
            out.append(("class_embedding.weight", (c.num_class_embeds, D)))
        oc = boc[0]
        for i, t in enumerate(c.down_block_types):
            ic, oc = oc, boc[i]
            for j in range(L):
                res(f"down_blocks.{i}.resnets.{j}", ic if j == 0 else oc, oc)
                if t == "AttnDownBlock2D":
                    attn(f"down_blocks.{i}.attentions.{j}", oc)
            if i != len(boc) - 1:
                conv(f"down_blocks.{i}.downsamplers.0.conv", oc, oc, 3)
        res("mid_block.resnets.0", boc[-1], boc[-1])
        if c.add_attention:
            attn("mid_block.attentions.0", boc[-1])
        res("mid_block.resnets.1", boc[-1], boc[-1])
        rev = list(reversed(boc))
        oc = rev[0]
        for i, t in enumerate(c.up_block_types):
            prev, oc = oc, rev[i]
            ic = rev[min(i + 1, len(boc) - 1)]
            for j in range(L + 1):
                skip = ic if j == L else oc
                rin = prev if j == 0 else oc
                res(f"up_blocks.{i}.resnets.{j}", rin + skip, oc)
                if t == "AttnUpBlock2D":
                    attn(f"up_blocks.{i}.attentions.{j}", oc)
            if i != len(boc) - 1:
                conv(f"up_blocks.{i}.upsamplers.0.conv", oc, oc, 3)
        norm("conv_norm_out", boc[0])
        conv("conv_out", boc[0], c.out_channels, 3)
        return out

    def _build_parameter_tree(self):
        table = self._param_table()
        shapes = dict(table)
        self.class_embedding = None
        for name, shape in table:
            parts = name.split(".")
            if parts[0] == "class_embedding":
                self.class_embedding = nn.Embedding(shape[0], shape[1])  # N(0,1) init, like the reference
                continue
            mod = self
            for p in parts[:-1]:
                if not hasattr(mod, p) or getattr(mod, p) is None:
                    mod.add_module(p, _Holder())
                mod = getattr(mod, p)
            leaf = parts[-1]
            t = torch.empty(shape, dtype=torch.float32)
            if leaf == "weight" and len(shape) == 1:
                nn.init.ones_(t)
            elif leaf == "bias" and ".norm" in "." + name or leaf == "bias" and "group_norm" in name or leaf == "bias" and name.startswith("conv_norm_out"):
                nn.init.zeros_(t)
            elif leaf == "weight":
                nn.init.kaiming_uniform_(t, a=math.sqrt(5))  # torch default for Conv2d / Linear
            else:
                wshape = shapes[name[: -len("bias")] + "weight"]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
                bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
                nn.init.uniform_(t, -bound, bound)
            mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))

    # ------------------------------------------------------------------------------------------------------------
    # properties the reference reads
    # ------------------------------------------------------------------------------------------------------------
    @property
    def dtype(self) -> torch.dtype:
        return self.conv_in.weight.dtype

    @property
    def device(self) -> torch.device:
        return self.conv_in.weight.device

    @property
    def precision(self) -> str:
        return self._precision

    def set_precision(self, precision: str):
        """'fp16' / 'bf16' (tensor-core product path, fp32 accumulation) or 'fp32' (validation mode, SIMT fp32 kernels)."""
        precision = precision.lower()
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {list(_PRECISIONS)}")
        if precision != self._precision:
            self._precision = precision
            self._destroy_handle()
        return self

    # ------------------------------------------------------------------------------------------------------------
    # C handle management
    # ------------------------------------------------------------------------------------------------------------
    def _destroy_handle(self):
        d = self.__dict__
        h = d.get("_handle")
        if h is not None:
            try:
                _lib.lib().pd_unet_destroy(h)
            except Exception:  # pragma: no cover - interpreter shutdown
                pass
        d["_handle"] = None
        d["_synced_version"] = None
        d["_synced_fingerprint"] = None
        d["_plan_key"] = None
        d["_workspace"] = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def _c_config(self) -> _lib.UnetConfig:
        c = self.config
        cc = _lib.UnetConfig()
        cc.in_channels, cc.out_channels = c.in_channels, c.out_channels
        cc.n_blocks = len(c.block_out_channels)
        for i, ch in enumerate(c.block_out_channels):
            cc.block_out_channels[i] = ch
            cc.down_attn[i] = int(c.down_block_types[i] == "AttnDownBlock2D")
            cc.up_attn[i] = int(c.up_block_types[i] == "AttnUpBlock2D")
        cc.layers_per_block = c.layers_per_block
        cc.attention_head_dim = c.attention_head_dim or 0   # None: one head of dim C (cond_unet_2d.py:176-178)
        cc.norm_num_groups = c.norm_num_groups
        cc.norm_eps = c.norm_eps
        cc.num_class_embeds = c.num_class_embeds or 0
        cc.flip_sin_to_cos = int(c.flip_sin_to_cos)
        cc.freq_shift = float(c.freq_shift)
        cc.downsample_padding = c.downsample_padding
        cc.mid_block_scale_factor = float(c.mid_block_scale_factor)
        cc.add_attention = int(c.add_attention)
        cc.precision = _PRECISIONS[self._precision]
        cc.max_microbatch = self._max_microbatch
        cc.conv_impl = 0 if self._conv_impl == "tcgen05" else 1
        cc.attn_impl = 0 if self._attn_impl == "mma" else 1
        return cc

    def _weights_version(self):
        """Cheap key of the parameter state: autograd version counter + storage address of every parameter.  It catches
        in-place optimiser steps, `load_state_dict`, `.to()` / `_apply` and re-assigned `.data`; it does NOT see a write
        through `param.data.copy_()` (what diffusers' `EMAModel.copy_to / restore` do, utils_training.py:674-676): those
        are caught by `weights_fingerprint()` on the whole-path entry points, or announced with `mark_dirty()`."""
        return tuple((p._version, p.data_ptr()) for p in self.parameters()) + (str(self.device), self._dirty_epoch)

    def mark_dirty(self):
        """Tell the model its parameters were changed out of band (e.g. `p.data.copy_(...)`): the library-owned,
        re-laid-out device copy is rebuilt on the next call."""
        self._dirty_epoch += 1
        return self

    def sync_weights(self):
        """Force the re-upload of the parameters into the library's buffers now."""
        self.mark_dirty()
        self._ensure_handle()
        return self

    @torch.no_grad()
    def weights_fingerprint(self) -> float:
        """Position-weighted sum of the parameters' L1 norms, in fp64 (one tiny device reduction + ONE host read).  The
        whole-path entry points (`_ddib`, `_inversion`, the fused pipeline loop: seconds of device work per call) compare it
        with the value taken at the last upload, so an EMA `copy_to` right before sampling is never silently ignored; the
        per-op `forward` stays free of host synchronisation and relies on `_weights_version()` / `mark_dirty()`."""
        ps = [p.detach() for p in self.parameters()]
        norms = torch.stack(torch._foreach_norm(ps, 1)).double()
        w = torch.arange(1, len(ps) + 1, dtype=torch.float64, device=norms.device)
        return float((norms * w).sum().item())

    def check_weights(self):
        """Re-upload the parameters if their fingerprint moved since the last upload (whole-path entry points call this)."""
        if self._handle is not None and self._synced_fingerprint is not None \
                and self._synced_version == self._weights_version() and self.weights_fingerprint() != self._synced_fingerprint:
            self.mark_dirty()
        return self

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.__dict__["_dirty_epoch"] = self.__dict__.get("_dirty_epoch", 0) + 1
        return out

    def _ensure_handle(self):
        dev = self.device
        if dev.type != "cuda":
            raise _lib.PhenDiffB200Error(
                f"CustomCondUNet2DModel lives on {dev}: move it to a CUDA device (there is no CPU fallback)")
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._handle is None or self._handle_device != dev:
                self._destroy_handle()
                h = C.c_void_p()
                cc = self._c_config()
                _lib.check(L.pd_unet_create(C.byref(cc), C.byref(h)))
                self._handle, self._handle_device = h, dev
                # host-side table and the library's table must agree (names and shapes)
                n = C.c_int32()
                _lib.check(L.pd_unet_num_params(h, C.byref(n)))
                mine = self._param_table()
                if n.value != len(mine):
                    raise _lib.PhenDiffB200Error(f"parameter table mismatch: library {n.value} vs host {len(mine)}")
            ver = self._weights_version()
            if self._synced_version != ver:
                sd = self.state_dict()
                for name, shape in self._param_table():
                    t = sd[name].detach().to(device=dev, dtype=torch.float32).contiguous()
                    arr = (C.c_int64 * len(shape))(*shape)
                    _lib.check(L.pd_unet_load_weight(self._handle, name.encode(), _lib.ptr(t), arr, len(shape)))
                _lib.check(L.pd_unet_finalize(self._handle, _lib.current_stream()))
                self._synced_version = ver
                self._synced_fingerprint = self.weights_fingerprint()
                self._plan_key = None
        return self._handle

    def _ensure_plan(self, B: int, H: int, W: int, guided: bool = False):
        """Static buffer plan + workspace for B samples of H x W.  guided: plan for the classifier-free-guidance pass
        (2B images per step: every sample once with its class row, once with the unconditional row; `pd_cfg_transfer`)."""
        h = self._ensure_handle()
        key = (B, H, W, bool(guided))
        if self._plan_key != key:
            L = _lib.lib()
            nbytes = C.c_size_t()
            _lib.check((L.pd_unet_plan_guided if guided else L.pd_unet_plan)(h, B, H, W, C.byref(nbytes)))
            ws = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device=self.device)
            off = (-ws.data_ptr()) % 1024
            _lib.check(L.pd_unet_bind_workspace(h, C.c_void_p(ws.data_ptr() + off), nbytes.value))
            self._workspace = ws
            self._plan_key = key
        return h

    def launch_count(self) -> int:
        if self._handle is None:
            return 0
        n = C.c_int64()
        _lib.check(_lib.lib().pd_unet_launch_count(self._handle, C.byref(n)))
        return n.value

    KERNEL_CLASSES = ("conv_tcgen05", "conv_simt", "groupnorm", "attention", "embedding", "conv_in", "conv_out_ddim", "upsample")

    def plan_info(self) -> dict:
        mb, tc, simt, ops = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().pd_unet_plan_info(self._handle, C.byref(mb), C.byref(tc), C.byref(simt), C.byref(ops)))
        return {"microbatch": mb.value, "tcgen05_layers": tc.value, "simt_layers": simt.value, "ops_per_forward": ops.value,
                "workspace_bytes": int(self._workspace.numel()) if self._workspace is not None else 0}

    def profile_begin(self, every_n: int = 50, max_samples: int = 64):
        _lib.check(_lib.lib().pd_unet_profile_begin(self._handle, every_n, max_samples))

    def profile_end(self) -> dict:
        """Per kernel class: summed device ms, launches and algorithmic FLOPs over the sampled forwards."""
        n = C.c_int32()
        _lib.check(_lib.lib().pd_unet_profile_end(self._handle, C.byref(n)))
        out = {"samples": n.value}
        for i, name in enumerate(self.KERNEL_CLASSES):
            ms, la, fl = C.c_double(), C.c_int64(), C.c_double()
            _lib.check(_lib.lib().pd_unet_profile_query(self._handle, i, C.byref(ms), C.byref(la), C.byref(fl)))
            out[name] = {"ms": ms.value, "launches": la.value, "flops": fl.value}
        return out

    def profile_ops(self) -> list:
        """Per recorded op after profile_end(): dict(name, cls, ms (mean per forward), flops)."""
        out = []
        for i in range(self.plan_info()["ops_per_forward"]):
            name, cls, ms, ns, fl = C.c_char_p(), C.c_int32(), C.c_double(), C.c_int32(), C.c_double()
            _lib.check(_lib.lib().pd_unet_profile_op(self._handle, i, C.byref(name), C.byref(cls), C.byref(ms), C.byref(ns), C.byref(fl)))
            out.append({"name": (name.value or b"").decode(), "cls": self.KERNEL_CLASSES[cls.value],
                        "ms": ms.value / max(ns.value, 1), "flops": fl.value})
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # A.7: diffusers 0.17-0.19 wrote attention weights under deprecated names; accept both spellings
        ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        fixed = {}
        for k, v in state_dict.items():
            parts = k.split(".")
            if "attentions" in parts and len(parts) >= 2 and parts[-2] in ren:
                parts[-2:-1] = ren[parts[-2]].split(".")
                k = ".".join(parts)
            own = dict(self.named_parameters()).get(k)
            if own is not None and v.numel() == own.numel() and v.shape != own.shape:
                v = v.reshape(own.shape)  # deprecated attention blocks stored 1x1-conv shaped linears
            fixed[k] = v
        self._dirty_epoch += 1   # `assign=True` resets version counters; never trust them across a load
        return super().load_state_dict(fixed, strict=strict, **kw)

    # ------------------------------------------------------------------------------------------------------------
    # forward (cond_unet_2d.py:244-362)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(
        self,
        sample: torch.FloatTensor,
        timestep: Union[torch.Tensor, float, int],
        class_labels: Optional[torch.Tensor] = None,
        class_emb: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ) -> Union[UNet2DOutput, Tuple]:
        if class_labels is not None and class_emb is not None:
            raise ValueError("Cannot specify both class_labels and class_emb")
        if self.class_embedding is not None and class_labels is None and class_emb is None:
            raise ValueError("either class_labels or class_emb should be provided when doing class conditioning")
        _lib.require_cuda(sample, "sample")
        if sample.dim() != 4 or sample.shape[1] != self.config.in_channels:
            raise ValueError(f"sample must be (batch, {self.config.in_channels}, height, width)")
        dev = sample.device
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0
        x = sample.to(torch.float32).contiguous()
        B, _, H, W = x.shape

        # 1. time (cond_unet_2d.py:276-287): scalar / 0-dim / 1-D -> (B,) on the device
        if not torch.is_tensor(timestep):
            ts = torch.full((B,), float(timestep), dtype=torch.float32, device=dev)
        elif timestep.dim() == 0 and not timestep.is_cuda:
            ts = torch.full((B,), float(timestep.item()), dtype=torch.float32, device=dev)   # CPU scalar (scheduler.timesteps): no device sync
        elif timestep.dim() == 0:
            ts = timestep.to(device=dev, dtype=torch.float32).expand(B).contiguous()          # device scalar: stays on the device
        else:
            ts = (timestep.to(dev).to(torch.float32) * torch.ones(B, dtype=torch.float32, device=dev)).contiguous()

        labels = emb = None
        if self.class_embedding is not None:
            if class_emb is not None:
                emb = class_emb.to(device=dev, dtype=torch.float32).contiguous()
                if emb.shape != (B, self.time_embed_dim):
                    raise ValueError(f"class_emb must be (batch, {self.time_embed_dim})")
            else:
                labels = class_labels.to(device=dev, dtype=torch.int64).contiguous()
                if labels.shape != (B,):
                    raise ValueError("class_labels must be (batch,)")

        with torch.cuda.device(dev):
            h = self._ensure_plan(B, H, W)
            out = torch.empty((B, self.config.out_channels, H, W), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().pd_unet_forward(h, _lib.ptr(x), _lib.ptr(ts), _lib.ptr(labels), _lib.ptr(emb),
                                                  _lib.ptr(out), _lib.current_stream()))
        if not return_dict:
            return (out,)
        return UNet2DOutput(sample=out)

    # ------------------------------------------------------------------------------------------------------------
    # persistence in the diffusers directory layout (config.json + diffusion_pytorch_model.safetensors)
    # ------------------------------------------------------------------------------------------------------------
    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **kw):
        self.save_config(save_directory)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file

            save_file(sd, os.path.join(save_directory, "diffusion_pytorch_model.safetensors"))
        else:
            torch.save(sd, os.path.join(save_directory, "diffusion_pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, **kw):
        d = os.path.join(path, subfolder) if subfolder else path
        model = cls.from_config(cls.load_config(d), **kw)
        st = os.path.join(d, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file

            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(d, "diffusion_pytorch_model.bin"), map_location="cpu")
        model.load_state_dict(sd)
        return model
