"""ConditionalDDIMPipeline — drop-in for the reference's conditional DDIM pipeline
(reference: src/pipeline_conditional_ddim/pipeline_conditionial_ddim.py:27-361).

`__call__` keeps the reference's signature, checks and semantics (start image, partial trajectories, forward noising,
classifier-free guidance with the "imagen"/"CFG" equations).  Two execution routes, both CUDA-only:
  * fused route (no CFG, eta == 0, labels given — exactly how `_ddib` calls the pipeline): the whole n-step loop is
    one C-ABI call (`pd_ddib_transfer`) in which every scheduler update is fused into the UNet's conv_out epilogue;
  * general route: UNet forward (`pd_unet_forward`) + `pd_cfg_combine` + `pd_ddim_step` per step.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass
from typing import List, Literal, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from .cond_unet_2d import CustomCondUNet2DModel
from .schedulers import DDIMScheduler, randn_tensor

DEFAULT_NUM_INFERENCE_STEPS = 50


@dataclass
class ImagePipelineOutput:
    images: Union[List, np.ndarray]


class ConditionalDDIMPipeline:
    def __init__(self, unet, scheduler):
        # make sure scheduler can always be converted to DDIM (pipeline:44-45)
        scheduler = DDIMScheduler.from_config(scheduler.config)
        self.unet = unet
        self.scheduler = scheduler
        self._progress_bar_config = {}
        self.fused = os.environ.get("PHENDIFF_B200_FUSED", "1") != "0"

    # -- diffusers DiffusionPipeline surface used by the reference -------------------------------------
    @property
    def components(self):
        return {"unet": self.unet, "scheduler": self.scheduler}

    @property
    def device(self) -> torch.device:
        return self.unet.device

    @property
    def _execution_device(self):
        return self.device

    def to(self, *args, **kwargs):
        kwargs.pop("silence_dtype_warnings", None)
        self.unet.to(*args, **kwargs)
        return self

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def progress_bar(self, iterable=None, total=None):
        cfg = dict(self._progress_bar_config)
        if cfg.get("disable", False) or os.environ.get("PHENDIFF_B200_PROGRESS", "0") == "0":
            return iterable if iterable is not None else range(total)
        from tqdm.auto import tqdm

        return tqdm(iterable, **cfg) if iterable is not None else tqdm(total=total, **cfg)

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image

        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype("uint8")
        if images.shape[-1] == 1:
            return [Image.fromarray(image.squeeze(), mode="L") for image in images]
        return [Image.fromarray(image) for image in images]

    def enable_model_cpu_offload(self, gpu_id=0):
        raise _lib.PhenDiffB200Error("CPU offload is out of scope: phendiff_b200 keeps the model resident in HBM (no CPU path)")

    def save_pretrained(self, save_directory: str, **kw):
        os.makedirs(save_directory, exist_ok=True)
        self.unet.save_pretrained(os.path.join(save_directory, "unet"))
        self.scheduler.save_config(os.path.join(save_directory, "scheduler"))
        with open(os.path.join(save_directory, "model_index.json"), "w", encoding="utf-8") as f:
            json.dump({"_class_name": "ConditionalDDIMPipeline",
                       "unet": ["phendiff_b200", type(self.unet).__name__],
                       "scheduler": ["phendiff_b200", type(self.scheduler).__name__]}, f, indent=2)

    @classmethod
    def from_pretrained(cls, path: str, **kw):
        unet = CustomCondUNet2DModel.from_pretrained(path, subfolder="unet", **kw)
        sched = DDIMScheduler.from_config(DDIMScheduler.load_config(path, subfolder="scheduler"))
        return cls(unet, sched)

    @classmethod
    def download(cls, *a, **k):
        raise _lib.PhenDiffB200Error("there is no network access on the target systems; use from_pretrained on a local directory")

    # -- checks (pipeline:91-137) ------------------------------------------------------------------------------------
    def check_inputs(self, class_labels=None, class_emb=None, w=None, generator=None, frac_diffusion_skipped=None,
                     start_image=None) -> None:
        assert class_labels is None or (
            isinstance(class_labels, torch.Tensor) and class_labels.ndim == 1
        ), "class_labels must be a 1D tensor of shape (batch_size,) if not None."
        assert class_emb is None or (
            isinstance(class_emb, torch.Tensor) and class_emb.ndim == 2
        ), "class_emb must be a 2D tensor of shape (batch_size, emb_dim) if not None."
        assert class_labels is None or class_emb is None, "Cannot pass both class_labels and class_emb."
        batch_size = class_labels.shape[0] if class_labels is not None else class_emb.shape[0]
        assert (
            isinstance(w, float) or isinstance(w, int) or w is None or (w.ndim == 1 and batch_size == w.shape[0])
        ), "w must be a 1D tensor of shape (batch_size,) if not None and not a single int/float."
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(
                f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                f" size of {batch_size} through class conditioning. Make sure the batch size matches the length of the generators.")
        assert (frac_diffusion_skipped is not None and start_image is not None) or (
            frac_diffusion_skipped is None and start_image is None
        ), "Either pass both frac_diffusion_skipped and start_image or none of them."
        if frac_diffusion_skipped is not None:
            assert (
                isinstance(frac_diffusion_skipped, float) or isinstance(frac_diffusion_skipped, int)
            ) and 0 <= frac_diffusion_skipped <= 1, f"frac_diffusion_skipped must be a float (or int) between 0 and 1; got {frac_diffusion_skipped}."

    # -- the loop (pipeline:139-361) ---------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(
        self,
        class_labels: Optional[torch.Tensor],
        class_emb: Optional[torch.Tensor] = None,
        w: Union[int, float, torch.Tensor, None] = None,
        generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
        eta: float = 0.0,
        num_inference_steps: int = DEFAULT_NUM_INFERENCE_STEPS,
        use_clipped_model_output: Optional[bool] = None,
        output_type: Optional[str] = "pil",
        return_dict: bool = True,
        start_image: Optional[torch.Tensor] = None,
        add_forward_noise_to_image: bool = True,
        frac_diffusion_skipped: Optional[float] = None,
        guidance_eqn: Literal["imagen", "CFG"] = "imagen",
    ) -> Union[ImagePipelineOutput, Tuple]:
        self.check_inputs(class_labels, class_emb, w, generator, frac_diffusion_skipped, start_image)
        if num_inference_steps is None:
            num_inference_steps = DEFAULT_NUM_INFERENCE_STEPS
        batch_size = class_labels.shape[0] if class_labels is not None else class_emb.shape[0]
        device = self._execution_device
        if device.type != "cuda":
            raise _lib.PhenDiffB200Error("the pipeline must live on a CUDA device (pipe.to('cuda')); there is no CPU fallback")
        cfg = self.unet.config
        if isinstance(cfg.sample_size, int):
            image_shape = (batch_size, cfg.in_channels, cfg.sample_size, cfg.sample_size)
        else:
            image_shape = (batch_size, cfg.in_channels, *cfg.sample_size)

        if start_image is not None:
            image = start_image.to(device=device, dtype=torch.float32)
        else:
            image = randn_tensor(image_shape, generator, device, torch.float32)

        self.scheduler.set_timesteps(num_inference_steps)
        if frac_diffusion_skipped is not None:
            init_timestep = self.scheduler.config.num_train_timesteps * (1 - frac_diffusion_skipped)
            timesteps = self.scheduler.timesteps[self.scheduler.timesteps <= init_timestep]
        else:
            timesteps = self.scheduler.timesteps

        if add_forward_noise_to_image:
            noise = randn_tensor(image.shape, generator, device, image.dtype)
            image = self.scheduler.add_noise(image, noise, timesteps[0].repeat(batch_size))

        do_classifier_free_guidance = (
            isinstance(w, torch.Tensor)
            or (guidance_eqn == "imagen" and (isinstance(w, float) or isinstance(w, int)) and w > 1)
            or (guidance_eqn == "CFG" and (isinstance(w, float) or isinstance(w, int)) and w > 0)
        )
        if do_classifier_free_guidance and guidance_eqn not in ("imagen", "CFG"):
            raise ValueError(f"Unknown guidance equation '{guidance_eqn}'; should be 'imagen' or 'CFG'")

        if class_labels is not None:
            class_labels = class_labels.to(device)
        if class_emb is not None:
            class_emb = class_emb.to(device)

        self.unet.check_weights()   # an out-of-band `param.data.copy_` (EMA copy_to) must not sample from stale weights
        fused_ok = (self.fused_route_ok() and not do_classifier_free_guidance and eta == 0.0 and class_emb is None
                    and len(timesteps) > 0)
        fused_cfg_ok = (self.fused_route_ok() and do_classifier_free_guidance and eta == 0.0 and class_emb is None
                        and class_labels is not None and len(timesteps) > 0)
        if fused_ok:
            image = self._fused_generate(image, class_labels, timesteps, use_clipped_model_output)
        elif fused_cfg_ok:
            # one pass over 2B images per step (conditional + unconditional copies), the guidance combine and
            # the scheduler update in the conv_out epilogue (`pd_cfg_transfer`)
            if isinstance(w, torch.Tensor):
                w_dev = w.to(device=device, dtype=torch.float32).contiguous()
            else:
                w_dev = torch.full((batch_size,), float(w), dtype=torch.float32, device=device)
            image = self._fused_guided_generate(image, class_labels, w_dev, guidance_eqn, timesteps, use_clipped_model_output)
        else:
            if do_classifier_free_guidance:
                if isinstance(w, torch.Tensor):
                    w_dev = w.to(device=device, dtype=torch.float32).contiguous()
                else:
                    w_dev = torch.full((batch_size,), float(w), dtype=torch.float32, device=device)
                zero_emb = torch.zeros((batch_size, self.unet.time_embed_dim), device=device)
            for t in self.progress_bar(timesteps):
                cond_output = self.unet(sample=image, timestep=t, class_labels=class_labels, class_emb=class_emb).sample
                if do_classifier_free_guidance:
                    uncond_output = self.unet(sample=image, timestep=t, class_labels=None, class_emb=zero_emb).sample
                    guided_score = torch.empty_like(cond_output)
                    _lib.check(_lib.lib().pd_cfg_combine(
                        _lib.ptr(cond_output), _lib.ptr(uncond_output), _lib.ptr(w_dev),
                        0 if guidance_eqn == "imagen" else 1, _lib.ptr(guided_score), batch_size,
                        cond_output.numel() // batch_size, _lib.current_stream()))
                else:
                    guided_score = cond_output
                image = self.scheduler.step(guided_score, t, image, eta=eta,
                                            use_clipped_model_output=use_clipped_model_output,
                                            generator=generator).prev_sample

        image = self.postprocess(image)
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image,)
        return ImagePipelineOutput(images=image)

    # -- helpers -----------------------------------------------------------------------------------------------------
    def fused_route_ok(self) -> bool:
        """Whether the whole-path C entry point may replace the per-step loop for this model: it feeds x_t to conv_in as
        is (no `center_input_sample` rescale, cond_unet_2d.py:271-273) and indexes the class table by label."""
        return bool(self.fused and not self.unet.config.center_input_sample)

    def postprocess(self, image: torch.Tensor) -> np.ndarray:
        """(image / 2 + 0.5).clamp(0, 1) -> cpu -> NHWC numpy (pipeline:349-350), as one kernel + one D2H copy."""
        B, Cc, H, W = image.shape
        out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=image.device)
        with torch.cuda.device(image.device):
            _lib.check(_lib.lib().pd_denorm_nhwc(_lib.ptr(image.contiguous()), _lib.ptr(out), B, Cc, H, W,
                                                 _lib.current_stream()))
        return out.cpu().numpy()

    def _run_fused(self, x: torch.Tensor, src_labels, tgt_labels, steps: List[_lib.StepCoeffs], n_inv: int, n_gen: int):
        """x (B,C,H,W) fp32 CUDA, updated in place through n_inv inversion + n_gen generation steps."""
        B, _, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            self.unet.check_weights()   # an out-of-band `param.data.copy_` (EMA copy_to) must not sample from stale weights
            h = self.unet._ensure_plan(B, H, W)
            arr = (_lib.StepCoeffs * len(steps))(*steps)
            src = src_labels.to(device=dev, dtype=torch.int64).contiguous() if src_labels is not None else None
            tgt = tgt_labels.to(device=dev, dtype=torch.int64).contiguous() if tgt_labels is not None else None
            _lib.check(_lib.lib().pd_ddib_transfer(h, _lib.ptr(x), _lib.ptr(src), _lib.ptr(tgt), arr, n_inv, n_gen,
                                                   _lib.current_stream()))
        return x

    def _fused_guided_generate(self, image, class_labels, w_dev, guidance_eqn, timesteps, use_clipped_model_output):
        x = image.to(torch.float32).contiguous().clone()
        steps = [self.scheduler.step_coeffs(t, 0.0, use_clipped_model_output) for t in timesteps]
        B, _, H, W = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            h = self.unet._ensure_plan(B, H, W, guided=True)
            arr = (_lib.StepCoeffs * len(steps))(*steps)
            labels = class_labels.to(device=dev, dtype=torch.int64).contiguous()
            _lib.check(_lib.lib().pd_cfg_transfer(h, _lib.ptr(x), _lib.ptr(labels), _lib.ptr(w_dev),
                                                  0 if guidance_eqn == "imagen" else 1, arr, len(steps), _lib.current_stream()))
        return x

    def _fused_generate(self, image, class_labels, timesteps, use_clipped_model_output):
        x = image.to(torch.float32).contiguous().clone()
        steps = [self.scheduler.step_coeffs(t, 0.0, use_clipped_model_output) for t in timesteps]
        return self._run_fused(x, None, class_labels, steps, 0, len(steps))
