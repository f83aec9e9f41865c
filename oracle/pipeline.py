"""ORACLE (test infrastructure, never shipped on the product path).

Restatement of the reference's denoising loop and DDIB drivers on top of the oracle UNet/schedulers:
  * ConditionalDDIMPipeline.__call__  src/pipeline_conditional_ddim/pipeline_conditionial_ddim.py:139-361
  * _inversion                        src/utils_Img2Img.py:763-800
  * _ddib                             src/utils_Img2Img.py:566-612
Parity status: blocks and schedulers are pinned against diffusers' published known answers, the loop itself follows the
reference lines cited above (see oracle/unet.py, oracle/schedulers.py, tests/test_oracle_published_kats.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from .schedulers import OracleDDIMInverseScheduler, OracleDDIMScheduler

DEFAULT_NUM_INFERENCE_STEPS = 50


class OraclePipeline:
    def __init__(self, unet, scheduler):
        # pipeline:44-45 — the scheduler is always re-created as a DDIMScheduler from the given config
        self.unet = unet
        self.scheduler = OracleDDIMScheduler.from_config(scheduler.config)

    @property
    def device(self):
        return self.unet.device

    def check_inputs(self, class_labels, class_emb, w, generator, frac_diffusion_skipped, start_image):
        # pipeline:91-137
        assert class_labels is None or (isinstance(class_labels, torch.Tensor) and class_labels.ndim == 1)
        assert class_emb is None or (isinstance(class_emb, torch.Tensor) and class_emb.ndim == 2)
        assert class_labels is None or class_emb is None
        batch_size = class_labels.shape[0] if class_labels is not None else class_emb.shape[0]
        assert isinstance(w, (float, int)) or w is None or (w.ndim == 1 and batch_size == w.shape[0])
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError("generator list length != batch size")
        assert (frac_diffusion_skipped is not None and start_image is not None) or (
            frac_diffusion_skipped is None and start_image is None)
        if frac_diffusion_skipped is not None:
            assert isinstance(frac_diffusion_skipped, (float, int)) and 0 <= frac_diffusion_skipped <= 1

    @torch.no_grad()
    def __call__(self, class_labels, class_emb=None, w=None, generator=None, eta=0.0,
                 num_inference_steps=DEFAULT_NUM_INFERENCE_STEPS, use_clipped_model_output=None,
                 output_type="numpy", return_dict=True, start_image=None, add_forward_noise_to_image=True,
                 frac_diffusion_skipped=None, guidance_eqn="imagen", return_raw=False, trace=None):
        self.check_inputs(class_labels, class_emb, w, generator, frac_diffusion_skipped, start_image)
        if num_inference_steps is None:
            num_inference_steps = DEFAULT_NUM_INFERENCE_STEPS
        batch_size = class_labels.shape[0] if class_labels is not None else class_emb.shape[0]
        device = self.device
        ss = self.unet.config.sample_size
        shape = (batch_size, self.unet.config.in_channels, ss, ss) if isinstance(ss, int) else \
            (batch_size, self.unet.config.in_channels, *ss)
        if start_image is not None:
            image = start_image
        else:
            image = torch.randn(shape, generator=generator, dtype=self.unet.dtype).to(device)
        self.scheduler.set_timesteps(num_inference_steps)
        if frac_diffusion_skipped is not None:
            init_timestep = self.scheduler.config.num_train_timesteps * (1 - frac_diffusion_skipped)
            timesteps = self.scheduler.timesteps[self.scheduler.timesteps <= init_timestep]
        else:
            timesteps = self.scheduler.timesteps
        if add_forward_noise_to_image:
            noise = torch.randn(image.shape, generator=generator, dtype=image.dtype).to(device)
            image = self.scheduler.add_noise(image, noise, timesteps[0].repeat(batch_size))
        do_cfg = isinstance(w, torch.Tensor) or (guidance_eqn == "imagen" and isinstance(w, (float, int)) and w > 1) \
            or (guidance_eqn == "CFG" and isinstance(w, (float, int)) and w > 0)
        for t in timesteps:
            cond = self.unet(sample=image, timestep=t, class_labels=class_labels, class_emb=class_emb).sample
            if do_cfg:
                uncond = self.unet(sample=image, timestep=t, class_labels=None,
                                   class_emb=torch.zeros((batch_size, self.unet.time_embed_dim)).to(device)).sample
                ww = w.view(-1, 1, 1, 1) if isinstance(w, torch.Tensor) else w
                if guidance_eqn == "imagen":
                    guided = uncond + ww * (cond - uncond)
                elif guidance_eqn == "CFG":
                    guided = cond + ww * (cond - uncond)
                else:
                    raise ValueError(f"Unknown guidance equation '{guidance_eqn}'; should be 'imagen' or 'CFG'")
            else:
                guided = cond
            if trace is not None:   # (t, x_t, model output) per step, for teacher-forced per-step comparisons
                trace.append((int(t), image.clone(), guided.clone()))
            image = self.scheduler.step(guided, t, image, eta=eta, use_clipped_model_output=use_clipped_model_output,
                                        generator=generator).prev_sample
        if return_raw:
            return image
        image = (image / 2 + 0.5).clamp(0, 1)
        image = image.cpu().permute(0, 2, 3, 1).numpy()
        if not return_dict:
            return (image,)
        return SimpleNamespace(images=image)


@torch.no_grad()
def oracle_inversion(pipe, input_images, class_labels, num_inference_steps, variant="0.18.2", trace=None):
    # utils_Img2Img.py:763-800
    gauss = input_images.clone().detach()
    inv = OracleDDIMInverseScheduler.from_config(pipe.scheduler.config, variant=variant)
    inv.set_timesteps(num_inference_steps)
    for t in inv.timesteps:
        model_output = pipe.unet(gauss, t, class_labels).sample
        if trace is not None:
            trace.append((int(t), gauss.clone(), model_output.clone()))
        gauss = inv.step(model_output, t, gauss).prev_sample
    return gauss


@torch.no_grad()
def oracle_ddib(pipe, clean_images, orig_class_labels, target_class_labels, num_inference_steps,
                variant="0.18.2", return_raw=False, trace_inv=None, trace_gen=None):
    # utils_Img2Img.py:566-612 (ConditionalDDIMPipeline branch)
    inverted = oracle_inversion(pipe, clean_images, orig_class_labels, num_inference_steps, variant, trace=trace_inv)
    out = pipe(class_labels=target_class_labels, w=0, num_inference_steps=num_inference_steps,
               start_image=inverted, add_forward_noise_to_image=False, frac_diffusion_skipped=0,
               return_raw=return_raw, trace=trace_gen)
    return out if return_raw else out.images


def oracle_custom_guided_generation(pipe, input_images, target_class_labels, p, guidance_loss_scale, num_inference_steps, trace=None):
    # utils_Img2Img.py:701-760 (_custom_guided_generation): per step, gradient of the per-image L_p distance between the predicted
    # clean image and `input_images` w.r.t. the current images (through the UNet and through the scheduler's x0 formula), a gradient
    # step on the images, then the scheduler update with the model output computed BEFORE the gradient step
    images = input_images.clone().detach()
    pipe.scheduler.set_timesteps(num_inference_steps)
    for t in pipe.scheduler.timesteps:
        images = images.detach().requires_grad_()
        model_output = pipe.unet(images, t, target_class_labels).sample
        x0 = pipe.scheduler.step(model_output, t, images).pred_original_sample
        losses = torch.linalg.vector_norm(x0 - input_images, dim=(1, 2, 3), ord=p)          # Lp_loss, utils_Img2Img.py:245-270
        guidance_grad = torch.autograd.grad([losses[i] for i in range(len(input_images))], images)[0]
        if trace is not None:
            trace.append((int(t), images.detach().clone(), model_output.detach().clone(), losses.detach().clone(), guidance_grad.clone()))
        images = images.detach() - guidance_loss_scale * guidance_grad
        images = pipe.scheduler.step(model_output.detach(), t, images).prev_sample
    return images.detach()


def oracle_linear_interp_custom_guidance_inverted_start(pipe, clean_images, orig_class_labels, target_class_labels, p, guidance_loss_scale,
                                                        num_inference_steps, variant="0.18.2"):
    # utils_Img2Img.py:651-698 (ConditionalDDIMPipeline branch): inversion, then guided generation from (and towards!) the inverted Gaussian
    inverted = oracle_inversion(pipe, clean_images, orig_class_labels, num_inference_steps, variant)
    return oracle_custom_guided_generation(pipe, inverted, target_class_labels, p, guidance_loss_scale, num_inference_steps)
