"""DDIM / inverse-DDIM schedulers with the diffusers 0.18.2 call surface the reference uses
(`from_config`, `.config`, `set_timesteps`, `.timesteps`, `step(...).prev_sample/.pred_original_sample`,
`add_noise`, `get_velocity`, `.alphas_cumprod`; reference call sites: pipeline_conditionial_ddim.py:45,248,267,340-347,
utils_Img2Img.py:776-798, utils_training.py:256,420,430; semantics: SURVEY.md Appendix A.3/A.4).

Host side (this file): the alpha-bar table and timestep grid — tiny fp32 host tables, built once.
Device side: every `step` / `add_noise` / `get_velocity` on tensors is one fused CUDA kernel behind the C ABI
(`pd_ddim_step`, `pd_axpby_per_sample`); in the whole-path entry point the update is fused into the UNet's conv_out
epilogue instead (`step_coeffs`).  There is no CPU implementation here: CPU tensors raise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np
import torch

from . import _lib
from .config import ConfigMixin


def randn_tensor(shape, generator, device, dtype=torch.float32):
    """diffusers.utils.randn_tensor semantics: one generator draws the whole batch on the generator's own device; a LIST
    of generators draws sample i from generator i (shape (1, ...) each), so a per-sample seed reproduces that sample
    whatever the batch composition."""
    shape = tuple(shape)
    if isinstance(generator, (list, tuple)):
        if len(generator) != shape[0]:
            raise ValueError(f"generator list of length {len(generator)} for a batch of {shape[0]}")
        parts = [torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=dtype).to(device) for g in generator]
        return torch.cat(parts, dim=0)
    gdev = generator.device if isinstance(generator, torch.Generator) else device
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)


@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


def _make_betas(n, beta_start, beta_end, schedule, trained_betas) -> torch.Tensor:
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(beta_start**0.5, beta_end**0.5, n, dtype=torch.float32) ** 2
    if schedule == "squaredcos_cap_v2":
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return torch.tensor([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=torch.float32)
    raise NotImplementedError(f"{schedule} does is not implemented")


def _zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    # arXiv 2305.08891 Algorithm 1: shift sqrt(alpha-bar) so the last one is exactly 0, keep the first
    root = torch.cumprod(1.0 - betas, dim=0).sqrt()
    first, last = root[0].clone(), root[-1].clone()
    root = (root - last) * (first / (first - last))
    abar = root**2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    return 1 - alphas


def _timestep_int(t) -> int:
    if torch.is_tensor(t):
        return int(t.item())
    return int(t)


class _SchedulerBase(ConfigMixin):
    config_name = "scheduler_config.json"
    order = 1
    init_noise_sigma = 1.0

    def scale_model_input(self, sample, timestep=None):
        return sample

    def __len__(self):
        return self.config.num_train_timesteps

    # -- shared device helpers ----------------------------------------------------------------------------------
    def _coeffs(self, a_t: float, a_next: float, sigma: float, timestep: float, use_clipped: bool) -> _lib.StepCoeffs:
        c = self.config
        if c.prediction_type not in _lib.PD_PRED:
            raise ValueError(f"prediction_type given as {c.prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
        f32 = np.float32
        a_t, a_next, sigma = f32(a_t), f32(a_next), f32(sigma)
        co = _lib.StepCoeffs()
        co.pred_type = _lib.PD_PRED[c.prediction_type]
        co.clip = int(bool(c.clip_sample))
        co.use_clipped_model_output = int(bool(use_clipped))
        co.clip_range = float(c.get("clip_sample_range", 1.0))
        co.sqrt_alpha = float(np.sqrt(a_t))
        co.sqrt_beta = float(np.sqrt(f32(1) - a_t))
        co.sqrt_alpha_next = float(np.sqrt(a_next))
        with np.errstate(invalid="ignore"):
            co.dir_coef = float(np.sqrt(f32(1) - a_next - sigma * sigma))
        co.sigma = float(sigma)
        co.timestep = float(timestep)
        return co

    def _run_step(self, co: _lib.StepCoeffs, model_output, sample, noise=None):
        _lib.require_cuda(sample, "sample")
        _lib.require_cuda(model_output, "model_output")
        if sample.dtype != torch.float32 or model_output.dtype != torch.float32:
            raise _lib.PhenDiffB200Error("scheduler.step expects fp32 sample and model_output (x_t stays fp32, SURVEY §5)")
        x = sample.contiguous()
        m = model_output.contiguous()
        prev = torch.empty_like(x)
        x0 = torch.empty_like(x)
        nz = noise.contiguous() if noise is not None else None
        _lib.check(_lib.lib().pd_ddim_step(co, _lib.ptr(x), _lib.ptr(m), _lib.ptr(nz), _lib.ptr(prev), _lib.ptr(x0),
                                           x.numel(), _lib.current_stream()))
        return prev, x0


class DDIMScheduler(_SchedulerBase):
    """Generation direction."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_one: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 thresholding: bool = False, dynamic_thresholding_ratio: float = 0.995, clip_sample_range: float = 1.0,
                 sample_max_value: float = 1.0, timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False):
        self.register_to_config(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
            set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
            thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
            clip_sample_range=clip_sample_range, sample_max_value=sample_max_value, timestep_spacing=timestep_spacing,
            rescale_betas_zero_snr=rescale_betas_zero_snr)
        if thresholding:
            raise NotImplementedError("dynamic thresholding is not used by any shipped PhenDiff config and is not implemented")
        self.betas = _make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        if rescale_betas_zero_snr:
            self.betas = _zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self._dev_tables = {}

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        N = c.num_train_timesteps
        if num_inference_steps > N:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`: {N}")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -(N / num_inference_steps))).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self.timesteps = torch.from_numpy(ts).to(device)

    def _alpha_pair(self, timestep: int):
        prev_t = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[timestep])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step_coeffs(self, timestep, eta: float = 0.0, use_clipped_model_output=False) -> _lib.StepCoeffs:
        """Host scalars of one generation step (consumed by pd_ddim_step or the fused conv_out epilogue)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = _timestep_int(timestep)
        a_t, a_prev = self._alpha_pair(t)
        sigma = 0.0
        if eta:
            f32 = np.float32
            var = (f32(1) - f32(a_prev)) / (f32(1) - f32(a_t)) * (f32(1) - f32(a_t) / f32(a_prev))
            sigma = float(f32(eta) * np.sqrt(var))
        return self._coeffs(a_t, a_prev, sigma, t, bool(use_clipped_model_output))

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict: bool = True):
        co = self.step_coeffs(timestep, eta, use_clipped_model_output)
        noise = None
        if eta > 0:
            if variance_noise is not None and generator is not None:
                raise ValueError("Cannot pass both generator and variance_noise. Please make sure that either `generator` or `variance_noise` stays `None`.")
            noise = variance_noise
            if noise is None:
                noise = randn_tensor(model_output.shape, generator, model_output.device, model_output.dtype)
        prev, x0 = self._run_step(co, model_output, sample, noise)
        if not return_dict:
            return (prev,)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def _per_sample(self, timesteps, device):
        ac = self.alphas_cumprod
        t = timesteps.to("cpu").long().flatten()
        sa = (ac[t] ** 0.5).to(device=device, dtype=torch.float32).contiguous()
        sb = ((1 - ac[t]) ** 0.5).to(device=device, dtype=torch.float32).contiguous()
        return sa, sb

    def _axpby(self, a, b, ca, cb):
        _lib.require_cuda(a, "samples")
        a = a.contiguous().float()
        b = b.contiguous().float()
        out = torch.empty_like(a)
        B = a.shape[0]
        _lib.check(_lib.lib().pd_axpby_per_sample(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ca), _lib.ptr(cb), _lib.ptr(out), B,
                                                  a.numel() // B, _lib.current_stream()))
        return out

    def add_noise(self, original_samples, noise, timesteps):
        sa, sb = self._per_sample(timesteps, original_samples.device)
        return self._axpby(original_samples, noise, sa, sb)

    def get_velocity(self, sample, noise, timesteps):
        sa, sb = self._per_sample(timesteps, sample.device)
        return self._axpby(noise, sample, sa, -sb)


class DDIMInverseScheduler(_SchedulerBase):
    """Inversion direction.  `from_config(ddim_config)` silently drops the keys this class does not
    declare (utils_Img2Img.py:776-778).  Default behaviour is the reference's pinned diffusers 0.18.2: un-rescaled
    table, "leading" ascending timesteps, final alpha-bar 0.  `variant=">=0.19"` switches to the later semantics."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, clip_sample: bool = True,
                 set_alpha_to_zero: bool = True, steps_offset: int = 0, prediction_type: str = "epsilon",
                 clip_sample_range: float = 1.0, variant: str = "0.18.2", timestep_spacing: str = "leading",
                 rescale_betas_zero_snr: bool = False, set_alpha_to_one: bool = True):
        new = variant != "0.18.2"
        cfg = dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                   beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                   steps_offset=steps_offset, prediction_type=prediction_type, clip_sample_range=clip_sample_range)
        if new:
            cfg.update(timestep_spacing=timestep_spacing, rescale_betas_zero_snr=rescale_betas_zero_snr,
                       set_alpha_to_one=set_alpha_to_one)
        else:
            cfg.update(set_alpha_to_zero=set_alpha_to_zero)
        self.register_to_config(**cfg)
        object.__setattr__(self, "variant", variant)
        self.betas = _make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        if new and rescale_betas_zero_snr:
            self.betas = _zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        if new:
            self.initial_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        else:
            self.final_alpha_cumprod = torch.tensor(0.0) if set_alpha_to_zero else self.alphas_cumprod[-1]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps).copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kwargs):
        variant = kwargs.get("variant", "0.18.2")
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        if variant == "0.18.2":
            # the 0.18.2 class knows neither of these keys: they must not leak in from a DDIMScheduler config
            for k in ("timestep_spacing", "rescale_betas_zero_snr", "set_alpha_to_one"):
                cfg.pop(k, None)
        return super().from_config(cfg, variant=variant)

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        N = c.num_train_timesteps
        if num_inference_steps > N:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`: {N}")
        self.num_inference_steps = num_inference_steps
        if self.variant == "0.18.2" or c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (N // num_inference_steps)).round().copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(N, 0, -(N / num_inference_steps))[::-1]).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported. Please make sure to choose one of 'leading' or 'trailing'.")
        self.timesteps = torch.from_numpy(ts).to(device)

    def step_coeffs(self, timestep) -> _lib.StepCoeffs:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = _timestep_int(timestep)
        N = self.config.num_train_timesteps
        r = N // self.num_inference_steps
        if self.variant == "0.18.2":
            a_t = float(self.alphas_cumprod[t])
            a_next = float(self.alphas_cumprod[t + r]) if t + r < N else float(self.final_alpha_cumprod)
        else:
            a_t = float(self.alphas_cumprod[t - r]) if t - r >= 0 else float(self.initial_alpha_cumprod)
            a_next = float(self.alphas_cumprod[t])
        return self._coeffs(a_t, a_next, 0.0, t, False)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             variance_noise=None, return_dict: bool = True):
        prev, x0 = self._run_step(self.step_coeffs(timestep), model_output, sample)
        if not return_dict:
            return (prev, x0)
        return DDIMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)
