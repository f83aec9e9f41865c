"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement, in plain PyTorch fp32, of the two third-party schedulers the reference's hot path
calls: `diffusers==0.18.2` `DDIMScheduler` and `DDIMInverseScheduler` (pinned by the reference at
environment.yaml:80; the package itself is absent from /root/reference and from this image, so this
is a restatement of its published algorithm).

PARITY PINNED FOR `DDIMScheduler` AND THE >= 0.19 `DDIMInverseScheduler`, UNPINNED FOR THE 0.18.2 INVERSE PAIRING:
the reference itself ships no tests, golden vectors or KATs for this path and diffusers cannot be imported
here, but diffusers' own test-suite publishes full-loop known answers (tests/schedulers/test_scheduler_ddim.py and
test_scheduler_ddim_inverse.py: epsilon / v-prediction / set_alpha_to_one on and off); this restatement reproduces all
eight (tests/test_oracle_published_kats.py).  The 0.18.2 inverse scheduler differs from the pinned >= 0.19 one only in
which (alpha-bar_t, alpha-bar_next) pair a step reads (`step`, ~10 lines) — that pairing is restated from SURVEY A.4 and no
published vector could be matched to it.  Further anchors:
  * generation:  src/pipeline_conditional_ddim/pipeline_conditionial_ddim.py:45,248,267,340-347
  * inversion:   src/utils_Img2Img.py:776-798
  * configs:     models_configs/noise_scheduler/*.json
and the derived known-answer values of SURVEY.md Appendix A.6 (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch


class _Config(dict):
    """Attribute-access dict standing in for diffusers' FrozenDict config."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas):
    # diffusers scheduling_ddim.py __init__ (0.18.2): schedule construction, all fp32 on CPU
    if trained_betas is not None:
        return torch.tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        b = [min(1 - f((i + 1) / num_train_timesteps) / f(i / num_train_timesteps), 0.999)
             for i in range(num_train_timesteps)]
        return torch.tensor(b, dtype=torch.float32)
    raise NotImplementedError(f"{beta_schedule} does is not implemented")


def rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    # Algorithm 1 of arXiv 2305.08891 as diffusers 0.18.2 states it (SURVEY A.3 "Table")
    alphas = 1.0 - betas
    alphas_cumprod = torch.cumprod(alphas, dim=0)
    s = alphas_cumprod.sqrt()
    s0 = s[0].clone()
    sT = s[-1].clone()
    s = s - sT
    s = s * (s0 / (s0 - sT))
    abar = s**2
    alphas = abar[1:] / abar[:-1]
    alphas = torch.cat([abar[0:1], alphas])
    return 1 - alphas


_DDIM_KEYS = dict(
    num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", trained_betas=None,
    clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", thresholding=False,
    dynamic_thresholding_ratio=0.995, clip_sample_range=1.0, sample_max_value=1.0,
    timestep_spacing="leading", rescale_betas_zero_snr=False,
)

_DDIM_INV_KEYS = dict(
    num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", trained_betas=None,
    clip_sample=True, set_alpha_to_zero=True, steps_offset=0, prediction_type="epsilon", clip_sample_range=1.0,
)


def _filter(cfg: dict, keys: dict) -> dict:
    # ConfigMixin.from_config keeps the constructor's own keys and silently drops the rest
    # (this is how utils_Img2Img.py:776-778 turns a DDIM config into an inverse-scheduler config)
    out = dict(keys)
    for k, v in cfg.items():
        if k in keys:
            out[k] = v
    return out


def _x0_eps(cfg, model_output, sample, alpha_prod_t, beta_prod_t):
    # step 3 of both schedulers: x0 / eps from the prediction type
    if cfg.prediction_type == "epsilon":
        x0 = (sample - beta_prod_t ** (0.5) * model_output) / alpha_prod_t ** (0.5)
        eps = model_output
    elif cfg.prediction_type == "sample":
        x0 = model_output
        eps = (sample - alpha_prod_t ** (0.5) * x0) / beta_prod_t ** (0.5)
    elif cfg.prediction_type == "v_prediction":
        x0 = (alpha_prod_t**0.5) * sample - (beta_prod_t**0.5) * model_output
        eps = (alpha_prod_t**0.5) * model_output + (beta_prod_t**0.5) * sample
    else:
        raise ValueError(f"prediction_type given as {cfg.prediction_type} must be one of `epsilon`, `sample`, or `v_prediction`")
    return x0, eps


class OracleDDIMScheduler:
    """diffusers 0.18.2 DDIMScheduler (generation direction), SURVEY A.3."""

    order = 1

    def __init__(self, **kwargs):
        self.config = _Config(_filter(kwargs, _DDIM_KEYS))
        c = self.config
        self.betas = _betas(c.num_train_timesteps, c.beta_start, c.beta_end, c.beta_schedule, c.trained_betas)
        if c.rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c.num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config):
        return cls(**{k: v for k, v in dict(config).items() if not k.startswith("_")})

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _get_variance(self, timestep, prev_timestep):
        a = self.alphas_cumprod[timestep]
        ap = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        return ((1 - ap) / (1 - a)) * (1 - a / ap)

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        N = c.num_train_timesteps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            step_ratio = N // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            step_ratio = N / num_inference_steps
            ts = np.round(np.arange(N, 0, -step_ratio)).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        c = self.config
        timestep = int(timestep)
        prev_timestep = timestep - c.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        x0, eps = _x0_eps(c, model_output, sample, alpha_prod_t, beta_prod_t)
        if c.thresholding:
            raise NotImplementedError("dynamic thresholding is used by no shipped config (SURVEY A.3)")
        elif c.clip_sample:
            x0 = x0.clamp(-c.clip_sample_range, c.clip_sample_range)
        variance = self._get_variance(timestep, prev_timestep)
        std_dev_t = eta * variance ** (0.5)
        if use_clipped_model_output:
            eps = (sample - alpha_prod_t ** (0.5) * x0) / beta_prod_t ** (0.5)
        direction = (1 - alpha_prod_t_prev - std_dev_t**2) ** (0.5) * eps
        prev_sample = alpha_prod_t_prev ** (0.5) * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype).to(model_output.device)
            prev_sample = prev_sample + std_dev_t * variance_noise
        if not return_dict:
            return (prev_sample,)
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=x0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while len(sa.shape) < len(original_samples.shape):
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        ac = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        timesteps = timesteps.to(sample.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        sa = sa.flatten()
        sb = sb.flatten()
        while len(sa.shape) < len(sample.shape):
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * noise - sb * sample

    def __len__(self):
        return self.config.num_train_timesteps


class OracleDDIMInverseScheduler:
    """diffusers DDIMInverseScheduler, SURVEY A.4.

    variant="0.18.2" (default, the pinned release): knows neither `rescale_betas_zero_snr` nor
    `timestep_spacing` -> un-rescaled table, "leading" ascending timesteps, final alpha-bar 0.
    variant=">=0.19": the later behaviour (trailing/rescale aware, (t - r -> t) pairs), switchable.
    """

    order = 1

    def __init__(self, variant: str = "0.18.2", **kwargs):
        self.variant = variant
        if variant == "0.18.2":
            self.config = _Config(_filter(kwargs, _DDIM_INV_KEYS))
        else:
            keys = dict(_DDIM_INV_KEYS)
            keys.pop("set_alpha_to_zero")
            keys.update(set_alpha_to_one=True, timestep_spacing="leading", rescale_betas_zero_snr=False)
            self.config = _Config(_filter(kwargs, keys))
        c = self.config
        self.betas = _betas(c.num_train_timesteps, c.beta_start, c.beta_end, c.beta_schedule, c.trained_betas)
        if variant != "0.18.2" and c.rescale_betas_zero_snr:
            self.betas = rescale_zero_terminal_snr(self.betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        if variant == "0.18.2":
            self.final_alpha_cumprod = torch.tensor(0.0) if c.set_alpha_to_zero else self.alphas_cumprod[-1]
        else:
            self.initial_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, c.num_train_timesteps).copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, variant: str = "0.18.2"):
        return cls(variant=variant, **{k: v for k, v in dict(config).items() if not k.startswith("_")})

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        N = c.num_train_timesteps
        if self.variant == "0.18.2" or c.timestep_spacing == "leading":
            step_ratio = N // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * step_ratio).round().copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            step_ratio = N / num_inference_steps
            ts = np.round(np.arange(N, 0, -step_ratio)[::-1]).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output=False,
             variance_noise=None, return_dict: bool = True):
        c = self.config
        timestep = int(timestep)
        r = c.num_train_timesteps // self.num_inference_steps
        if self.variant == "0.18.2":
            nxt = timestep + r
            alpha_prod_t = self.alphas_cumprod[timestep]
            alpha_prod_t_prev = self.alphas_cumprod[nxt] if nxt < c.num_train_timesteps else self.final_alpha_cumprod
        else:
            prv = timestep - r
            alpha_prod_t = self.alphas_cumprod[prv] if prv >= 0 else self.initial_alpha_cumprod
            alpha_prod_t_prev = self.alphas_cumprod[timestep]
        beta_prod_t = 1 - alpha_prod_t
        x0, eps = _x0_eps(c, model_output, sample, alpha_prod_t, beta_prod_t)
        if c.clip_sample:
            x0 = x0.clamp(-c.clip_sample_range, c.clip_sample_range)
        direction = (1 - alpha_prod_t_prev) ** (0.5) * eps
        prev_sample = alpha_prod_t_prev ** (0.5) * x0 + direction
        if not return_dict:
            return (prev_sample, x0)
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=x0)

    def __len__(self):
        return self.config.num_train_timesteps
