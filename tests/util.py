"""Shared helpers for the parity tests (oracle on one side, the C-ABI CUDA path on the other)."""
import math

import torch

from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS


def make_pair(denoiser="small_denoiser_config", sample_size=64, precision="fp32", seed=0, device="cuda", **kw):
    """Oracle UNet (CPU fp32) and the product UNet sharing one seed-0 state_dict (SURVEY §8d synthetic inputs)."""
    from oracle import OracleCondUNet2D
    from phendiff_b200 import CustomCondUNet2DModel

    cfg = dict(DENOISER_CONFIGS[denoiser])
    cfg["sample_size"] = sample_size
    torch.manual_seed(seed)
    oracle = OracleCondUNet2D(**cfg).eval()
    model = CustomCondUNet2DModel.from_config(cfg, precision=precision, **kw)
    model.load_state_dict(oracle.state_dict())
    model = model.to(device).eval()
    return oracle, model


def synth_images(batch, size, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(batch, 3, size, size, generator=g) * 0.5).clamp(-1, 1)
    labels = torch.arange(batch) % 2
    return x, labels


def psnr(a, b, data_range=2.0):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    return 10 * math.log10(data_range**2 / mse)


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()
