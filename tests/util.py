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


def diffusers_order_init(unet, seed=0):
    """Re-draws the oracle UNet's parameters from `torch.manual_seed(seed)` in the order diffusers' `UNet2DModel.__init__` creates
    its modules (conv_in, time embedding, [class embedding], per down block resnet_i / attention_i interleaved then the
    downsampler, mid block resnet / attention / resnet, up blocks likewise, conv_out; inside a ResnetBlock2D conv1,
    time_emb_proj, conv2, conv_shortcut; inside an attention block q, k, v, out).  With PyTorch's default initialisers this
    reproduces the weights diffusers' own tests get from `torch.manual_seed(0)` + constructing the model, which is what makes
    their published output vectors usable here (tests/test_oracle_published_kats.py)."""
    torch.manual_seed(seed)

    def resnet(r):
        for m in (r.conv1, r.time_emb_proj, r.conv2):
            m.reset_parameters()
        if r.conv_shortcut is not None:
            r.conv_shortcut.reset_parameters()

    def attn(a):
        for m in (a.to_q, a.to_k, a.to_v, a.to_out[0]):
            m.reset_parameters()

    def block(b, samplers):
        for i, r in enumerate(b.resnets):
            resnet(r)
            if b.has_attn:
                attn(b.attentions[i])
        if samplers is not None:
            samplers[0].conv.reset_parameters()

    unet.conv_in.reset_parameters()
    unet.time_embedding.linear_1.reset_parameters()
    unet.time_embedding.linear_2.reset_parameters()
    if unet.class_embedding is not None:
        unet.class_embedding.reset_parameters()
    for b in unet.down_blocks:
        block(b, b.downsamplers)
    mid = unet.mid_block
    resnet(mid.resnets[0])
    if mid.attentions[0] is not None:
        attn(mid.attentions[0])
    resnet(mid.resnets[1])
    for b in unet.up_blocks:
        block(b, b.upsamplers)
    unet.conv_out.reset_parameters()
    return unet


# diffusers tests/pipelines/ddim/test_ddim.py::DDIMPipelineFastTests::test_inference: the dummy UNet2DModel and the published
# corner of the generated image (image[0, -3:, -3:, -1], 2 inference steps, default DDIMScheduler, generator seed 0)
DDIM_FAST_TEST_UNET = dict(block_out_channels=(32, 64), layers_per_block=2, sample_size=32, in_channels=3, out_channels=3,
                           down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))
DDIM_FAST_TEST_SLICE = [1.000e00, 5.717e-01, 4.717e-01, 1.000e00, 0.000e00, 1.000e00, 3.000e-04, 0.000e00, 9.000e-04]
