"""Small kernel-level and model-level cases sized for `compute-sanitizer` (memcheck / racecheck / synccheck, SURVEY §5: the
mbarrier / TMA / TMEM protocols need it): `bash tools/gpu.sh <tag> san` runs this file under each tool and keeps the logs in
profiles/.  They also run in the ordinary `-m gpu` suite (a few seconds)."""
import pytest
import torch

from tests.test_gpu_kernels import (GN_CONV_CASES, HALO_CASES, TC_CASES, _attention_case, _conv_case, _conv_ex,
                                    test_conv_halo_fused_groupnorm as _gn_conv_check)
from tests.util import make_pair, synth_images

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [HALO_CASES[0], HALO_CASES[2], HALO_CASES[7], HALO_CASES[10]])
def test_san_conv_halo(build_lib, case):
    n, h, w, cin, cout, k, pad, kw = case
    r = _conv_ex(build_lib, 2, 2, n, h, w, cin, cout, k, pad, **kw)
    assert (r["got"] - r["ref"]).abs().max().item() <= 2e-3 * max(1.0, r["ref"].abs().max().item())


def test_san_conv_tap_stride2(build_lib):
    got, ref = _conv_case(build_lib, 1, 2, *TC_CASES[8], out_scale=1.0)   # the per-tap kernel (Downsample2D)
    assert (got - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("case", [GN_CONV_CASES[0], GN_CONV_CASES[2], GN_CONV_CASES[3]])
def test_san_conv_halo_gn(build_lib, case):
    _gn_conv_check(build_lib, case, 2)


def test_san_conv_halo_upsample_and_stats(build_lib):
    r = _conv_ex(build_lib, 2, 2, 1, 16, 16, 64, 64, 3, 1, upsample=True)
    assert (r["got"] - r["ref"]).abs().max().item() <= 2e-3 * max(1.0, r["ref"].abs().max().item())
    r = _conv_ex(build_lib, 2, 2, 2, 32, 32, 64, 128, 3, 1, addvec=True, stats_cw=4)
    assert (r["got"] - r["ref"]).abs().max().item() <= 2e-3 * max(1.0, r["ref"].abs().max().item())


@pytest.mark.parametrize("mode", ["mma_fp16", "mmav3_bf16", "mmatc_fp16", "mmatc3_fp16", "simt_fp32"])
def test_san_attention(build_lib, mode):
    n, s, c = 1, 256, 64
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(n, s, 3 * c, generator=g) * 0.8
    got, ref, bf = _attention_case(build_lib, mode, qkv, n, s, c)
    assert (got - ref).abs().max().item() <= {0: 1e-4, 1: 3e-2, 2: 4e-3}[bf]


@pytest.mark.parametrize("precision", ["fp16", "fp32"])
def test_san_unet_forward_and_ddib(build_lib, precision):
    """Whole graph through every kernel the product path launches (super_small @ 32x32, batch 2) + 2+2 fused DDIB steps."""
    from phendiff_b200 import ConditionalDDIMPipeline, DDIMScheduler, ddib_transfer
    from phendiff_b200.reference_configs import SCHEDULER_CONFIGS

    oracle, model = make_pair("super_small", 32, precision)
    x, labels = synth_images(2, 32)
    with torch.no_grad():
        ref = oracle(x, torch.tensor(7), labels).sample
    got = model(x.cuda(), torch.tensor(7), labels.cuda()).sample.cpu()
    assert (got - ref).abs().max().item() <= (1e-4 if precision == "fp32" else 1e-2)
    pipe = ConditionalDDIMPipeline(model, DDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]))
    out = ddib_transfer(pipe, x, labels, 1 - labels, 2)
    assert out.shape == (2, 3, 32, 32) and bool(torch.isfinite(out).all())


def test_san_training_step_tensor_core(build_lib):
    """One bf16 mixed-precision training step + one input-gradient backward on a two-level 128-channel UNet at 32x32: every kernel of
    the training path the sanitizer has not seen above — the tcgen05 weight-gradient kernel (3x3, 1x1, stride-2 phase view), the halo
    kernel as dgrad (incl. the sub-pixel phase mode), the mma.sync attention forward / backward, GroupNorm backward, casts."""
    from oracle import OracleCondUNet2D
    from phendiff_b200 import CustomCondUNet2DModel, DDIMScheduler
    from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS
    from phendiff_b200.training import DenoiserTrainer

    cfg = dict(DENOISER_CONFIGS["small_denoiser_config"], sample_size=32, block_out_channels=(128, 128), layers_per_block=1,
               down_block_types=("DownBlock2D", "AttnDownBlock2D"), up_block_types=("AttnUpBlock2D", "UpBlock2D"))
    torch.manual_seed(0)
    oracle = OracleCondUNet2D(**cfg)
    model = CustomCondUNet2DModel.from_config(cfg, precision="fp16")
    model.load_state_dict(oracle.state_dict())
    model = model.cuda()
    for p in oracle.parameters():          # the oracle stays on the CPU: under the sanitizer its cuDNN kernels would be instrumented too
        p.requires_grad_(True)
    sched = DDIMScheduler.from_config(SCHEDULER_CONFIGS["1k_epsilon_pred"])
    B = 1
    x, labels = synth_images(B, 32)
    g = torch.Generator().manual_seed(2)
    noise, t = torch.randn(x.shape, generator=g).cuda(), torch.tensor([321]).cuda()
    x, labels = x.cuda(), labels.cuda()
    trainer = DenoiserTrainer(model, sched, B, 32, mixed_precision="bf16", use_ema=True)
    loss = trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=t)
    counts = trainer.tensor_core_counts()
    assert counts["conv_wgrad"] >= 8, counts
    noisy = sched.add_noise(x, noise, t)
    ref_loss = torch.nn.functional.mse_loss(oracle(noisy.cpu(), t.cpu(), class_labels=labels.cpu()).sample, noise.cpu())
    ref_loss.backward()
    num = sum((gr.cpu() - dict(oracle.named_parameters())[n].grad).double().pow(2).sum().item() for n, gr in trainer.named_gradients())
    den = sum(p.grad.double().pow(2).sum().item() for p in oracle.parameters())
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * ref_loss.item()
    assert (num / den) ** 0.5 <= 1e-2
    trainer.optimizer_step()
    m = trainer.forward_only(noisy, t.float(), labels)
    gin = trainer.input_gradient(torch.ones_like(m))
    assert torch.isfinite(gin).all() and gin.abs().max().item() > 0
