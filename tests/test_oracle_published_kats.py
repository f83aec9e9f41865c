"""Pins the oracle against the PUBLISHED known-answer tests of the third-party package the reference's arithmetic lives in.

The reference pins `diffusers==0.18.2` (environment.yaml:80) and calls its blocks / schedulers from
src/cond_unet_2d/cond_unet_2d.py:131-242, src/pipeline_conditional_ddim/pipeline_conditionial_ddim.py:45,248,340-347 and
src/utils_Img2Img.py:776-798.  The package cannot be installed here, but its own test-suite holds hard-coded expected values
for exactly the layers and schedulers on this path:

  * tests/models/test_layers_utils.py: EmbeddingsTests.test_sinoid_embeddings_hardcoded, ResnetBlock2DTests.test_resnet_default /
    test_restnet_with_use_in_shortcut, Upsample2DBlockTests.test_upsample_with_conv, Downsample2DBlockTests.test_downsample_with_conv,
    AttentionBlockTests.test_attention_block_default / test_attention_block_sd
  * tests/schedulers/test_scheduler_ddim.py: test_full_loop_no_noise / _with_v_prediction / _with_set_alpha_to_one /
    _with_no_set_alpha_to_one
  * tests/schedulers/test_scheduler_ddim_inverse.py (the >= 0.19 scheduler): the same four loops
  * tests/pipelines/ddim/test_ddim.py: DDIMPipelineFastTests.test_inference (whole seed-0 UNet2DModel + DDIMScheduler loop)

Provenance: the expected numbers below are those published constants, restated (the test files are not in this container);
the inputs are fully determined by `torch.manual_seed(0)` + PyTorch's default layer initialisation in diffusers' parameter
creation order, or by closed-form tensors.  An independently written restatement reproducing all 18 vectors to the printed
precision is what pins it; a wrong block (or a wrong recollection of a constant) fails here.

What stays unpinned: the 0.18.2 `DDIMInverseScheduler` index pairing (t -> t + r, `set_alpha_to_zero`), for which no
published vector could be matched — it shares `_x0_eps`, clipping and the update formula with the pinned variants.
"""
import functools

import pytest
import torch

from oracle.schedulers import OracleDDIMInverseScheduler, OracleDDIMScheduler
from oracle.unet import Attention, Downsample2D, ResnetBlock2D, Timesteps, Upsample2D

# diffusers prints 4 decimals and compares with atol 1e-3; the restatement lands within 5e-5 of every printed value
ATOL = 1.0e-4


def _corner(t):
    return t[0, -1, -3:, -3:].flatten()


def _check(out, shape, expected):
    assert tuple(out.shape) == shape
    got = _corner(out)
    assert torch.allclose(got, torch.tensor(expected), atol=ATOL, rtol=0), got.tolist()


# ---- tests/models/test_layers_utils.py -----------------------------------------------------------------------------

def test_sinusoid_embeddings_hardcoded():
    ts = torch.arange(128)
    t1 = Timesteps(64, flip_sin_to_cos=False, downscale_freq_shift=1)(ts)
    t2 = Timesteps(64, flip_sin_to_cos=True, downscale_freq_shift=0)(ts)    # the reference's setting (cond_unet_2d.py:131-137)
    assert (t1.abs() <= 1.0).all() and (t2.abs() <= 1.0).all()
    assert torch.allclose(t1[23:26, 47:50].flatten(),
                          torch.tensor([0.9646, 0.9804, 0.9892, 0.9615, 0.9787, 0.9882, 0.9582, 0.9769, 0.9872]), atol=ATOL, rtol=0)
    assert torch.allclose(t2[23:26, 47:50].flatten(),
                          torch.tensor([0.3019, 0.2280, 0.1716, 0.3146, 0.2377, 0.1790, 0.3272, 0.2474, 0.1864]), atol=ATOL, rtol=0)


@torch.no_grad()
def test_resnet_default():
    torch.manual_seed(0)
    sample, temb = torch.randn(1, 32, 64, 64), torch.randn(1, 128)
    block = ResnetBlock2D(32, 32, 128, 32, 1e-6)
    _check(block(sample, temb), (1, 32, 64, 64), [-1.9010, -0.2974, -0.8245, -1.3533, 0.8742, -0.9645, -2.0584, 1.3387, -0.4746])


@torch.no_grad()
def test_resnet_with_in_shortcut():
    torch.manual_seed(0)
    sample, temb = torch.randn(1, 32, 64, 64), torch.randn(1, 128)
    block = ResnetBlock2D(32, 32, 128, 32, 1e-6)
    block.conv_shortcut = torch.nn.Conv2d(32, 32, 1, 1, 0)     # use_in_shortcut=True: created after conv2, as in diffusers
    _check(block(sample, temb), (1, 32, 64, 64), [0.2226, -1.0791, -0.1629, 0.3659, -0.2889, -1.2376, 0.0582, 0.9206, 0.0044])


@torch.no_grad()
def test_upsample_with_conv():
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 32, 32)
    _check(Upsample2D(32)(sample), (1, 32, 64, 64), [0.7145, 1.3773, 0.3492, 0.8448, 1.0839, -0.3341, 0.5956, 0.1250, -0.4841])


@torch.no_grad()
def test_downsample_with_conv():
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 64, 64)
    _check(Downsample2D(32, 1)(sample), (1, 32, 32, 32), [0.9267, 0.5878, 0.3337, 1.2321, -0.1191, -0.3984, -0.7532, -0.0715, -0.3913])


@torch.no_grad()
def test_attention_block_default():
    torch.manual_seed(0)
    sample = torch.randn(1, 32, 64, 64)
    attn = Attention(32, 1, 32, 1e-6)           # channels=32, num_head_channels=1
    _check(attn(sample), (1, 32, 64, 64), [-1.4975, -0.0038, -0.7847, -1.4567, 1.1220, -0.8962, -1.7394, 1.1319, -0.5427])


@torch.no_grad()
def test_attention_block_sd():
    torch.manual_seed(0)
    sample = torch.randn(1, 512, 64, 64)
    attn = Attention(512, 512, 32, 1e-6)        # channels=512, one head
    _check(attn(sample), (1, 512, 64, 64), [-0.6621, -0.0156, -3.2766, 0.8025, -0.8609, 0.2820, 0.0905, -1.1179, -3.2126])


# ---- tests/schedulers/test_scheduler_ddim*.py ------------------------------------------------------------------------

def _dummy_sample_deter():
    b, c, h, w = 4, 3, 8, 8
    n = b * c * h * w
    return (torch.arange(n).reshape(c, h, w, b) / n).permute(3, 0, 1, 2)


def _dummy_model(sample, t):
    t = t.reshape(-1, *(1,) * (sample.dim() - 1)).to(sample.dtype)
    return sample * t / (t + 1)


def _full_loop(cls, **overrides):
    cfg = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    cfg.update(overrides)
    sched = cls(**cfg)
    sched.set_timesteps(10)
    x = _dummy_sample_deter()
    for t in sched.timesteps:
        x = sched.step(_dummy_model(x, t), t, x, 0.0).prev_sample
    return x.abs().sum().item(), x.abs().mean().item()


@pytest.mark.parametrize("overrides,exp_sum,exp_mean", [
    ({}, 172.0067, 0.223967),
    ({"prediction_type": "v_prediction"}, 52.5302, 0.0684),
    ({"set_alpha_to_one": True, "beta_start": 0.01}, 149.8295, 0.1951),
    ({"set_alpha_to_one": False, "beta_start": 0.01}, 149.0784, 0.1941),
])
def test_ddim_scheduler_full_loops(overrides, exp_sum, exp_mean):
    s, m = _full_loop(OracleDDIMScheduler, **overrides)
    assert abs(s - exp_sum) < 1e-2 and abs(m - exp_mean) < 1e-3, (s, m)


@pytest.mark.parametrize("overrides,exp_sum,exp_mean", [
    ({}, 671.6816, 0.8746),
    ({"prediction_type": "v_prediction"}, 1394.2185, 1.8154),
    ({"set_alpha_to_one": True, "beta_start": 0.01}, 539.9622, 0.7031),
    ({"set_alpha_to_one": False, "beta_start": 0.01}, 542.6722, 0.7066),
])
def test_ddim_inverse_scheduler_full_loops_later_release(overrides, exp_sum, exp_mean):
    s, m = _full_loop(functools.partial(OracleDDIMInverseScheduler, ">=0.19"), **overrides)
    assert abs(s - exp_sum) < 1e-2 and abs(m - exp_mean) < 1e-3, (s, m)


# ---- tests/pipelines/ddim/test_ddim.py: the whole UNet graph + the DDIM loop --------------------------------------------

@torch.no_grad()
def test_ddim_pipeline_fast_test_whole_unet():
    """DDIMPipelineFastTests.test_inference: `torch.manual_seed(0)`; UNet2DModel(block_out_channels=(32, 64), layers_per_block=2,
    sample_size=32, DownBlock2D + AttnDownBlock2D / AttnUpBlock2D + UpBlock2D); DDIMScheduler(); generator seed 0; 2 steps;
    numpy output.  The UNet is the reference's graph without the class embedding (cond_unet_2d.py builds exactly these
    blocks), so this pins the ASSEMBLY — skip connections, attention placement, time embedding, conv_in / conv_norm_out /
    conv_out — together with the scheduler loop and the (x / 2 + 0.5).clamp(0, 1) post-processing."""
    import numpy as np

    from oracle.unet import OracleCondUNet2D
    from tests.util import DDIM_FAST_TEST_SLICE, DDIM_FAST_TEST_UNET, diffusers_order_init

    unet = diffusers_order_init(OracleCondUNet2D(**DDIM_FAST_TEST_UNET).eval(), 0)
    sched = OracleDDIMScheduler()
    torch.manual_seed(0)
    image = torch.randn(1, 3, 32, 32)
    sched.set_timesteps(2)
    for t in sched.timesteps:
        image = sched.step(unet(image, t).sample, t, image, eta=0.0, use_clipped_model_output=None).prev_sample
    out = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    assert out.shape == (1, 32, 32, 3)
    got = out[0, -3:, -3:, -1].flatten()
    assert np.abs(got - np.array(DDIM_FAST_TEST_SLICE)).max() < 1e-4, got.tolist()   # diffusers' own tolerance: 1e-3


# ---- tests/schedulers/test_scheduler_ddim.py: test_variance, test_steps_offset -------------------------------------------

def test_ddim_scheduler_variance_and_steps_offset():
    """The eta > 0 variance term (`ConditionalDDIMPipeline.__call__` exposes `eta`, pipeline_conditionial_ddim.py:146) and the
    `steps_offset` timestep grid, against the constants of diffusers' test_variance / test_steps_offset."""
    cfg = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    sched = OracleDDIMScheduler(**cfg)
    for t, prev, expected in ((0, 0, 0.0), (420, 400, 0.14771), (980, 960, 0.32460), (487, 486, 0.00979), (999, 998, 0.02)):
        assert abs(float(sched._get_variance(t, prev)) - expected) < 1e-5, (t, prev)
    sched = OracleDDIMScheduler(**dict(cfg, steps_offset=1))
    sched.set_timesteps(5)
    assert torch.equal(sched.timesteps, torch.LongTensor([801, 601, 401, 201, 1]))
