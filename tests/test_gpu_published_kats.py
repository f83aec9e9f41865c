"""GPU: the PRODUCT schedulers (host coefficient tables + `pd_ddim_step` on the device, through the C ABI) against the
full-loop known answers published in diffusers' own test-suite (tests/schedulers/test_scheduler_ddim.py and
test_scheduler_ddim_inverse.py) — the same constants tests/test_oracle_published_kats.py pins the oracle with, and the
same tolerances diffusers uses (|sum| within 1e-2, |mean| within 1e-3).  The "model" of those tests is the closed form
x * t / (t + 1); it is evaluated with torch on the device and is test scaffolding, not the path under test.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dummy_sample_deter():
    b, c, h, w = 4, 3, 8, 8
    n = b * c * h * w
    return (torch.arange(n).reshape(c, h, w, b) / n).permute(3, 0, 1, 2).contiguous()


def _full_loop(sched):
    sched.set_timesteps(10)
    x = _dummy_sample_deter().cuda()
    for t in sched.timesteps:
        tf = float(t)
        m = x * tf / (tf + 1.0)
        x = sched.step(m, t, x, 0.0).prev_sample
    x = x.cpu()
    return x.abs().sum().item(), x.abs().mean().item()


_BASE = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)


@pytest.mark.parametrize("overrides,exp_sum,exp_mean", [
    ({}, 172.0067, 0.223967),
    ({"prediction_type": "v_prediction"}, 52.5302, 0.0684),
    ({"set_alpha_to_one": True, "beta_start": 0.01}, 149.8295, 0.1951),
    ({"set_alpha_to_one": False, "beta_start": 0.01}, 149.0784, 0.1941),
])
def test_product_ddim_scheduler_published_full_loops(build_lib, overrides, exp_sum, exp_mean):
    from phendiff_b200 import DDIMScheduler

    s, m = _full_loop(DDIMScheduler(**dict(_BASE, **overrides)))
    assert abs(s - exp_sum) < 1e-2 and abs(m - exp_mean) < 1e-3, (s, m)


@pytest.mark.parametrize("overrides,exp_sum,exp_mean", [
    ({}, 671.6816, 0.8746),
    ({"prediction_type": "v_prediction"}, 1394.2185, 1.8154),
    ({"set_alpha_to_one": True, "beta_start": 0.01}, 539.9622, 0.7031),
    ({"set_alpha_to_one": False, "beta_start": 0.01}, 542.6722, 0.7066),
])
def test_product_ddim_inverse_scheduler_published_full_loops(build_lib, overrides, exp_sum, exp_mean):
    from phendiff_b200 import DDIMInverseScheduler

    s, m = _full_loop(DDIMInverseScheduler(variant=">=0.19", **dict(_BASE, **overrides)))
    assert abs(s - exp_sum) < 1e-2 and abs(m - exp_mean) < 1e-3, (s, m)


@pytest.mark.parametrize("precision,atol", [("fp32", 1e-3), ("fp16", 2e-2)])
def test_product_unet_and_ddim_loop_published_pipeline_vector(build_lib, precision, atol):
    """diffusers tests/pipelines/ddim/test_ddim.py::DDIMPipelineFastTests::test_inference through the PRODUCT: the CUDA UNet
    (unconditional here: the reference's graph without the class embedding) with the seed-0 weights diffusers' constructor
    draws, two DDIM steps on the device, de-normalisation kernel; compared with the published corner of the image at
    diffusers' own tolerance (fp32 validation mode) / a 16-bit bound (fp16 storage)."""
    import numpy as np

    from oracle.unet import OracleCondUNet2D
    from phendiff_b200 import CustomCondUNet2DModel, DDIMScheduler
    from tests.util import DDIM_FAST_TEST_SLICE, DDIM_FAST_TEST_UNET, diffusers_order_init

    oracle = diffusers_order_init(OracleCondUNet2D(**DDIM_FAST_TEST_UNET).eval(), 0)
    model = CustomCondUNet2DModel.from_config(dict(DDIM_FAST_TEST_UNET), precision=precision)
    model.load_state_dict(oracle.state_dict())
    model = model.to("cuda").eval()
    sched = DDIMScheduler()
    torch.manual_seed(0)
    image = torch.randn(1, 3, 32, 32).cuda()
    sched.set_timesteps(2)
    with torch.no_grad():
        for t in sched.timesteps:
            image = sched.step(model(image, t).sample, t, image, eta=0.0, use_clipped_model_output=None).prev_sample
    out = (image / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy()
    got = out[0, -3:, -3:, -1].flatten()
    assert np.abs(got - np.array(DDIM_FAST_TEST_SLICE)).max() < atol, got.tolist()
