"""GPU parity tests, edge cases of the path: ragged / prime batches against the micro-batch planner, batch 1, extents the
tensor-core tiles do not cover (the in-CUDA fallback routes), non-square samples, more than two classes, per-sample
timesteps and `class_emb` conditioning (`timestep` scalar / 0-dim / 1-D, labels xor embedding).

Same bars as tests/test_gpu_unet.py: fp32 validation mode <= 1e-4 per-step eps, 16-bit modes <= 1e-2 (fp16), images >= 40 dB.
"""
import pytest
import torch

from tests.util import make_pair, psnr, synth_images

pytestmark = pytest.mark.gpu


def _sched(name="3k_steps_clipping_rescaling"):
    from phendiff_b200.reference_configs import SCHEDULER_CONFIGS

    return SCHEDULER_CONFIGS[name]


@pytest.mark.parametrize("batch,cap,expect_mb", [(1, 64, 1), (7, 4, 4), (6, 4, 3), (10, 64, 10), (12, 8, 6), (11, 4, 4), (13, 8, 7)])
def test_ddib_batches_against_the_microbatch_planner(build_lib, batch, cap, expect_mb):
    """The planner runs the largest divisor of the batch that fits the cap when one exists within a factor 2 of it; otherwise
    (prime batches) even micro-batches whose ragged last one runs on padded scratch copies: every image's trajectory must be
    the one it has when transferred alone in the oracle, and the padding must never leak into the real outputs."""
    from oracle import OracleDDIMScheduler, OraclePipeline, oracle_ddib
    from phendiff_b200 import ConditionalDDIMPipeline, DDIMScheduler, ddib_transfer

    oracle, model = make_pair("super_small", 32, "fp16", max_microbatch=cap)
    x, src = synth_images(batch, 32)
    tgt = 1 - src
    o_pipe = OraclePipeline(oracle, OracleDDIMScheduler.from_config(_sched()))
    pipe = ConditionalDDIMPipeline(model, DDIMScheduler.from_config(_sched()))
    got = ddib_transfer(pipe, x, src, tgt, 5).cpu()
    assert model.plan_info()["microbatch"] == expect_mb
    ref = oracle_ddib(o_pipe, x, src, tgt, 5, return_raw=True)
    assert got.shape == ref.shape == (batch, 3, 32, 32)
    for i in range(batch):
        p = psnr(ref[i], got[i])
        assert p >= 40.0, f"batch {batch} cap {cap}: image {i} PSNR {p:.1f} dB"


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("fp16", 1e-2)])
@pytest.mark.parametrize("h,w", [(40, 40), (32, 64), (48, 16), (24, 24)])
def test_forward_extents_outside_the_tensor_core_tiles(build_lib, h, w, precision, bar):
    """Extents whose levels are not multiples of the 16x8-pixel halo tile (40 -> 20 -> 10, 24 -> 12 -> 6) or are not
    square: layers the tcgen05 kernels do not tile take the in-CUDA fallback (per-tap tcgen05 or SIMT fp32 accumulate),
    never the CPU; the result must not depend on which kernel ran."""
    oracle, model = make_pair("super_small", 32, precision)
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 3, h, w, generator=g) * 0.5).clamp(-1, 1)
    labels = torch.tensor([1, 0])
    for t in (3, 2500):
        with torch.no_grad():
            ref = oracle(x, torch.tensor(t), labels).sample
        got = model(x.cuda(), torch.tensor(t), labels.cuda()).sample.cpu()
        assert got.shape == ref.shape
        err = (got - ref).abs().max().item()
        assert err <= bar, f"{h}x{w} {precision} t={t}: eps max-abs err {err:.3e}"


def test_many_classes_per_sample_timesteps_and_class_emb(build_lib):
    """num_class_embeds > 2 with repeated / missing labels (the embedding rows are de-duplicated per class on the DDIB
    route), a 1-D per-sample `timestep`, and conditioning through `class_emb` instead of labels."""
    from oracle import OracleCondUNet2D
    from phendiff_b200 import CustomCondUNet2DModel
    from phendiff_b200.reference_configs import DENOISER_CONFIGS

    cfg = dict(DENOISER_CONFIGS["super_small"], sample_size=32, num_class_embeds=5)
    torch.manual_seed(3)
    oracle = OracleCondUNet2D(**cfg).eval()
    model = CustomCondUNet2DModel.from_config(cfg, precision="fp32")
    model.load_state_dict(oracle.state_dict())
    model = model.to("cuda").eval()
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(6, 3, 32, 32, generator=g) * 0.5).clamp(-1, 1)
    labels = torch.tensor([4, 0, 4, 2, 2, 4])                      # classes 1 and 3 absent, 4 three times
    with torch.no_grad():
        ref = oracle(x, torch.tensor(700), labels).sample
    got = model(x.cuda(), torch.tensor(700), labels.cuda()).sample.cpu()
    assert (got - ref).abs().max().item() <= 1e-4
    ts = torch.tensor([0, 10, 999, 1500, 2999, 7])                  # one timestep per sample
    with torch.no_grad():
        ref = oracle(x, ts, labels).sample
    got = model(x.cuda(), ts.cuda(), labels.cuda()).sample.cpu()
    assert (got - ref).abs().max().item() <= 1e-4
    emb = torch.randn(6, model.time_embed_dim, generator=g) * 0.1   # class_emb xor class_labels (cond_unet_2d.py:268-269)
    with torch.no_grad():
        ref = oracle(x, torch.tensor(42), class_emb=emb).sample
    got = model(x.cuda(), torch.tensor(42), class_emb=emb.cuda()).sample.cpu()
    assert (got - ref).abs().max().item() <= 1e-4
    with pytest.raises(ValueError):
        model(x.cuda(), torch.tensor(42), labels.cuda(), class_emb=emb.cuda())
    with pytest.raises(ValueError):
        model(x.cuda(), torch.tensor(42))


def test_single_step_and_same_class_transfer(build_lib):
    """n = 1 (one inversion step lands exactly on eps-hat for the 0.18.2 inverse scheduler) and a same-class
    transfer, which must reconstruct the input up to the discretisation error the oracle shows too."""
    from oracle import OracleDDIMScheduler, OraclePipeline, oracle_ddib
    from phendiff_b200 import ConditionalDDIMPipeline, DDIMScheduler, ddib_transfer

    oracle, model = make_pair("super_small", 32, "fp32")
    x, src = synth_images(2, 32)
    o_pipe = OraclePipeline(oracle, OracleDDIMScheduler.from_config(_sched("1k_epsilon_pred")))
    pipe = ConditionalDDIMPipeline(model, DDIMScheduler.from_config(_sched("1k_epsilon_pred")))
    for n, tgt in ((1, 1 - src), (8, src)):
        ref = oracle_ddib(o_pipe, x, src, tgt, n, return_raw=True)
        got = ddib_transfer(pipe, x, src, tgt, n).cpu()
        ok = ~(torch.isnan(ref) | torch.isnan(got))
        assert torch.equal(torch.isnan(ref), torch.isnan(got))
        p = psnr(ref[ok], got[ok])
        assert p >= 40.0, f"n={n}: PSNR {p:.1f} dB"


def test_forward_ragged_tail_per_sample_inputs(build_lib):
    """Per-op route on a ragged batch (7 images, cap 4 -> micro-batches 4 + 3 padded to 4): per-sample timesteps with integer
    labels, then with `class_emb`; every real image must match the oracle and the caller's buffers beyond the batch stay
    untouched (the tail is computed on scratch copies)."""
    oracle, model = make_pair("super_small", 32, "fp32", max_microbatch=4)
    x, labels = synth_images(7, 32)
    t = torch.tensor([3, 250, 1000, 1999, 2999, 7, 1500])
    with torch.no_grad():
        ref = oracle(x, t, labels).sample
    got = model(x.cuda(), t.cuda(), labels.cuda()).sample.cpu()
    assert model.plan_info()["microbatch"] == 4
    assert got.shape == ref.shape == (7, 3, 32, 32)
    err = (got - ref).abs().max().item()
    assert err <= 1e-4, f"ragged forward with labels: {err:.3e}"
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(7, oracle.time_embed_dim, generator=g) * 0.1
    with torch.no_grad():
        ref = oracle(x, t, class_emb=emb).sample
    got = model(x.cuda(), t.cuda(), class_emb=emb.cuda()).sample.cpu()
    err = (got - ref).abs().max().item()
    assert err <= 1e-4, f"ragged forward with class_emb: {err:.3e}"


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_single_head_unconditional_config(build_lib, precision, bar):
    """models_configs/denoiser/orig_google_ddpm_model_denoiser.json: `attention_head_dim: null` (ONE head of dim C = 512,
    cond_unet_2d.py:176-178,192-194,222-224), six levels, no class table.  Forward parity against the oracle at 64x64 and the
    whole-path route (unconditional models ride the fused route too: one embedding row, labels ignored as the reference does)."""
    from oracle import OracleDDIMScheduler, OraclePipeline, oracle_ddib
    from phendiff_b200 import ConditionalDDIMPipeline, DDIMScheduler, ddib_transfer
    from phendiff_b200.reference_configs import SCHEDULER_CONFIGS

    oracle, model = make_pair("orig_google_ddpm_model_denoiser", 64, precision)
    assert model.class_embedding is None and model.config.attention_head_dim is None
    x, labels = synth_images(2, 64)
    for t in (3, 999):
        with torch.no_grad():
            ref = oracle(x, torch.tensor(t)).sample
        got = model(x.cuda(), torch.tensor(t)).sample.cpu()
        err = (got - ref).abs().max().item()
        print(f"[single-head {precision}] t={t}: max abs err {err:.3e} (ref max {ref.abs().max():.3f})")
        assert err <= bar, f"t={t}: {err:.3e}"
    # (v-prediction scheduler: with 3 steps of the 1k epsilon scheduler x0 = (x - sqrt(1-a) eps) / sqrt(a) amplifies the 16-bit
    # eps error 50-fold at the first step — 105 dB in the fp32 mode, 28 dB in fp16 — which says nothing about this model)
    sched = SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]
    o_pipe = OraclePipeline(oracle, OracleDDIMScheduler.from_config(sched))
    pipe = ConditionalDDIMPipeline(model, DDIMScheduler.from_config(sched))
    assert pipe.fused_route_ok()
    ref = oracle_ddib(o_pipe, x, labels, 1 - labels, 3, return_raw=True)
    got = ddib_transfer(pipe, x, labels, 1 - labels, 3).cpu()
    ok = ~(torch.isnan(ref) | torch.isnan(got))
    p = psnr(ref[ok], got[ok])
    print(f"[single-head {precision}] DDIB 3+3 steps PSNR {p:.1f} dB")
    assert p >= 40.0


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("fp16", 1e-2)])
def test_sd21_width_denoiser_config(precision, tol):
    """models_configs/denoiser/SD_2-1_config.json: the fourth shipped denoiser JSON — the same CondUNet2D class at SD-2.1 widths (320 / 640 /
    1280 / 1280: 64-channel N tiles, 10-channel GroupNorm groups with a 2-channel statistics chunk, attention at three levels incl.
    the input resolution; 642 M parameters).  Forward vs the oracle at 32x32 (the tcgen05 kernels take the 32- and 16-pixel levels, the
    8- and 4-pixel levels fall back to the CUDA-core kernels)."""
    oracle, model = make_pair("SD_2-1_config", 32, precision)
    oracle = oracle.cuda()
    x, labels = synth_images(1, 32)
    x, labels = x.cuda(), labels.cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for t in (3, 999):
        with torch.no_grad():
            ref = oracle(x, torch.tensor(t, device="cuda"), labels).sample
        got = model(x, torch.tensor(t), labels).sample
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        print(f"[SD-2.1 widths {precision}] t={t}: max err / max|ref| = {err:.3e}; plan {model.plan_info()}")
        assert err <= tol


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-2), ("bf16", 4e-2)])
def test_pair_kernel_for_linear_layers(precision, tol, monkeypatch):
    """PHENDIFF_B200_LIN2CTA=1: the q/k/v and out-proj projections on the cta_group::2 kernel (thread-block cluster of two SMs, M = 256
    per MMA, each CTA loading half of the weight tile, multicast commits, remote mbarrier arrivals).  Same parity bar as the default
    route, and the launch counter proves the kernel ran."""
    from phendiff_b200 import _lib

    monkeypatch.setenv("PHENDIFF_B200_LIN2CTA", "1")
    before = _lib.lib().pd_debug_pair_kernel_launches()
    oracle, model = make_pair("small_denoiser_config", 64, precision)
    x, labels = synth_images(4, 64)
    with torch.no_grad():
        ref = oracle(x, torch.tensor(1500), labels).sample
    got = model(x.cuda(), torch.tensor(1500), labels.cuda()).sample.cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    ran = _lib.lib().pd_debug_pair_kernel_launches() - before
    print(f"[pair kernel {precision}] max err / max|ref| = {err:.3e}; cta_group::2 launches {ran}")
    assert ran >= 12       # 6 attention blocks x (qkv + out-proj)
    assert err <= tol
