"""Training step against torch.autograd on the oracle.

The oracle UNet is plain PyTorch, so `loss.backward()` on it IS the reference's `accelerator.backward(loss)` for this model
(src/utils_training.py:436).  Every parameter gradient of the CUDA path (`pd_train_step_grad`) is compared with autograd's, for
the three losses of utils_training.py:415-433 and for the unconditional pass of classifier-free-guidance training (:508-516);
clip + AdamW + EMA (`pd_adamw_step`) against `torch.nn.utils.clip_grad_norm_` + `torch.optim.AdamW` + the EMAModel recurrence.
Tolerances are fp32 round-off of two different summation orders: 2e-3 of each gradient's own max (most land below 1e-4)."""
import math

import pytest
import torch

from phendiff_b200.reference_configs import SCHEDULER_CONFIGS
from tests.util import make_pair, synth_images

pytestmark = pytest.mark.gpu


def _oracle_loss(oracle, sched, clean, labels, noise, timesteps, ptype, uncond):
    import torch.nn.functional as F

    noisy = sched.add_noise(clean, noise, timesteps)
    if uncond:
        out = oracle(noisy, timesteps, class_labels=None, class_emb=torch.zeros(clean.shape[0], oracle.time_embed_dim, device=clean.device)).sample
    else:
        out = oracle(noisy, timesteps, class_labels=labels).sample
    if ptype == "epsilon":
        return F.mse_loss(out, noise), out
    if ptype == "sample":
        a = sched.alphas_cumprod.to(clean.device)[timesteps].view(-1, 1, 1, 1)
        return ((a / (1 - a)) * F.mse_loss(out, clean, reduction="none")).mean(), out
    vel = sched.get_velocity(clean, noise, timesteps)
    return F.mse_loss(out, vel), out


def _setup(denoiser, size, B, sched_name, seed=0):
    from oracle import OracleDDIMScheduler
    from phendiff_b200 import DDIMScheduler
    from phendiff_b200.training import DenoiserTrainer

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    oracle, model = make_pair(denoiser, size, "fp32", seed=seed)
    oracle = oracle.cuda().train()
    for p in oracle.parameters():
        p.requires_grad_(True)
    cfg = dict(SCHEDULER_CONFIGS["1k_epsilon_pred" if sched_name == "sample" else sched_name])
    if sched_name == "sample":
        cfg["prediction_type"] = "sample"
    osched = OracleDDIMScheduler.from_config(cfg)
    sched = DDIMScheduler.from_config(cfg)
    return oracle, model, osched, sched, DenoiserTrainer


def _inputs(B, size, seed=5):
    x, labels = synth_images(B, size, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    noise = torch.randn(x.shape, generator=g)
    timesteps = torch.randint(0, 1000, (B,), generator=g)
    return x.cuda(), labels.cuda(), noise.cuda(), timesteps.cuda()


def _compare_grads(oracle, trainer, tol=2e-3):
    ref = {n: p.grad for n, p in oracle.named_parameters()}
    worst = (0.0, None)
    for name, g in trainer.named_gradients():
        r = ref[name]
        if r is None:                         # parameter unused in this pass (class table in the unconditional pass)
            assert g.abs().max().item() == 0.0, name
            continue
        scale = max(r.abs().max().item(), 1e-6)
        err = (g - r).abs().max().item() / scale
        if err > worst[0]:
            worst = (err, name)
        assert err <= tol, f"{name}: max|dg| / max|g| = {err:.3e} (max|g| = {scale:.3e})"
    return worst


@pytest.mark.parametrize("sched_name,uncond", [("1k_epsilon_pred", False), ("1k_epsilon_pred", True),
                                               ("3k_steps_clipping_rescaling", False), ("sample", False)])
def test_gradients_match_autograd(sched_name, uncond):
    B, size = 3, 32
    oracle, model, osched, sched, Trainer = _setup("super_small", size, B, sched_name)
    ptype = sched.config.prediction_type
    x, labels, noise, timesteps = _inputs(B, size)
    if sched_name == "sample":
        timesteps = timesteps.clamp(min=50)          # SNR weights explode at t -> 0; keep the comparison well scaled
    loss_ref, out_ref = _oracle_loss(oracle, osched, x, labels, noise, timesteps, ptype, uncond)
    loss_ref.backward()
    trainer = Trainer(model, sched, B, size)
    loss, out = trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps, do_unconditional_pass=uncond,
                                               return_model_output=True)
    assert (out - out_ref).abs().max().item() <= 2e-4 * max(1.0, out_ref.abs().max().item())
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item()) + 1e-7
    worst = _compare_grads(oracle, trainer)
    print(f"{sched_name} uncond={uncond}: loss {loss.item():.6f} vs {loss_ref.item():.6f}; worst gradient {worst[1]} {worst[0]:.2e}")
    if uncond:
        g = dict(trainer.named_gradients())["class_embedding.weight"]
        assert g.abs().max().item() == 0.0


def test_gradients_small_denoiser_64():
    """The trainable config the reference ships for 128^2 data (small_denoiser: three attention levels), at 64^2."""
    B, size = 2, 64
    oracle, model, osched, sched, Trainer = _setup("small_denoiser_config", size, B, "1k_epsilon_pred")
    x, labels, noise, timesteps = _inputs(B, size, seed=9)
    loss_ref, _ = _oracle_loss(oracle, osched, x, labels, noise, timesteps, "epsilon", False)
    loss_ref.backward()
    trainer = Trainer(model, sched, B, size)
    loss = trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps)
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    worst = _compare_grads(oracle, trainer)
    print(f"small_denoiser 64^2: worst gradient {worst[1]} {worst[0]:.2e}")


def test_gradients_accumulate():
    B, size = 2, 32
    oracle, model, osched, sched, Trainer = _setup("super_small", size, B, "1k_epsilon_pred")
    x, labels, noise, timesteps = _inputs(B, size)
    trainer = Trainer(model, sched, B, size)
    trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps)
    g1 = trainer.grads.clone()
    trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps)
    assert torch.allclose(trainer.grads, 2 * g1, rtol=1e-5, atol=1e-8)
    trainer.zero_grad()
    assert trainer.grads.abs().max().item() == 0.0


@pytest.mark.parametrize("max_norm,use_ema", [(1.0, True), (0.0, False), (1e-3, True)])
def test_clip_adamw_ema_matches_torch(max_norm, use_ema):
    from phendiff_b200 import _lib
    from phendiff_b200.training import ema_decay_at

    torch.manual_seed(3)
    n = 100_003
    p0 = torch.randn(n, device="cuda")
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    ema = p.clone() if use_ema else None
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.95, 0.999), weight_decay=1e-2, eps=1e-8)
    ref_ema = p0.clone()
    scratch, norm = torch.zeros(1185, device="cuda"), torch.zeros(1, device="cuda")      # PD_ADAMW_SCRATCH_FLOATS
    for step in range(1, 6):
        g = torch.randn(n, device="cuda") * (0.1 * step)
        ref.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref], max_norm) if max_norm > 0 else g.norm()
        opt.step()
        decay = ema_decay_at(step, max_decay=0.9999, inv_gamma=1.0, power=0.75)
        ref_ema.sub_((1 - decay) * (ref_ema - ref.data))
        _lib.check(_lib.lib().pd_adamw_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(ema), n, 3e-3, 0.95, 0.999, 1e-8,
                                           1e-2, step, max_norm, decay, _lib.ptr(scratch), _lib.ptr(norm), _lib.current_stream()))
        assert abs(norm.item() - ref_norm.item()) <= 1e-5 * ref_norm.item()
        assert (p - ref.data).abs().max().item() <= 2e-6, step
        if use_ema:
            assert (ema - ref_ema).abs().max().item() <= 2e-6, step


def test_ema_decay_schedule():
    from phendiff_b200.training import ema_decay_at

    assert ema_decay_at(1) == 0.0
    assert math.isclose(ema_decay_at(2, inv_gamma=1.0, power=0.75), 1 - 2 ** -0.75)
    assert ema_decay_at(10 ** 9) == 0.9999
    assert math.isclose(ema_decay_at(5, use_ema_warmup=False), 5 / 14)


def test_training_reduces_loss_and_inference_sees_new_weights():
    """A few optimizer steps on one fixed batch: the loss goes down, the module's parameters (views of the flat vector) move, and
    the inference route re-reads them."""
    B, size = 4, 32
    oracle, model, osched, sched, Trainer = _setup("super_small", size, B, "1k_epsilon_pred")
    x, labels, noise, timesteps = _inputs(B, size, seed=11)
    trainer = Trainer(model, sched, B, size, learning_rate=2e-3, use_ema=True)
    w0 = model.conv_in.weight.detach().clone()
    t0 = torch.full((B,), 500, device="cuda")
    before = model(x, t0, class_labels=labels).sample.clone()
    losses = [trainer.step(x, labels, noise=noise, timesteps=timesteps, do_unconditional_pass=False).item() for _ in range(12)]
    assert losses[-1] < 0.7 * losses[0], losses
    assert trainer.global_step == 12 and trainer.grads.abs().max().item() == 0.0
    assert (model.conv_in.weight - w0).abs().max().item() > 0
    after = model(x, t0, class_labels=labels).sample
    assert (after - before).abs().max().item() > 1e-4
    # parity of the updated weights with the same 12 steps of torch on the oracle
    opt = torch.optim.AdamW(oracle.parameters(), lr=2e-3, betas=(0.95, 0.999), weight_decay=1e-6, eps=1e-8)
    for _ in range(12):
        loss_ref, _ = _oracle_loss(oracle, osched, x, labels, noise, timesteps, "epsilon", False)
        opt.zero_grad()
        loss_ref.backward()
        torch.nn.utils.clip_grad_norm_(oracle.parameters(), 1.0)
        opt.step()
    assert abs(loss_ref.item() - losses[-1]) <= 0.05 * abs(loss_ref.item()), (loss_ref.item(), losses[-1])
    ema = trainer.ema_state()
    assert ema is not None and (ema["conv_in.weight"] - model.conv_in.weight).abs().max().item() > 0


def _rel_l2(g, r):
    return ((g - r).norm() / r.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("tc_mask,B,size", [(1, 2, 64), (2, 2, 64), (4, 2, 64), (8, 2, 64), (15, 2, 64), (31, 2, 64), (63, 2, 64), (127, 2, 64), (127, 1, 128)])
def test_bf16_tensor_core_gradients(tc_mask, B, size, monkeypatch):
    """mixed_precision="bf16": convolutions on the tcgen05 kernels with bf16 operands (forward = 1, dgrad = 2, wgrad = 4; 8 = attention on mma.sync; 16 = GroupNorm outputs in bf16 only + fused
    q/k/v GEMM; 32 = conv_in / conv_out gradients as zero-padded tensor-core GEMMs; 64 = GroupNorm statistics
    from the conv's forward epilogue; all = 127)
    against fp32 autograd on the oracle.  bf16 operand rounding is 2^-9 relative per element; a gradient tensor is a sum of many
    such products, so the bar is on the tensor as a whole: relative L2 error <= 3e-2 per parameter tensor (measured ~5e-3) and
    <= 1e-2 over the whole gradient vector (the VERDICT's "bf16 <= 1e-2 rel")."""
    monkeypatch.setenv("PHENDIFF_B200_TRAIN_TC", str(tc_mask))
    # (127, 1, 128): the resolution of the headline / training configs — 128-pixel rows are one K block of the weight-gradient kernel
    # (130-row boxes), the stride-2 phase view at 64-pixel output rows, conv_in / conv_out padded GEMMs at full width
    oracle, model, osched, sched, Trainer = _setup("small_denoiser_config", size, B, "1k_epsilon_pred")
    x, labels, noise, timesteps = _inputs(B, size, seed=9)
    loss_ref, out_ref = _oracle_loss(oracle, osched, x, labels, noise, timesteps, "epsilon", False)
    loss_ref.backward()
    trainer = Trainer(model, sched, B, size, mixed_precision="bf16")
    loss, out = trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps, return_model_output=True)
    counts = trainer.tensor_core_counts()
    if tc_mask & 1:
        assert counts["conv_forward"] > 50, counts
    if tc_mask & 4:
        assert counts["conv_wgrad"] > 50, counts
    fwd_err = (out - out_ref).abs().max().item() / out_ref.abs().max().item()
    ref = {n: p.grad for n, p in oracle.named_parameters()}
    worst, num, den = (0.0, None), 0.0, 0.0
    rms_all = math.sqrt(sum(r.double().pow(2).sum().item() for r in ref.values()) / sum(r.numel() for r in ref.values()))
    for name, g in trainer.named_gradients():
        r = ref[name]
        # tensors whose true gradient is (numerically) zero — the key bias: softmax is invariant to it — are measured against the
        # gradient's overall scale instead of their own ~1e-9 norm
        e = ((g - r).norm() / max(r.norm().item(), 1e-2 * rms_all * math.sqrt(r.numel()))).item()
        num += (g - r).double().pow(2).sum().item()
        den += r.double().pow(2).sum().item()
        if e > worst[0]:
            worst = (e, name)
    total = math.sqrt(num / den)
    print(f"bf16 tc_mask={tc_mask} B={B} @{size}: {counts}; forward max err {fwd_err:.2e}; loss {loss.item():.6f} vs {loss_ref.item():.6f}; "
          f"whole-gradient rel L2 {total:.2e}; worst tensor {worst[1]} {worst[0]:.2e}")
    assert fwd_err <= 2e-2
    assert abs(loss.item() - loss_ref.item()) <= 1e-2 * abs(loss_ref.item())
    assert total <= 1e-2
    assert worst[0] <= 3e-2, worst


def test_trainer_step_with_cfg_coin_and_ema_copy():
    """`DenoiserTrainer.step` as the reference's loop body runs it for classifier-free-guidance training (proba_uncond > 0, bf16 mixed
    precision, EMA): the shared-seed coin takes both branches, the loss stays finite and decreases on a fixed batch, and
    `copy_ema_to_model` (EMAModel.copy_to, utils_training.py:674-676) swaps the module's parameters for the shadow ones."""
    B, size = 4, 64
    oracle, model, osched, sched, Trainer = _setup("small_denoiser_config", size, B, "3k_steps_clipping_rescaling")
    x, labels, noise, timesteps = _inputs(B, size, seed=17)
    trainer = Trainer(model, sched, B, size, learning_rate=5e-4, use_ema=True, proba_uncond=0.5, mixed_precision="bf16", shared_seed=3)
    coin = torch.Generator().manual_seed(3)
    expected = [bool(torch.rand(1, generator=coin).item() < 0.5) for _ in range(8)]
    assert any(expected) and not all(expected)
    losses = [trainer.step(x, labels, noise=noise, timesteps=timesteps).item() for _ in range(8)]
    assert all(math.isfinite(v) for v in losses) and min(losses[4:]) < losses[0], losses
    assert trainer.global_step == 8 and 0.0 < trainer.cur_decay_value < 1.0
    before = model.conv_in.weight.detach().clone()
    trainer.copy_ema_to_model()
    assert (model.conv_in.weight - before).abs().max().item() > 0
    assert torch.equal(trainer.params, trainer.ema)
    out = model(x, torch.full((B,), 500, device="cuda"), class_labels=labels).sample     # the inference route re-reads the swapped weights
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("mixed", ["no", "bf16"])
def test_gradients_sd21_width_config(mixed):
    """The fourth shipped denoiser JSON (SD-2.1 widths: 320 / 640 / 1280 / 1280, attention at three levels, 642 M parameters) through the
    training step at 32x32: channel counts that are multiples of 64 but not of 128 (the weight-gradient kernel falls back to CUDA cores
    for them), 2560-channel GroupNorms, S = 1024 / 256 / 64 attention."""
    B, size = 1, 32
    oracle, model, osched, sched, Trainer = _setup("SD_2-1_config", size, B, "1k_epsilon_pred")
    x, labels, noise, timesteps = _inputs(B, size, seed=4)
    loss_ref, _ = _oracle_loss(oracle, osched, x, labels, noise, timesteps, "epsilon", False)
    loss_ref.backward()
    trainer = Trainer(model, sched, B, size, mixed_precision=mixed)
    loss = trainer.diffusion_and_backward(x, labels, noise=noise, timesteps=timesteps)
    ref = {n: p.grad for n, p in oracle.named_parameters()}
    num = sum((g - ref[n]).double().pow(2).sum().item() for n, g in trainer.named_gradients())
    den = sum(r.double().pow(2).sum().item() for r in ref.values())
    rel = math.sqrt(num / den)
    print(f"[SD-2.1 widths, mixed_precision={mixed}] loss {loss.item():.6f} vs {loss_ref.item():.6f}; whole-gradient rel L2 {rel:.2e}; {trainer.tensor_core_counts()}")
    assert abs(loss.item() - loss_ref.item()) <= (1e-4 if mixed == "no" else 1e-2) * abs(loss_ref.item())
    assert rel <= (1e-4 if mixed == "no" else 1e-2)
