"""CPU tests of the oracle itself: known-answer values, committed golden fixtures, parameter
counts and the structural properties of the restated schedulers.  (The reference has no tests or vectors of
its own; the third-party package's PUBLISHED known answers are in tests/test_oracle_published_kats.py, these tests add the
derived KATs and guard the oracle against drift from its committed outputs.)"""
import os

import pytest
import torch

from oracle import (OracleCondUNet2D, OracleDDIMInverseScheduler, OracleDDIMScheduler, OraclePipeline, oracle_ddib,
                    oracle_inversion)
from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# (N, abar[0], abar[N/2], abar[N-2], abar[N-1]) rescaled table, then the un-rescaled (inverse 0.18.2) one
KAT = {
    "1k_epsilon_pred": (1000, (0.99990010, 0.31781536, 7.986e-08, 0.0), (0.99989998, 0.33127463, 7.4838e-04, 7.3341e-04)),
    "3k_steps_clipping_rescaling": (3000, (0.99989998, 0.036693018, 4.09e-14, 0.0), (0.99989998, 0.036699165, 4.049e-10, 3.968e-10)),
    "better_SD_config": (3000, (0.99999011, 0.12434004, 1.10e-11, 0.0), None),
    "SD_orig_config": (1000, (0.99914998, 0.27633247, 4.7167e-03, 4.6601e-03), (0.99914998, 0.27633247, 4.7167e-03, 4.6601e-03)),
}


@pytest.mark.parametrize("name", list(KAT))
def test_scheduler_tables_known_answers(name):
    N, gen, inv = KAT[name]
    s = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS[name])
    assert s.config.num_train_timesteps == N
    idx = [0, N // 2, N - 2, N - 1]
    for got, want in zip(s.alphas_cumprod[idx].tolist(), gen):
        assert got == pytest.approx(want, rel=2e-3, abs=1e-16)
    if inv is not None:
        i = OracleDDIMInverseScheduler.from_config(s.config)
        for got, want in zip(i.alphas_cumprod[idx].tolist(), inv):
            assert got == pytest.approx(want, rel=2e-3, abs=1e-16)


def test_timestep_grids_known_answers():
    s = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["1k_epsilon_pred"])
    s.set_timesteps(100)
    assert s.timesteps[:3].tolist() == [999, 989, 979] and s.timesteps[-2:].tolist() == [19, 9]
    s.set_timesteps(10)
    assert s.timesteps.tolist() == [999, 899, 799, 699, 599, 499, 399, 299, 199, 99]
    i = OracleDDIMInverseScheduler.from_config(s.config)
    i.set_timesteps(10)
    assert i.timesteps.tolist() == list(range(0, 1000, 100))
    s3 = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"])
    s3.set_timesteps(100)
    assert s3.timesteps[:2].tolist() == [2999, 2969] and s3.timesteps[-1].item() == 29
    i3 = OracleDDIMInverseScheduler.from_config(s3.config)
    i3.set_timesteps(100)
    assert i3.timesteps[:3].tolist() == [0, 30, 60] and i3.timesteps[-1].item() == 2970
    so = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["SD_orig_config"])
    so.set_timesteps(100)
    assert so.timesteps[:2].tolist() == [991, 981] and so.timesteps[-1].item() == 1  # leading, steps_offset 1
    with pytest.raises(ValueError):
        s.set_timesteps(1001)


def test_scheduler_golden_fixture():
    kat = torch.load(os.path.join(GOLDEN, "scheduler_kat.pt"))
    for name, c in SCHEDULER_CONFIGS.items():
        s = OracleDDIMScheduler.from_config(c)
        inv = OracleDDIMInverseScheduler.from_config(s.config)
        s.set_timesteps(100)
        inv.set_timesteps(100)
        N = s.config.num_train_timesteps
        idx = [0, N // 2, N - 2, N - 1]
        assert torch.equal(s.alphas_cumprod[idx], kat[name]["alphas_cumprod"])
        assert torch.equal(inv.alphas_cumprod[idx], kat[name]["inv_alphas_cumprod"])
        assert torch.equal(s.timesteps, kat[name]["timesteps"]) and torch.equal(inv.timesteps, kat[name]["inv_timesteps"])


def test_inverse_scheduler_drops_unknown_keys_and_last_step_is_eps():
    s = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"])
    inv = OracleDDIMInverseScheduler.from_config(s.config)
    assert "rescale_betas_zero_snr" not in inv.config and "timestep_spacing" not in inv.config  # A.4 (0.18.2)
    assert float(inv.alphas_cumprod[-1]) > 0  # un-rescaled table
    inv.set_timesteps(10)
    x, m = torch.randn(1, 3, 4, 4), torch.randn(1, 3, 4, 4)
    t = inv.timesteps[-1]
    out = inv.step(m, t, x).prev_sample
    a = inv.alphas_cumprod[int(t)]
    assert torch.allclose(out, a.sqrt() * m + (1 - a).sqrt() * x)  # a' = 0 => x_T = eps_hat (v-prediction)
    new = OracleDDIMInverseScheduler.from_config(s.config, variant=">=0.19")
    assert float(new.alphas_cumprod[-1]) == 0.0
    new.set_timesteps(10)
    assert new.timesteps.tolist() == [299, 599, 899, 1199, 1499, 1799, 2099, 2399, 2699, 2999]


def test_zero_snr_epsilon_first_step_is_clamped_inf():
    s = OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["1k_epsilon_pred"])
    s.set_timesteps(10)
    x = torch.tensor([[1.0, -2.0, 0.5]])
    m = torch.tensor([[0.5, 1.0, 0.5]])
    r = s.step(m, s.timesteps[0], x)
    assert r.pred_original_sample[0, 0] == 1.0 and r.pred_original_sample[0, 1] == -1.0
    assert torch.isnan(r.pred_original_sample[0, 2])  # (x - m)/0 with x == m


def test_step_prediction_types_consistent():
    """x0/eps of the three prediction types describe the same point when fed consistent model outputs."""
    base = dict(SCHEDULER_CONFIGS["SD_orig_config"], clip_sample=False)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(2, 3, 4, 4, generator=g), torch.randn(2, 3, 4, 4, generator=g)
    outs = []
    for pt in ("epsilon", "sample", "v_prediction"):
        s = OracleDDIMScheduler.from_config(dict(base, prediction_type=pt))
        s.set_timesteps(20)
        t = s.timesteps[5]
        xt = s.add_noise(x0, eps, t.repeat(2))
        m = {"epsilon": eps, "sample": x0, "v_prediction": s.get_velocity(x0, eps, t.repeat(2))}[pt]
        outs.append(s.step(m, t, xt).prev_sample)
    assert torch.allclose(outs[0], outs[1], atol=2e-5) and torch.allclose(outs[0], outs[2], atol=2e-5)


@pytest.mark.parametrize("name,count", [("small_denoiser_config", 62826243), ("super_small", 15725443)])
def test_unet_parameter_counts(name, count):
    m = OracleCondUNet2D(**DENOISER_CONFIGS[name])
    assert sum(p.numel() for p in m.parameters()) == count
    keys = set(m.state_dict())
    for k in ("conv_in.weight", "time_embedding.linear_1.weight", "class_embedding.weight",
              "down_blocks.0.resnets.0.time_emb_proj.weight", "down_blocks.2.attentions.1.to_out.0.bias",
              "down_blocks.1.downsamplers.0.conv.weight", "mid_block.attentions.0.group_norm.weight",
              "up_blocks.0.resnets.2.conv_shortcut.weight", "up_blocks.1.upsamplers.0.conv.bias", "conv_norm_out.weight"):
        assert k in keys, k  # Appendix A.7 layout


def test_unet_golden_and_errors():
    g = torch.load(os.path.join(GOLDEN, "unet_super_small_32.pt"))
    torch.manual_seed(0)
    m = OracleCondUNet2D(**dict(DENOISER_CONFIGS["super_small"], sample_size=32)).eval()
    with torch.no_grad():
        eps = m(g["x"], g["t"], g["labels"]).sample
    assert (eps - g["eps"]).abs().max().item() <= 1e-5
    with pytest.raises(ValueError):
        m(g["x"], g["t"], g["labels"], torch.zeros(2, 256))
    with pytest.raises(ValueError):
        m(g["x"], g["t"])
    # class_emb == table rows reproduces the label path
    with torch.no_grad():
        e2 = m(g["x"], g["t"], class_emb=m.class_embedding(g["labels"])).sample
    assert torch.allclose(eps, e2, atol=1e-6)


def test_ddib_golden_and_same_class_roundtrip():
    g = torch.load(os.path.join(GOLDEN, "ddib_super_small_32_n3.pt"))
    torch.manual_seed(0)
    m = OracleCondUNet2D(**dict(DENOISER_CONFIGS["super_small"], sample_size=32)).eval()
    pipe = OraclePipeline(m, OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]))
    out = oracle_ddib(pipe, g["x"], g["src"], g["tgt"], g["n"], return_raw=True)
    assert (out - g["out"]).abs().max().item() <= 1e-4
    # per-sample independence: the batch is embarrassingly parallel (the property multi-GPU sharding rests on, §8e)
    one = oracle_ddib(pipe, g["x"][1:], g["src"][1:], g["tgt"][1:], g["n"], return_raw=True)
    assert (one - out[1:]).abs().max().item() <= 1e-4
    lat = oracle_inversion(pipe, g["x"], g["src"], 3)
    assert lat.shape == g["x"].shape and torch.isfinite(lat).all()
    imgs = oracle_ddib(pipe, g["x"], g["src"], g["tgt"], 2)
    assert imgs.shape == (2, 32, 32, 3) and imgs.min() >= 0 and imgs.max() <= 1
