"""GPU parity on the configuration the headline number is quoted on (BASELINE.json configs[1]; the reference pairs
small_denoiser_config @ 128x128 with 100 inference steps, SLURM_launch_script_DDIM_perc_a100.sh:105-112, and runs it
through utils_Img2Img.py:566-612): the micro-batch planner's 64-image passes, the 16x16-pixel dual-accumulator tiles at
128x128, four micro-batches per call and 100 + 100 free-running steps.

The oracle runs ON THE GPU here, in fp32 with TF32 off (cuDNN / cuBLAS fp32 kernels), in chunks — the CPU oracle would
take minutes at this size.  It is the checker only; the product path is the C-ABI library.

Bars (BASELINE.json north_star): per-step eps max-abs error <= 1e-2 (16-bit storage), <= 1e-4 (fp32 validation mode),
final translated images >= 40 dB PSNR.  fp16 (the reference's own autocast type, the product default) must meet them;
bf16 is measured through the same tests, printed, and held to the documented deviation.
"""
import pytest
import torch

from tests.util import make_pair, psnr, synth_images

pytestmark = pytest.mark.gpu

SCHED = "3k_steps_clipping_rescaling"
# max-abs / rms bars per storage type; bf16's max-abs is the recorded deviation from the 1e-2 the spec states
MAXABS = {"fp16": 1e-2, "bf16": 4e-2, "fp32": 1e-4}
RMS = {"fp16": 1e-3, "bf16": 5e-3, "fp32": 2e-5}


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _oracle_forward(oracle, x, t, labels, chunk=16):
    """fp32 oracle on the device, `chunk` images at a time (the explicit softmax holds B x heads x S x S fp32)."""
    outs = []
    with torch.no_grad():
        for i in range(0, x.shape[0], chunk):
            outs.append(oracle(x[i:i + chunk], torch.tensor(t), labels[i:i + chunk]).sample)
    return torch.cat(outs)


@pytest.mark.parametrize("precision,batch", [("fp16", 64), ("fp16", 256), ("bf16", 64), ("fp32", 2)])
def test_headline_forward_128(build_lib, precision, batch):
    """small_denoiser @ 128x128: one micro-batch of 64, four of them (batch 256), t at both ends and the middle."""
    oracle, model = make_pair("small_denoiser_config", 128, precision)
    oracle = oracle.cuda()
    x, labels = synth_images(batch, 128)
    x, labels = x.cuda(), labels.cuda()
    worst = 0.0
    for t in (0, 1500, 2999):
        ref = _oracle_forward(oracle, x, t, labels)
        got = model(x, torch.tensor(t), labels).sample
        err = (got - ref).abs().max().item()
        rms = (got - ref).pow(2).mean().sqrt().item()
        worst = max(worst, err)
        print(f"[headline fwd {precision} B={batch}] t={t}: max abs err {err:.3e}, rms {rms:.3e} (ref max {ref.abs().max():.3f})")
        assert err <= MAXABS[precision], f"t={t}: {precision} eps max-abs err {err:.3e}"
        assert rms <= RMS[precision], f"t={t}: {precision} eps rms err {rms:.3e}"
    if precision != "fp32":
        assert model.plan_info()["microbatch"] == 64, model.plan_info()


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_headline_ddib_n100_128(build_lib, precision):
    """100 + 100 steps at 128x128, batch 8: free-running PSNR >= 40 dB against the oracle's images AND teacher-forced
    per-step eps (the product UNet fed the oracle's own x_t) at every 10th step of both directions."""
    from oracle import OracleDDIMScheduler, OraclePipeline, oracle_ddib
    from phendiff_b200 import ConditionalDDIMPipeline, DDIMScheduler, ddib_transfer
    from phendiff_b200.reference_configs import SCHEDULER_CONFIGS

    n = 100
    oracle, model = make_pair("small_denoiser_config", 128, precision)
    oracle = oracle.cuda()
    x, src = synth_images(8, 128)
    x, src = x.cuda(), src.cuda()
    tgt = 1 - src
    o_pipe = OraclePipeline(oracle, OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS[SCHED]))
    pipe = ConditionalDDIMPipeline(model, DDIMScheduler.from_config(SCHEDULER_CONFIGS[SCHED]))
    tr_inv, tr_gen = [], []
    ref = oracle_ddib(o_pipe, x, src, tgt, n, return_raw=True, trace_inv=tr_inv, trace_gen=tr_gen)
    assert len(tr_inv) == n and len(tr_gen) == n
    worst = {"inv": 0.0, "gen": 0.0}
    for name, trace, labels in (("inv", tr_inv, src), ("gen", tr_gen, tgt)):
        for k in list(range(0, n, 10)) + [n - 1]:
            t, xt, eps = trace[k]
            ok = ~(torch.isnan(xt).flatten(1).any(1) | torch.isnan(eps).flatten(1).any(1))
            if not bool(ok.any()):
                continue
            got = model(torch.nan_to_num(xt), t, labels).sample
            err = (got - eps)[ok].abs().max().item()
            worst[name] = max(worst[name], err)
    print(f"[headline ddib {precision}] teacher-forced eps max-abs err: inversion {worst['inv']:.3e}, generation {worst['gen']:.3e}")
    assert max(worst.values()) <= MAXABS[precision]
    got = ddib_transfer(pipe, x, src, tgt, n)
    ok = ~(torch.isnan(ref) | torch.isnan(got))
    assert bool(ok.any())
    p = psnr(ref[ok].cpu(), got[ok].cpu())
    print(f"[headline ddib {precision}] n={n} free-running PSNR {p:.1f} dB, max abs {float((ref - got)[ok].abs().max()):.3e}, "
          f"NaN pixels ref/got {int(torch.isnan(ref).sum())}/{int(torch.isnan(got).sum())}")
    assert p >= 40.0
