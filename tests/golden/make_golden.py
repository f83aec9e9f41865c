"""Generates the committed golden fixtures from the ORACLE (the reference has no vectors of its own and diffusers 0.18.2 is
not importable here; the oracle's blocks and schedulers are pinned by tests/test_oracle_published_kats.py, see oracle/ headers).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import OracleCondUNet2D, OracleDDIMInverseScheduler, OracleDDIMScheduler, OraclePipeline, oracle_ddib  # noqa: E402
from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(0)
    cfg = dict(DENOISER_CONFIGS["super_small"], sample_size=32)
    unet = OracleCondUNet2D(**cfg).eval()
    g = torch.Generator().manual_seed(1234)
    x = (torch.randn(2, 3, 32, 32, generator=g) * 0.5).clamp(-1, 1)
    labels = torch.tensor([0, 1])
    t = torch.tensor(1499)
    with torch.no_grad():
        eps = unet(x, t, labels).sample
        emb = unet.embed(x, t, labels)
    torch.save({"x": x, "t": t, "labels": labels, "eps": eps, "emb": emb}, os.path.join(HERE, "unet_super_small_32.pt"))

    pipe = OraclePipeline(unet, OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]))
    out = oracle_ddib(pipe, x, labels, 1 - labels, 3, return_raw=True)
    torch.save({"x": x, "src": labels, "tgt": 1 - labels, "n": 3, "out": out}, os.path.join(HERE, "ddib_super_small_32_n3.pt"))

    # scheduler known-answer tables
    kat = {}
    for name, c in SCHEDULER_CONFIGS.items():
        s = OracleDDIMScheduler.from_config(c)
        inv = OracleDDIMInverseScheduler.from_config(s.config)
        s.set_timesteps(100)
        inv.set_timesteps(100)
        N = s.config.num_train_timesteps
        idx = [0, N // 2, N - 2, N - 1]
        kat[name] = {"alphas_cumprod": s.alphas_cumprod[idx].clone(), "inv_alphas_cumprod": inv.alphas_cumprod[idx].clone(),
                     "timesteps": s.timesteps.clone(), "inv_timesteps": inv.timesteps.clone()}
    torch.save(kat, os.path.join(HERE, "scheduler_kat.pt"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
