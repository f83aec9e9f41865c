#!/usr/bin/env python
"""MEASURED CPU baseline of BASELINE.json configs[0]: the oracle (CPU restatement of the reference's
diffusers path) on the host cores of the box it runs on — small_denoiser and super_small at 64x64, batch 4, 10 + 10 DDIM steps,
whole run timed (no extrapolation).  Prints one JSON line per model with the CPU model and the threads used."""
import json
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import OracleCondUNet2D, OracleDDIMScheduler, OraclePipeline, oracle_ddib  # noqa: E402
from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS  # noqa: E402


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    x = (torch.randn(4, 3, 64, 64, generator=g) * 0.5).clamp(-1, 1)
    src = torch.arange(4) % 2
    for name in ("small_denoiser_config", "super_small"):
        torch.manual_seed(0)
        unet = OracleCondUNet2D(**dict(DENOISER_CONFIGS[name], sample_size=64)).eval()
        pipe = OraclePipeline(unet, OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]))
        oracle_ddib(pipe, x[:1], src[:1], 1 - src[:1], 1, return_raw=True)   # warm-up (thread pool, oneDNN primitives)
        t0 = time.perf_counter()
        oracle_ddib(pipe, x, src, 1 - src, 10, return_raw=True)
        dt = time.perf_counter() - t0
        print(json.dumps({"config": f"BASELINE configs[0]: {name} @64x64, batch 4, 10+10 DDIM steps, fp32 CPU oracle", "seconds": dt,
                          "images_per_s": 4 / dt, "cpu": cpu_model(), "logical_cpus": os.cpu_count(),
                          "torch_threads": torch.get_num_threads(), "mkldnn": torch.backends.mkldnn.is_available()}), flush=True)


if __name__ == "__main__":
    main()
