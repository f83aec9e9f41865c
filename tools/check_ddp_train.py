#!/usr/bin/env python
"""Data-parallel training step on real NCCL (run under torchrun on 2 GPUs: `bash tools/gpu.sh <tag> ddpcheck` inside `gpurun --gpus 2`).
Each rank runs forward + backward on ITS shard, the flat gradient vector is averaged with one all-reduce; rank 0 also runs both
shards locally (gradient accumulation) and checks that the all-reduced vector equals their mean, and that after the optimizer
step every rank holds bit-identical parameters."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge

    ge.build()
    from phendiff_b200 import CustomCondUNet2DModel, DDIMScheduler
    from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS
    from phendiff_b200.training import DenoiserTrainer

    B, size = 4, 64
    torch.manual_seed(0)
    unet = CustomCondUNet2DModel.from_config(dict(DENOISER_CONFIGS["small_denoiser_config"], sample_size=size)).to(dev)
    sched = DDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"])
    tr = DenoiserTrainer(unet, sched, B, size, learning_rate=1e-3, use_ema=True, mixed_precision="bf16")
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(world * B, 3, size, size, generator=g) * 0.5).clamp(-1, 1).to(dev)
    noise = torch.randn(world * B, 3, size, size, generator=g).to(dev)
    ts = torch.randint(0, 3000, (world * B,), generator=g).to(dev)
    labels = (torch.arange(world * B) % 2).to(dev)
    sl = slice(rank * B, (rank + 1) * B)
    loss = tr.diffusion_and_backward(x[sl], labels[sl], noise=noise[sl], timesteps=ts[sl])
    tr.all_reduce_gradients()
    reduced = tr.grads.clone()
    ok = True
    if rank == 0:
        tr.zero_grad()
        for r in range(world):
            s2 = slice(r * B, (r + 1) * B)
            tr.diffusion_and_backward(x[s2], labels[s2], noise=noise[s2], timesteps=ts[s2])
        ref = tr.grads / world
        err = ((reduced - ref).norm() / ref.norm()).item()
        print(f"[ddpcheck] world {world}: all-reduced gradient vs mean of the shard gradients computed on rank 0: rel L2 {err:.3e} (loss rank0 {loss.item():.5f})")
        # bf16 mode is not bit-reproducible run to run (fp32 atomics reorder, and a last-bit change of a gradient can flip a bf16 rounding):
        # the floor measured here is ~2e-4; the all-reduce itself is exact
        ok = err <= 1e-3
    tr.grads.copy_(reduced)
    tr.optimizer_step()
    fp = torch.stack([tr.params.double().sum(), tr.params.double().abs().sum(), tr.ema.double().sum()])
    gathered = [torch.empty_like(fp) for _ in range(world)]
    dist.all_gather(gathered, fp)
    if rank == 0:
        same = all(torch.equal(gathered[0], t) for t in gathered)
        print(f"[ddpcheck] parameters / EMA after the optimizer step identical on all ranks: {same}")
        ok = ok and same
        print("[ddpcheck] OK" if ok else "[ddpcheck] FAILED")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
