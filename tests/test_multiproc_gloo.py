"""World-size-2 gloo test of the multi-GPU host logic: batch-shard, run the class transfer per rank with no
data-path collective, all-gather the outputs once, and compare with the unsharded result.  On the CPU box the per-rank
transfer is the oracle (the CUDA path needs a GPU); what is under test is the sharding + gather plumbing bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import OracleCondUNet2D, OracleDDIMScheduler, OraclePipeline, oracle_ddib
        from phendiff_b200.reference_configs import DENOISER_CONFIGS, SCHEDULER_CONFIGS
        from phendiff_b200.sharding import gather_outputs, shard_range

        torch.manual_seed(0)
        unet = OracleCondUNet2D(**dict(DENOISER_CONFIGS["super_small"], sample_size=16)).eval()
        pipe = OraclePipeline(unet, OracleDDIMScheduler.from_config(SCHEDULER_CONFIGS["3k_steps_clipping_rescaling"]))
        g = torch.Generator().manual_seed(1234)
        x = (torch.randn(total, 3, 16, 16, generator=g) * 0.5).clamp(-1, 1)
        src = torch.arange(total) % 2
        lo, hi = shard_range(total, rank, world)
        local = oracle_ddib(pipe, x[lo:hi], src[lo:hi], 1 - src[lo:hi], 2, return_raw=True)
        full = gather_outputs(local, total)
        if rank == 0:
            ref = oracle_ddib(pipe, x, src, 1 - src, 2, return_raw=True)
            q.put((tuple(full.shape), float((full - ref).abs().max())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_sharded_transfer_equals_unsharded_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    shape, err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert shape == (total, 3, 16, 16)
    assert err <= 1e-4, f"sharded vs unsharded differ by {err:.3e}"


def _train_worker(rank, world, port, q):
    """Data-parallel training step: each rank computes the gradient of the oracle's loss on ITS half of the batch,
    `average_gradients` all-reduces the flat vector once; the result must equal the gradient of the loss on the whole batch
    (mse is a mean over the batch, so the average of the two half-batch gradients is the full-batch gradient)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn.functional as F

        from oracle import OracleCondUNet2D
        from phendiff_b200.reference_configs import DENOISER_CONFIGS
        from phendiff_b200.sharding import average_gradients, shard_range

        torch.manual_seed(0)
        unet = OracleCondUNet2D(**dict(DENOISER_CONFIGS["super_small"], sample_size=16))
        g = torch.Generator().manual_seed(7)
        total = 4
        x = torch.randn(total, 3, 16, 16, generator=g)
        noise = torch.randn(total, 3, 16, 16, generator=g)
        t = torch.randint(0, 1000, (total,), generator=g)
        labels = torch.arange(total) % 2

        def flat_grad(lo, hi):
            unet.zero_grad()
            F.mse_loss(unet(x[lo:hi], t[lo:hi], class_labels=labels[lo:hi]).sample, noise[lo:hi]).backward()
            return torch.cat([p.grad.flatten() for p in unet.parameters()])

        lo, hi = shard_range(total, rank, world)
        flat = average_gradients(flat_grad(lo, hi))
        if rank == 0:
            ref = flat_grad(0, total)
            q.put(float((flat - ref).abs().max() / ref.abs().max()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gradient_all_reduce_equals_full_batch_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err <= 1e-5, f"averaged shard gradients differ from the full-batch gradient by {err:.3e}"
