"""World-size-2 `gloo` tests (CPU) of the N>1 path's host logic: env-id sharding, the advantage-statistics
all-reduce and the flat gradient all-reduce (opendog_b200/train.py). The GPU job runs the same functions over NCCL."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opendog_b200.train import (allreduce_advantage_stats, allreduce_flat_grads, shard_range, stats_mean_std)
    T, N = 6, 10
    g = torch.Generator().manual_seed(0)
    adv_all = torch.randn(T, world * N, generator=g, dtype=torch.float64)
    lo, hi = shard_range(rank, world, N)
    mine = adv_all[:, lo:hi]
    stats = torch.tensor([mine.sum(), (mine ** 2).sum(), mine.numel()], dtype=torch.float64)
    allreduce_advantage_stats(stats)
    mean, std = stats_mean_std(stats)
    # gradient averaging
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    x_all = torch.randn(world * 8, 5, generator=g)
    loss = net(x_all[rank * 8:(rank + 1) * 8]).pow(2).mean()
    loss.backward()
    flat = allreduce_flat_grads(list(net.parameters()))
    q.put((rank, (lo, hi), float(mean), float(std), flat.numpy().copy(),
           [p.grad.numpy().copy() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_and_gradients_match_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 10) and res[1][1] == (10, 20)
    # single-process reference
    g = torch.Generator().manual_seed(0)
    adv_all = torch.randn(6, 20, generator=g, dtype=torch.float64)
    for r in res:
        assert abs(r[2] - float(adv_all.mean())) < 1e-12
        assert abs(r[3] - float(adv_all.std())) < 1e-12          # torch.std: unbiased, as sim2real/train.py:564
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 2))
    x_all = torch.randn(16, 5, generator=g)
    net(x_all).pow(2).mean().backward()                           # mean over both shards == average of shard means
    ref = [p.grad.numpy() for p in net.parameters()]
    for r in res:
        for a, b in zip(r[5], ref):
            assert np.abs(a - b).max() < 1e-6
    assert np.array_equal(res[0][4], res[1][4])                   # both ranks hold the same flat buffer
