"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in vamb_b200/parallel.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vamb_b200 import parallel as par

    out = {}
    # gradient arena: mean of the two ranks' "gradients"
    g = torch.full((1000,), float(rank + 1))
    par.allreduce_mean_(g)
    out["grad"] = float(g[0]), float(g[-1])
    # every rank takes the same number of steps
    out["steps"] = par.agree_min(100 + 7 * rank)
    # BatchNorm running statistics
    bn = torch.nn.BatchNorm1d(4)
    bn.running_mean.fill_(rank)
    bn.running_var.fill_(2.0 + rank)
    par.average_running_stats_([bn])
    out["bn"] = float(bn.running_mean[0]), float(bn.running_var[0])
    # ragged row gather
    lo, hi = par.shard_rows(11, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1).repeat(1, 3)
    full = par.gather_rows(local)
    out["gather"] = full[:, 0].tolist()
    # data-parallel SGD on a toy quadratic keeps the replicas identical and equals full-batch SGD
    torch.manual_seed(0)
    w = torch.zeros(3)
    x = torch.arange(24, dtype=torch.float32).reshape(8, 3) / 10
    y = x @ torch.tensor([1.0, -2.0, 0.5])
    lo, hi = par.shard_rows(8, rank, world)
    for _ in range(50):
        grad = 2 * x[lo:hi].T @ (x[lo:hi] @ w - y[lo:hi]) / (hi - lo)
        par.allreduce_mean_(grad)
        w -= 0.05 * grad
    out["w"] = w.tolist()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["grad"] == (1.5, 1.5)
        assert res[r]["steps"] == 100
        assert res[r]["bn"] == (0.5, 2.5)
        assert res[r]["gather"] == [float(i) for i in range(11)]
    assert res[0]["w"] == res[1]["w"]
    # equals single-process full-batch gradient descent (shards have equal size)
    w = torch.zeros(3)
    x = torch.arange(24, dtype=torch.float32).reshape(8, 3) / 10
    y = x @ torch.tensor([1.0, -2.0, 0.5])
    for _ in range(50):
        w -= 0.05 * (2 * x.T @ (x @ w - y) / 8)
    assert np.allclose(res[0]["w"], w.tolist(), atol=1e-5)


def test_shard_rows_partition():
    from vamb_b200 import parallel as par

    for n in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [par.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        par.shard_rows(10, 2, 2)
