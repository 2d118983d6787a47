"""CPU, world_size 2, gloo: the data-parallel host logic (sharding, bucketed gradient
all-reduce with unused parameters, statistics all-reduce)."""
import os
import socket

import torch
import torch.multiprocessing as mp
import torch.nn as nn

from clc_b200 import dist as cd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.unused = nn.Linear(3, 3)  # like the reference's never-called feature_alignment
        self.b = nn.Linear(16, 2)

    def forward(self, x):
        return self.b(torch.tanh(self.a(x)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = cd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = _Net()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    t = torch.randn(8, 2, generator=g)
    lo, hi = cd.shard_range(8, rank, world)
    red = cd.GradAllReducer(net, bucket_mb=0.0001)  # tiny buckets -> several async all-reduces
    # step 0 discovers which parameters receive gradients and in which order; steps 1.. use one hook per bucket
    for it in range(3):
        red.zero_grad()
        loss = ((net(x[lo:hi]) - t[lo:hi]) ** 2).mean()
        loss.backward()
        red.finish()
    # after finish() every gradient is a view into its (reduced) bucket: no scatter copy
    flats = {f.data_ptr(): f for f, _, _ in red._buckets}
    assert len(red._buckets) > 1
    assert any(f.data_ptr() <= net.a.weight.grad.data_ptr() < f.data_ptr() + 4 * f.numel() for f in flats.values())
    assert all(p.grad is None for p in net.unused.parameters())
    bpp, mse = cd.allreduce_stats(torch.tensor(-10.0 * (rank + 1)), -1.0, 3.0 * (rank + 1), 4.0)
    q.put((rank, {n: p.grad.numpy().copy() for n, p in net.named_parameters() if p.grad is not None}, bpp, mse))
    torch.distributed.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 24, 64):
        for w in (1, 2, 3, 8):
            spans = [cd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_grad_allreduce_and_stats_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process full-batch gradients: equal shards -> mean of shard grads == full-batch grad
    torch.manual_seed(0)
    net = _Net()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    t = torch.randn(8, 2, generator=g)
    ((net(x) - t) ** 2).mean().backward()
    for rank, grads, bpp, mse in res:
        assert "unused.weight" not in grads
        for n, p in net.named_parameters():
            if p.grad is not None:
                assert torch.allclose(torch.from_numpy(grads[n]), p.grad, atol=1e-6), n
        assert abs(bpp - (-(-10.0 - 20.0 - 2.0) / 8.0)) < 1e-12
        assert abs(mse - 9.0 / 24.0) < 1e-12
