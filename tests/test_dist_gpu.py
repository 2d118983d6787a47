"""GPU, world_size 2: the data-parallel exchange of the latent path through clc_peer_allreduce (one-shot
all-reduce over peer memory opened with CUDA IPC).  The two ranks are two processes that share cuda:0 (a test
box has one GPU; IPC works between processes on one device just as between devices), `gloo` carries the
handles.  Checks: bpp statistic = sum over ranks, EntropyBottleneck parameter gradients = MEAN over ranks (the
gradient of the global-batch bpp, like nn.DataParallel's), identical bits on both ranks, repeated steps and
CUDA-graph replays stay in step."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from clc_b200.latent_path import LatentPath
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    out = {}
    for seed_rank in range(world):        # every rank also computes every shard alone (no collective)
        lp = LatentPath(1, 256, 256, n_refs=2, train=True, match_mode="tc", fused_slices=True, device="cuda:0")
        lp.randomize(seed=100 + seed_rank)
        lp.step()
        torch.cuda.synchronize()
        out[f"log2_{seed_rank}"] = lp.log2.cpu().clone()
        out[f"geb_{seed_rank}"] = lp._eb_grads_flat.cpu().clone()
    lp = LatentPath(1, 256, 256, n_refs=2, train=True, match_mode="tc", fused_slices=True, device="cuda:0",
                    data_parallel=True, collective="peer")
    assert lp._peer is not None
    lp.randomize(seed=100 + rank)
    for _ in range(3):                    # repeated eager steps (step counter / slot parity)
        lp.step()
    torch.cuda.synchronize()
    out["dp_log2"] = lp.log2.cpu().clone()
    out["dp_geb"] = lp._eb_grads_flat.cpu().clone()
    out["dp_bpp"] = lp.bpp().item()
    lp.capture()
    for _ in range(2):
        lp.replay()
    torch.cuda.synchronize()
    out["graph_log2"] = lp.log2.cpu().clone()
    out["graph_geb"] = lp._eb_grads_flat.cpu().clone()
    dist.barrier()
    lp._peer.close()
    q.put((rank, {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_allreduce_latent_path_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = {r: {k: (torch.from_numpy(v) if not isinstance(v, float) else v) for k, v in d.items()} for r, d in res.items()}
    a, b = res[0], res[1]
    want_log2 = a["log2_0"] + a["log2_1"]
    want_geb = (a["geb_0"].double() + a["geb_1"].double()) / 2
    for r in (a, b):
        for tag in ("dp", "graph"):
            assert torch.allclose(r[f"{tag}_log2"], want_log2, rtol=1e-12, atol=1e-6), tag
            assert torch.allclose(r[f"{tag}_geb"].double(), want_geb, rtol=1e-4, atol=1e-9), tag
    assert torch.equal(a["dp_geb"], b["dp_geb"]) and torch.equal(a["dp_log2"], b["dp_log2"])   # fixed summation order
    n_pix = 2 * 256 * 256
    assert abs(a["dp_bpp"] - (-(want_log2[0] + want_log2[1]).item() / n_pix)) < 1e-9
