#!/usr/bin/env python
"""Context measurement (not the latent-path metric): the drop-in CLC.forward in eval mode, eager vs captured into
one CUDA graph (models.make_graphed_forward), with and without the forked mean / scale branches.
  python scripts/model_forward_timing.py [--N 128] [--size 256 256] [--batch 1] [--refs 3]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--size", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--refs", type=int, default=3)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--channels-last", action="store_true", help="weights and images in torch.channels_last")
    ap.add_argument("--tf32-matmul", action="store_true", help="TF32 tensor cores for the Linear layers (graphs only)")
    a = ap.parse_args()
    from clc_b200.models import CLC
    d = torch.device("cuda:0")
    torch.manual_seed(0)
    m = CLC(N=a.N, num_ref_frames=a.refs).eval().to(d)
    H, W = a.size
    x = torch.rand(a.batch, 3, H, W, device=d)
    refs = [torch.rand(a.batch, 3, H, W, device=d) for _ in range(a.refs)]
    if a.channels_last:
        m = m.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)
        refs = [r.contiguous(memory_format=torch.channels_last) for r in refs]

    def eager():
        with torch.no_grad():
            return m(x, refs)

    res = {"model": f"CLC(N={a.N})", "channels_last": a.channels_last, "tf32_matmul": a.tf32_matmul, "batch": a.batch, "image": [H, W], "n_refs": a.refs,
           "eager_ms": timed(eager, a.iters)}
    for fork in (False, True):
        run = m.make_graphed_forward(x, refs, fork_branches=fork, channels_last=a.channels_last,
                                     tf32_matmul=a.tf32_matmul)
        res["graph_forked_ms" if fork else "graph_ms"] = timed(lambda: run(x, refs), a.iters)
    res["kernel_nodes_note"] = "same kernels in all three; the difference is host launch overhead and branch overlap"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
