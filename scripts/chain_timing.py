#!/usr/bin/env python
"""Where does the step time go?  Times CUDA-graph replays (L2 flushed, CUDA events) of
the whole step, of each chain alone, and of prefixes of the match chain (stage mask)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from clc_b200.latent_path import LatentPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--no-fuse", action="store_true", help="separate CLM kernels instead of the fused chain")
ap.add_argument("--fuse-all", action="store_true", help="forward CLM fusion at any latent size (default: <= 512 pixels)")
a = ap.parse_args()
cfg = WORKLOADS[a.workload]
lp = LatentPath(cfg["B"], cfg["H"], cfg["W"], n_refs=cfg["R"], train=cfg["train"], fused_slices=True, device="cuda:0",
                fuse_chain=("all" if a.fuse_all else not a.no_fuse))
lp.randomize(seed=1)
lp.step()
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda:0")


def timeit(g):
    for _ in range(5):
        flush.zero_()
        g.replay()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def empty():
    pass


variants = {
    "whole step (forked)": lambda: lp.step(fork=True),
    "whole step (serial)": lambda: lp.step(fork=False),
    "match chain": lp.match_chain,
    "hyper chain": lp.hyper_chain,
    "slice chain": lp.slice_chain,
    "entropy (hyper || slices)": lp._entropy_forked,
}
for name, fn in variants.items():
    res = [timeit(lp._capture(fn)) for _ in range(3)]      # re-captured: capture-to-capture variance
    print(f"{name:32s} " + " ".join(f"{t:8.1f}" for t in res) + " us")
