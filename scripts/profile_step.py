#!/usr/bin/env python
"""Run a few un-graphed LatentPath steps of one workload -- the command the ncu captures under
profiles/ are taken from:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python scripts/profile_step.py --workload cfg2 --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from clc_b200.latent_path import LatentPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--match-mode", default="tc")
ap.add_argument("--per-slice", action="store_true")
a = ap.parse_args()
cfg = WORKLOADS[a.workload]
lp = LatentPath(cfg["B"], cfg["H"], cfg["W"], n_refs=cfg["R"], train=cfg["train"], match_mode=a.match_mode,
                fused_slices=not a.per_slice, device="cuda:0")
lp.randomize(seed=1)
for _ in range(a.steps):
    lp.step()
torch.cuda.synchronize()
print("bpp", lp.bpp().item())
