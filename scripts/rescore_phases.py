#!/usr/bin/env python
"""Bring-up: wall-clock phase breakdown of the re-scoring kernel (first 64 CTAs) from the %globaltimer stamps of
the debug build.  python scripts/rescore_phases.py [--workload cfg2]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from clc_b200 import _lib  # noqa: E402
import clc_b200.latent_path as LPm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
a = ap.parse_args()
cfg = dict(WORKLOADS[a.workload])
lp = LPm.LatentPath(cfg["B"], cfg["H"], cfg["W"], n_refs=cfg["R"], train=cfg["train"], match_mode="tc",
                    fused_slices=True, device="cuda:0")
lp.randomize(seed=1)
H = _lib.debug_lib()
real = LPm.call


def dbg_call(name, *args):
    rc = getattr(H, name)(*args)
    assert rc == 0, (name, rc)
    return rc


LPm.call = dbg_call
for _ in range(3):
    lp.match_chain()
torch.cuda.synchronize()
LPm.call = real
out = np.zeros((64, 16), dtype=np.int64)
assert H.clc_debug_rescore_stamps(out.ctypes.data_as(C.c_void_p)) == 0
names = ["launch->pdl_wait", "stage Q + lists", "re-score round 1", "select + certify", "(round 2)", "softmax", "blend",
         "write aligned + coef", "cluster barrier 1", "CLM from DSMEM", "cluster barrier 2"]
t0 = out[:, 0].min()
print("CTA start spread (us):", (out[:, 0].max() - t0) / 1e3)
d = np.diff(out[:, :12], axis=1) / 1e3
for i, n in enumerate(names):
    col = d[:, i]
    print(f"{n:24s} median {np.median(col):7.2f}  min {col.min():7.2f}  max {col.max():7.2f} us")
print(f"{'total':24s} median {np.median((out[:, 11] - out[:, 0]) / 1e3):7.2f} us; last end - first start {(out[:, :12].max() - t0) / 1e3:.2f} us")
if lp.train:
    out = np.zeros((64, 16), dtype=np.int64)
    assert H.clc_debug_bwd_stamps(out.ctypes.data_as(C.c_void_p)) == 0
    names = ["launch->pdl_wait", "idx/q/g/aligned/att loads", "G_r reduce+publish (w0-1)", "window pass 1 (loads+FMA)",
             "block reduction", "coefficients", "window pass 2 + red.add", "g_q red.add"]
    t0 = out[:, 0].min()
    print("\nmatch backward -- CTA start spread (us):", (out[:, 0].max() - t0) / 1e3)
    d = np.diff(out[:, :9], axis=1) / 1e3
    for i, n in enumerate(names):
        col = d[:, i]
        print(f"{n:28s} median {np.median(col):7.2f}  min {col.min():7.2f}  max {col.max():7.2f} us")
    print(f"{'total':28s} median {np.median((out[:, 8] - out[:, 0]) / 1e3):7.2f} us; last end - first start {(out[:, :9].max() - t0) / 1e3:.2f} us")
