#!/usr/bin/env python
"""Per-call / per-kernel device times of one LatentPath workload with everything L2-warm:
  (a) each C-ABI call repeated REP times inside its own CUDA graph (time per call incl. the
      in-graph launch gap), and
  (b) the library's per-kernel trace (one CUDA event after every kernel) over un-graphed steps.
Used to find which kernel of the serial match chain to shorten next."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from clc_b200 import _lib  # noqa: E402
from clc_b200.latent_path import LatentPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--rep", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--match-mode", default="tc")
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()
cfg = dict(WORKLOADS[a.workload])
if a.batch:
    cfg["B"] = a.batch
lp = LatentPath(cfg["B"], cfg["H"], cfg["W"], n_refs=cfg["R"], train=cfg["train"], match_mode=a.match_mode,
                fused_slices=True, device="cuda:0")
lp.randomize(seed=1)
lp.step()
torch.cuda.synchronize()

# (a) record the C-ABI calls of one step, then replay each one REP times in a graph
calls = []
_real_call = _lib.call


def orig(name, *args):
    """Replays go through the bring-up build (stage masks / experiment bits only exist there)."""
    rc = getattr(_lib.debug_lib(), name)(*args)
    assert rc == 0, (name, rc)
    return rc


def rec(name, *args):
    calls.append((name, args))
    return _real_call(name, *args)


import clc_b200.latent_path as LPm  # noqa: E402
import clc_b200.ops as OPSm  # noqa: E402
LPm.call = rec
OPSm.call = rec
lp.step()
LPm.call = _real_call
OPSm.call = _real_call
torch.cuda.synchronize()
print(f"# {a.workload}: {len(calls)} C-ABI calls per step; per-call time, L2-warm, {a.rep} back-to-back in a graph")
s = torch.cuda.Stream()
tot = 0.0
STAGES = {"clc_match_topk_tc": ["prepass", "gemm", "rescore+blend"], "clc_match_bwd": ["memset+main", "cl_to_nchw"],
          "clc_match_clm_fwd": ["prepass", "gemm", "rescore+blend+clm"], "clc_match_clm_bwd": ["main", "cl_to_nchw"]}
# debug variants: (stage bit, dbg bits << 8, label)
VARIANTS = {"clc_match_topk_tc": [(1, 1, "prepass: query role only"), (1, 2, "prepass: ref role only"),
                                  (2, 1, "gemm: shifts rounded to 8 rows (aligned descriptors)"),
                                  (2, 2, "gemm: no MMAs (TMA + epilogue only)"), (2, 4, "gemm: no A loads"),
                                  (4, 1, "rescore: no re-scoring loads"), (4, 2, "rescore: no blend")],
            "clc_match_bwd": [(1, 1, "main: no window atomics"), (1, 2, "main: no g_q atomics"), (1, 3, "main: no atomics")],
            "clc_match_clm_fwd": [(4, 1, "rescore: no re-scoring loads"), (4, 2, "rescore: no blend")],
            "clc_match_clm_bwd": [(1, 1, "main: no window atomics"), (1, 2, "main: no g_q atomics"), (1, 3, "main: no atomics")]}
expanded = []
for name, args in calls:
    expanded.append((name, args, 0xff, name))
    for i, lab in enumerate(STAGES.get(name, [])):
        expanded.append((name, args, 1 << i, f"  {name}[{lab}]"))
    for st_, dbg, lab in VARIANTS.get(name, []):
        expanded.append((name, args, st_ | (dbg << 8), f"    dbg {lab}"))
for name, args, mask, label in expanded:
    _lib.debug_lib().clc_debug_set_stage_mask(mask)
    with torch.cuda.stream(s):
        st = s.cuda_stream
        args2 = list(args)
        args2[-1] = st                       # the stream is the last argument of every entry point
        for _ in range(2):
            orig(name, *args2)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        args2[-1] = torch.cuda.current_stream().cuda_stream
        for _ in range(a.rep):
            orig(name, *args2)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (a.iters * a.rep)
    if mask == 0xff:
        tot += us
    print(f"{label:40s} {us:8.2f} us/call")
_lib.debug_lib().clc_debug_set_stage_mask(0xff)
print(f"{'sum':28s} {tot:8.2f} us")

# (b) per-kernel trace, warm
per = {}
stream = torch.cuda.current_stream().cuda_stream
for _ in range(a.iters):
    def traced():
        torch.cuda._sleep(1_500_000)
        _lib.lib().clc_trace_mark()
        lp.step()
    for name, ms in _lib.kernel_trace(traced, stream):
        t = per.setdefault(name, [0.0, 0])
        t[0] += ms
        t[1] += 1
print("# per-kernel trace (event after every kernel, host running ahead of the device), L2-warm")
for n, t in per.items():
    print(f"{n:36s} {1e3 * t[0] / t[1]:8.2f} us x{t[1] / a.iters:g}")

# (c) the two graphs of the end-to-end path, each alone (L2-warm)
lp.capture_split()
for name, g in (("graph: match chain", lp._g_match), ("graph: entropy chains (forked)", lp._g_entropy)):
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / a.iters:8.2f} us/replay")
lp.capture(fork=True)
for name in ("graph: full step (forked)",):
    lp.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        lp.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 1e3 / a.iters:8.2f} us/replay")
