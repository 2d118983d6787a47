#!/usr/bin/env python
"""bench.py -- Mpix/s of the CLC latent path (match + CLM + entropy stage, fwd+bwd) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the
                                                             # reference path on the host cores
Under torchrun (N>1) every rank runs the same per-GPU batch (weak scaling, images sharded
data-parallel); rank 0 prints ONE JSON line.

A "step" = one pass of clc_b200.latent_path.LatentPath over one batch of synthetic inputs:
  value : inputs resident in HBM, CUDA-event timed per step with an L2 flush between steps.
  e2e   : the same path from HOST buffers (clc_b200.latent_path.HostPipeline): every step uploads its inputs
          from pinned host memory and reads its bpp back; double-buffered, so the upload of step i+1 overlaps
          the kernels of step i.  The single-buffered step_host and the per-operator autograd modules are
          reported next to it.
  roofline : the dominant C-ABI call of the step, algorithmic bytes / its mean
          CUDA-event duration inside the same run, vs MEASURED_PEAKS.json.
  cpu_baseline : the oracle port (oracle/latent_path_oracle.py) timed on this box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (per-GPU batch, H, W, n_refs, train)      BASELINE.json configs[...]
    "cfg1": dict(B=1, H=256, W=256, R=3, train=False, desc="configs[0]: 1x3x256x256, 3 refs, forward"),
    "cfg2": dict(B=8, H=256, W=256, R=3, train=True, desc="configs[1]: training step, batch 8 of 256x256, n_refs=3"),
    "cfg3": dict(B=3, H=512, W=768, R=3, train=False, desc="configs[2]: Kodak-shaped 768x512 inference, 3 images per GPU"),
    "cfg4": dict(B=1, H=1280, W=2048, R=3, train=False, desc="configs[3]: CLIC-shaped 2048x1280 inference, 3-ref matching"),
    # configs[4]: n_refs sweep 1/3/5, data-parallel training, GLOBAL batch 64 -> 64 / n_gpus images per GPU
    # (strong scaling in the batch; --n-refs picks the sweep point)
    "cfg5": dict(B=64, H=256, W=256, R=3, train=True, global_batch=64,
                 desc="configs[4]: data-parallel training, global batch 64 of 256x256, n_refs sweep 1/3/5"),
}


def workload(args, world):
    """The per-GPU problem of this run: BASELINE.json config + the --n-refs / world-size dependent parts."""
    cfg = dict(WORKLOADS[args.workload])
    if args.n_refs is not None:
        cfg["R"] = args.n_refs
    if "global_batch" in cfg:
        cfg["B"] = max(1, cfg["global_batch"] // world)
    return cfg


def make_config(args, cfg, world):
    """`config` of the JSON line -- the SAME dict for both arms (`--impl ours` / `--impl reference`)."""
    fused = not args.per_slice
    use_graph = not args.no_graph
    return {"workload": f"{args.workload}: {cfg['desc']}", "per_gpu_batch": cfg["B"], "image": [cfg["H"], cfg["W"]],
            "n_refs": cfg["R"], "pass": "fwd+bwd" if cfg["train"] else "fwd", "match_mode": args.match_mode,
            "slice_launches": "fused" if fused else "per-slice", "cuda_graph": use_graph,
            "graph_branches": "serial" if args.no_fork else ("match | hyper -> slices" if world == 1
                                                              else "match | hyper | slices -> all-reduce"),
            "l2": "flushed between timed iterations (256 MB write)", "patch": 4, "k": 4,
            "collective": ("none" if world == 1 else
                           ("one-shot all-reduce kernel over NVLink peer memory" if args.collective == "peer" else "NCCL all-reduce")),
            "noise": "uploaded tensors" if args.host_noise else "generated in-kernel (Philox), like the reference's "
                     "on-device uniform_()"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_time(cfg, budget_s, steps=None, warmup=1, seed=1):
    """Time the oracle port of the path on the host cores.  Returns dict(value Mpix/s, ...)."""
    from oracle import latent_path_oracle as LO
    from oracle import clc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, H, W, R, train = cfg["B"], cfg["H"], cfg["W"], cfg["R"], cfg["train"]
    # bounded sample: shrink the batch (images are independent units) until one step is cheap
    est = 0.06 * (H * W / 65536.0) ** 1.6 * R / 3.0      # rough s / image on ~8-32 cores
    Bs = max(1, min(B, int(max(1.0, budget_s / 6.0 / max(est, 1e-3)))))
    g = torch.Generator().manual_seed(seed)
    h, w = H // 16, W // 16
    y = 3 * torch.randn(Bs, 320, h, w, generator=g)
    inp = dict(y=y, z=2 * torch.randn(Bs, 192, H // 64, W // 64, generator=g),
               refs=torch.randn(Bs, R, 320, h, w, generator=g) + (y / 6).unsqueeze(1),
               mu=torch.randn(Bs, 320, h, w, generator=g),
               scale=torch.exp(torch.empty(Bs, 320, h, w).uniform_(math.log(0.05), math.log(300.0), generator=g)),
               lrp=torch.randn(Bs, 320, h, w, generator=g), att=torch.randn(Bs, R, 1, h, w, generator=g),
               noise_y=torch.rand(Bs, 320, h, w, generator=g) - 0.5,
               noise_z=torch.rand(Bs, 192, H // 64, W // 64, generator=g) - 0.5,
               g_y_hat=1e-3 * torch.randn(Bs, 320, h, w, generator=g),
               g_fused=1e-3 * torch.randn(Bs, 320, h, w, generator=g))
    eb = O.EntropyBottleneck(192)
    for _ in range(warmup):
        LO.step(inp, eb, train=train)
        eb.zero_grad()
    times = []
    t_all = time.perf_counter()
    n = 0
    while True:
        t0 = time.perf_counter()
        LO.step(inp, eb, train=train)
        eb.zero_grad()
        times.append(time.perf_counter() - t0)
        n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and (time.perf_counter() - t_all > budget_s or n >= 50) and n >= 2:
            break
    t = sum(times) / len(times)
    return dict(value=Bs * H * W / t / 1e6, unit="Mpix/s", cores=cores, kind="port",
                sample=f"{Bs} of {B} images/step ({H}x{W}, {R} refs, {'fwd+bwd' if train else 'fwd'}), "
                       f"{n} timed steps, oracle port on torch CPU ({cores} threads)",
                ms_per_step=t * 1e3, steps=n, images=Bs)


class PublicPath:
    """The same operator sequence through the PUBLIC autograd API (what a user of the drop-in
    modules calls), used for the end-to-end number."""

    def __init__(self, cfg, device, match_mode):
        import clc_b200
        self.c = clc_b200
        self.cfg, self.dev, self.mode = cfg, device, match_mode
        self.gc = clc_b200.GaussianConditional(None).to(device).train(cfg["train"])
        self.eb = clc_b200.EntropyBottleneck(192).to(device).train(cfg["train"])
        self.npix = cfg["B"] * cfg["H"] * cfg["W"]
        self.functional_grads = False

    def step(self, d):
        from clc_b200 import ops
        c, train = self.c, self.cfg["train"]
        B, R = d["refs"].shape[0], d["refs"].shape[1]
        grad = ("y", "z", "refs", "mu", "scale", "lrp", "att")
        t = {k: (v.requires_grad_(True) if (train and k in grad) else v) for k, v in d.items()}
        aligned = c.match_and_gather(t["y"], t["refs"], 4, 4, 4, 15.0, True, False, self.mode)   # [B,R,C,h,w]
        fused = c.clm_fuse(aligned.transpose(0, 1), t["att"].transpose(0, 1), t["y"])
        _, lik_z, z_hat = self.eb(t["z"], noise=t.get("noise_z") if train else None, ste=True, want_outputs=False)
        liks, yh = [], []
        for i in range(5):
            sl = slice(64 * i, 64 * (i + 1))
            _, lik, y_hat = self.gc(t["y"][:, sl], t["scale"][:, sl], t["mu"][:, sl],
                                    noise=t["noise_y"][:, sl] if (train and "noise_y" in t) else None, ste=True,
                                    want_outputs=False)
            yh.append(ops.lrp_add_(y_hat, t["lrp"][:, sl]))
            liks.append(lik)
        bpp = -(ops.log2_sum(torch.cat(liks, 1)) + ops.log2_sum(lik_z)) / self.npix
        if train:
            loss = bpp.float() + (torch.cat(yh, 1) * t["g_y_hat"]).sum() + (fused * t["g_fused"]).sum()
            if self.functional_grads:
                # graph capture: gradients as returned tensors (no AccumulateGrad nodes tied to another stream)
                self.grads = torch.autograd.grad(loss, [t[k] for k in grad])
            else:
                loss.backward()
            return loss
        return bpp


class GraphedPublicPath:
    """The module-API step (PublicPath.step: forward AND backward through the drop-in autograd modules) captured
    once into a CUDA graph, the standard PyTorch recipe for a launch-bound training step (whole-network
    capture): per step the host inputs are copied into the graph's static tensors, the graph is replayed and the
    result read back.  In-kernel noise: the captured step advances the device noise stream, so every replay
    draws a fresh sample."""

    def __init__(self, pub, example):
        from clc_b200 import rng
        self.pub = pub
        pub.functional_grads = True
        self.static = {k: v.detach().clone() for k, v in example.items()}
        dev = next(iter(self.static.values())).device

        def step():
            for v in self.static.values():
                v.grad = None
            out = pub.step(self.static)
            rng.advance(dev)
            return out

        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        for v in self.static.values():
            v.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = pub.step(self.static)
            rng.advance(dev)

    def step(self, host_views):
        with torch.no_grad():
            for n, t in self.static.items():
                t.copy_(host_views[n], non_blocking=True)
        self.graph.replay()
        return self.out


def _make_path(cfg, args, dev, fused, rank):
    from clc_b200.latent_path import LatentPath
    lp = LatentPath(cfg["B"], cfg["H"], cfg["W"], n_refs=cfg["R"], train=cfg["train"], match_mode=args.match_mode,
                    fused_slices=fused, device=dev, data_parallel=True, device_noise=not args.host_noise,
                    collective=args.collective)
    lp.randomize(seed=1 + rank)
    return lp


def _time_steps(run, exchange, flush, K, barrier):
    """K steps, each bracketed by its own CUDA events (the L2 flush sits outside the bracket).
    Returns the summed device time in ms."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record()
        run()
        exchange()
        b.record()
    torch.cuda.synchronize()
    barrier()
    return sum(a.elapsed_time(b) for a, b in ev)


def run_ours(args):
    from clc_b200 import _lib
    from clc_b200 import dist as cdist
    rank, world, local = cdist.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = workload(args, world)
    K, Wm = args.steps, max(args.warmup, 3)
    fused = not args.per_slice
    lp = _make_path(cfg, args, dev, fused, rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    use_graph = not args.no_graph
    launches_per_step_calls = lp.step()  # C-ABI calls per step; also first-touch
    torch.cuda.synchronize()
    l0 = _lib.launches()
    lp.step()
    torch.cuda.synchronize()
    launches_per_step = _lib.launches() - l0          # kernels per step, counted inside the library
    if use_graph:
        lp.capture(fork=not args.no_fork)
    run = lp.replay if use_graph else lp.step

    def barrier():
        if world > 1:
            torch.distributed.barrier()

    def make_exchange(path):
        # the only collective the path owns -- ONE NCCL all-reduce per step of the EB parameter
        # gradients (training) + the 2-double bpp statistic -- is enqueued by LatentPath itself on its entropy
        # branch (captured into the graph), where it overlaps the match chain.
        return lambda: None

    exchange = make_exchange(lp)
    for _ in range(Wm):
        flush.zero_()
        run()
        exchange()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    t_wall = time.perf_counter()
    total_ms = _time_steps(run, exchange, flush, K, barrier)
    t_wall = time.perf_counter() - t_wall
    gpu_launches = launches_per_step * K
    tms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
    total_ms = tms.item()
    pix_per_step = world * cfg["B"] * cfg["H"] * cfg["W"]
    value = pix_per_step * K / (total_ms * 1e-3) / 1e6
    bpp_dev = lp.bpp().item()
    n_uncert = lp.n_uncertified() if args.match_mode == "tc" else None

    # ---- L2-warm variant (no flush), for context -------------------------------------------
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        run()
    b.record()
    torch.cuda.synchronize()
    warm_ms = a.elapsed_time(b) / K

    # ---- the other slice-launch granularity, same timing rules, for context -----------------
    lp2 = _make_path(cfg, args, dev, not fused, rank)
    lp2.step()
    torch.cuda.synchronize()
    if use_graph:
        lp2.capture(fork=not args.no_fork)
    run2 = lp2.replay if use_graph else lp2.step
    ex2 = make_exchange(lp2)
    for _ in range(Wm):
        flush.zero_()
        run2()
        ex2()
    t2 = torch.tensor([_time_steps(run2, ex2, flush, K, barrier)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t2, op=torch.distributed.ReduceOp.MAX)
    value_other = pix_per_step * K / (t2.item() * 1e-3) / 1e6
    del lp2

    # ---- end to end, host buffers: H2D of the step's inputs + D2H of its result inside the timed
    # region.  First the single-buffered LatentPath.step_host (pinned flat staging buffer, two graphs, upload of the
    # entropy inputs overlapped with the match chain).  Context: the same operator sequence through
    # the per-operator autograd modules (eager, per-slice calls, as a model makes them).
    host_flat, host_views = lp.host_staging()
    for n, v in host_views.items():
        v.copy_(getattr(lp, n))
    h2d = lp.h2d_bytes_per_step()
    lp.capture_split()
    for _ in range(Wm):
        lp.step_host(host_flat)
    torch.cuda.synchronize()
    barrier()
    e2e_ev = []
    for _ in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        bpp_host = lp.step_host(host_flat)          # uploads, runs, reads the result back (synchronises)
        b.record()
        e2e_ev.append((a, b))
    torch.cuda.synchronize()
    e2e1_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e_ev)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e1_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_single_value = pix_per_step * K / (e2e1_ms.item() * 1e-3) / 1e6

    # Headline: the double-buffered driver (HostPipeline): two LatentPath instances alternate, the upload of
    # step i+1 overlaps the kernels + read-back of step i.  Every step uploads its inputs from pinned host
    # memory and reads its bpp back; the L2 flush of each step is enqueued on the compute stream INSIDE the
    # timed region; one event pair brackets the K steps.
    from clc_b200.latent_path import HostPipeline
    lpb = _make_path(cfg, args, dev, fused, rank)
    lpb.step()
    host_flat_b, views_b = lpb.host_staging()
    for n, v in views_b.items():
        v.copy_(getattr(lpb, n))
    pipe = HostPipeline([lp, lpb])
    flats = [host_flat, host_flat_b]
    for i in range(Wm + (Wm & 1)):
        pipe.submit(i, flats[i % 2], before_compute=flush.zero_)
        pipe.result(i)
    torch.cuda.synchronize()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        pipe.submit(i, flats[i % 2], before_compute=flush.zero_)
        if i:
            bpp_host = pipe.result(i - 1)
    bpp_host = pipe.result(K - 1)
    b.record()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_value = pix_per_step * K / (e2e_ms.item() * 1e-3) / 1e6
    del lpb, pipe

    pub = PublicPath(cfg, dev, args.match_mode)
    names = lp.step_inputs
    for _ in range(Wm):
        d = {n: host_views[n].to(dev, non_blocking=True) for n in names}
        pub.step(d).item()
    torch.cuda.synchronize()
    barrier()
    mod_ev = []
    for _ in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        d = {n: host_views[n].to(dev, non_blocking=True) for n in names}
        res = pub.step(d)
        res_host = res.item()          # device -> host read of the step's result
        b.record()
        mod_ev.append((a, b))
    torch.cuda.synchronize()
    mod_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in mod_ev)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(mod_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_modules_value = pix_per_step * K / (mod_ms.item() * 1e-3) / 1e6

    # the same module-API step, captured once into a CUDA graph (whole-step capture, fwd + bwd)
    gpub = GraphedPublicPath(pub, {n: host_views[n].to(dev) for n in names})
    for _ in range(Wm):
        gpub.step(host_views).item()
    torch.cuda.synchronize()
    barrier()
    gmod_ev = []
    for _ in range(K):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res_host = gpub.step(host_views).item()          # uploads + replay + device -> host read of the result
        b.record()
        gmod_ev.append((a, b))
    torch.cuda.synchronize()
    gmod_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in gmod_ev)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(gmod_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_gmodules_value = pix_per_step * K / (gmod_ms.item() * 1e-3) / 1e6
    del gpub
    clocks = sampler.stop()

    # ---- per-kernel pass: the library records a CUDA event after EVERY kernel of the same step ---
    per = {}
    stream = torch.cuda.current_stream().cuda_stream
    def traced_step():
        # a spin kernel first, so the host enqueues the whole step while the device is still busy and
        # the kernels then run back to back (event-to-event time = kernel time, not host launch time)
        torch.cuda._sleep(1_500_000)
        _lib.lib().clc_trace_mark()
        lp.step()

    for _ in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        for name, ms in _lib.kernel_trace(traced_step, stream):
            t = per.setdefault(name, [0.0, 0])
            t[0] += ms
            t[1] += 1
    pk = peaks()
    work = lp.algorithmic_work()
    gathered = lp.gathered_bytes()
    breakdown = {n: {"us_per_step": 1e3 * t[0] / K, "launches_per_step": t[1] / K, "us_per_launch": 1e3 * t[0] / t[1]}
                 for n, t in per.items()}
    traffic_db = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic_db = json.load(open(tp))

    def roof(name):
        us = 1e3 * per[name][0] / per[name][1]
        kind, amount = work.get(name, ("bytes", None))
        traffic = traffic_db.get(f"{args.workload}:{name}")
        if kind == "flops":
            ach = amount / (us * 1e-6) / 1e12
            return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["tf"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf"], "traffic": traffic, "us_per_launch": us, "peak_source": pk["src"],
                    "algorithmic_flop_per_launch": amount}
        ach = (amount / (us * 1e-6) / 1e9) if amount else None
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": (ach / pk["hbm"]) if ach else None, "traffic": traffic, "us_per_launch": us,
                "peak_source": pk["src"], "algorithmic_bytes_per_launch": amount,
                "l2_bytes": gathered.get(name)}

    # dominant kernel = largest share of the step; kernels within 2 % of the largest are tied and the tie goes
    # to name order, so the headline does not flip between runs
    top = max(t[0] for t in per.values())
    dom = sorted(n for n in per if per[n][0] >= 0.98 * top)[0]
    rooflines = {n: roof(n) for n in per}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_time(cfg, budget_s=args.cpu_budget)
    this_mode, other_mode = ("fused", "per-slice") if fused else ("per-slice", "fused")
    line = {
        "metric": "Mpix/s of CLC latent path (match+CLM+entropy)", "value": value, "unit": "Mpix/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "strong" if "global_batch" in cfg else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": make_config(args, cfg, world),
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                "api": "clc_b200.latent_path.HostPipeline (double-buffered): per step pinned host staging buffer -> 2 "
                       "uploads -> L2 flush -> match graph || entropy graph -> bpp read back; upload of step i+1 "
                       "overlaps the kernels of step i",
                "ms_per_step": e2e_ms.item() / K, "bpp": bpp_host,
                "single_buffered": {"value": e2e_single_value, "ms_per_step": e2e1_ms.item() / K,
                                    "api": "clc_b200.LatentPath.step_host (no overlap between steps)"},
                "autograd_modules": {"value": e2e_modules_value, "ms_per_step": mod_ms.item() / K,
                                     "api": "clc_b200 per-operator autograd modules, eager, per-slice calls"},
                "autograd_modules_graphed": {"value": e2e_gmodules_value, "ms_per_step": gmod_ms.item() / K,
                                             "api": "the same per-operator autograd modules (fwd + bwd), step captured "
                                                    "once with torch.cuda.graph and replayed"}},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "roofline": rooflines[dom],
        "cpu_baseline": cpu,
        f"value_{other_mode.replace('-', '_')}_launches": value_other,
        "l2_warm_ms_per_step": warm_ms,
        "kernel_launches_per_step": launches_per_step,
        "abi_calls_per_step": launches_per_step_calls,
        "breakdown": breakdown,
        "rooflines": rooflines,
        "bpp": bpp_dev,
        "n_uncertified": n_uncert,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line))


def run_model_step(args):
    """Context run (NOT the latent-path metric): the reference's data-parallel TRAINING step at model level
    (train_CLC.py:119-184 with CLC(N=128, num_ref_frames=n), AdamW, grad clip 1.0, aux loss), one process per GPU,
    bucketed gradient all-reduce (clc_b200.dist.GradAllReducer, NCCL over NVLink) overlapped with the backward.
    Reports the step time, the same step without the all-reduce, and the bare all-reduce of the same buckets, i.e.
    how much of the exchange the backward hides."""
    from clc_b200 import dist as cdist
    from clc_b200.loss import RateDistortionLoss
    from clc_b200.models import CLC
    rank, world, local = cdist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = workload(args, world)
    B, H, W, R = cfg["B"], cfg["H"], cfg["W"], cfg["R"]
    torch.manual_seed(0)
    model = CLC(N=128, num_ref_frames=R).to(dev).train()
    crit = RateDistortionLoss(lmbda=0.013)
    main = [p for n, p in model.named_parameters() if not n.endswith(".quantiles")]
    aux = [p for n, p in model.named_parameters() if n.endswith(".quantiles")]
    opt, aux_opt = torch.optim.AdamW(main, lr=1e-4), torch.optim.AdamW(aux, lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    x = torch.rand(B, 3, H, W, device=dev, generator=g)
    refs = [torch.rand(B, 3, H, W, device=dev, generator=g) for _ in range(R)]

    def step(red):
        if red is not None:
            red.zero_grad()                              # gradients live in the communication buckets
        else:
            opt.zero_grad(set_to_none=True)
            aux_opt.zero_grad(set_to_none=True)
        out = model(x, refs)
        loss = crit(out, x)["loss"]
        loss.backward()
        model.aux_loss().backward()                     # (train_CLC.py:181-183; its 192x3 gradients ride the last bucket)
        if red is not None:
            red.finish()
        torch.nn.utils.clip_grad_norm_(main, 1.0)
        opt.step()
        aux_opt.step()
        return loss

    def timed(red, K, Wm):
        for _ in range(Wm):
            step(red)
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            loss = step(red)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / K], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return t.item(), loss.item()

    K, Wm = args.steps, max(args.warmup, 2)
    ms_nodp, _ = timed(None, K, Wm)                       # local step, no exchange (each rank drifts apart: timing only)
    red = cdist.GradAllReducer(model, bucket_mb=args.bucket_mb) if world > 1 else None
    ms_dp, loss = timed(red, K, Wm)
    n_grad = sum(p.grad.numel() for p in model.parameters() if p.grad is not None)
    ms_ar = None
    if world > 1:
        flat = torch.zeros(n_grad, device=dev)
        for _ in range(3):
            torch.distributed.all_reduce(flat)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            torch.distributed.all_reduce(flat)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_ar = t.item()
    if rank != 0:
        return
    pix = world * B * H * W
    exposed = max(ms_dp - ms_nodp, 0.0)
    line = {"impl": "ours", "mode": "model-step (context, not the latent-path metric)",
            "metric": "Mpix/s of the CLC(N=128) data-parallel training step", "value": pix / (ms_dp * 1e-3) / 1e6,
            "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_dp,
            "ms_per_step_without_allreduce": ms_nodp, "higher_is_better": True,
            "scaling": "strong" if "global_batch" in cfg else "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['desc']}", "model": "CLC(N=128, M=320)", "per_gpu_batch": B,
                       "global_batch": B * world, "image": [H, W], "n_refs": R, "optimizer": "AdamW + aux AdamW, clip 1.0",
                       "bucket_mb": args.bucket_mb},
            "grad_allreduce": {"elements": n_grad, "bytes": 4 * n_grad, "bare_ms": ms_ar, "exposed_ms": exposed,
                               "hidden_fraction": (1.0 - exposed / ms_ar) if ms_ar else None,
                               "bus_gbs": (2 * (world - 1) / world * 4 * n_grad / (ms_ar * 1e-3) / 1e9) if ms_ar else None},
            "loss": loss}
    print(json.dumps(line))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg = workload(args, world)
    r = cpu_port_time(cfg, budget_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "Mpix/s of CLC latent path (match+CLM+entropy)", "value": r["value"],
            "unit": "Mpix/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if "global_batch" in cfg else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(args, cfg, world),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--n-refs", type=int, default=None, help="references per image (configs[4] sweeps 1/3/5)")
    ap.add_argument("--match-mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--per-slice", action="store_true",
                    help="launch the GaussianConditional / LRP kernels once per channel slice (the model's call "
                         "pattern) instead of once over all slices (the isolated path's all-slices entry point)")
    ap.add_argument("--host-noise", action="store_true",
                    help="upload the U(-1/2,1/2) quantisation noise as tensors (bit-reproducible parity runs) instead "
                         "of generating it inside the kernels, as the reference generates it on the device")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="the latent path's per-step exchange: one-shot kernel over NVLink peer memory, or NCCL all-reduce")
    ap.add_argument("--model-step", action="store_true",
                    help="context run: CLC(N=128) model-level data-parallel training step with the bucketed gradient "
                         "all-reduce (not the latent-path metric)")
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fork", action="store_true", help="capture the step as one serial chain instead of two branches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.model_step:
        run_model_step(args)
    else:
        run_ours(args)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
