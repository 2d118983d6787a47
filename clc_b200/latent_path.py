"""The CLC conditional-latent hot path as ONE object over preallocated HBM buffers, driven
through the raw C ABI (no autograd, no allocation per step, CUDA-graph capturable):

    match (Pearson correlation GEMM + top-k)  ->  gather/blend  ->  CLM fusion
    -> EntropyBottleneck(z) -> 5 x [GaussianConditional + STE round, LRP add] -> bpp partial sums
    and the backward of all of it.

This is what bench.py times (`value`): the operator sequence of SURVEY.md 8a with the
parameter networks' outputs (mu / scale / lrp, attention logits, upstream gradients) supplied
as inputs, exactly as the reference's slice loop (CLC_run.py:535-590) sees them.  Training
through real models uses the autograd wrappers (ops.py, matching.py, clm.py) instead.
"""
import ctypes as C
import math

import torch

from . import matching, ops
from ._lib import call, lib, ptr


class LatentPath:
    """Buffers + launch sequence for a batch of B images of H x W pixels with n_refs references."""

    INPUT_NAMES = ("y", "z", "refs", "mu", "scale", "lrp", "att", "noise_y", "noise_z", "g_y_hat", "g_fused")
    MATCH_INPUTS = ("y", "refs", "att", "g_fused")                       # what match -> gather -> CLM reads
    ENTROPY_INPUTS = ("z", "mu", "scale", "lrp", "noise_y", "noise_z", "g_y_hat")
    TRAIN_ONLY = ("noise_y", "noise_z", "g_y_hat", "g_fused")

    def __init__(self, B, H, W, n_refs=3, M=320, num_slices=5, z_channels=192, train=True, patch=4, k=4,
                 temperature=15.0, match_mode="tc", gaussian_mask=True, fused_slices=False,
                 device="cuda", lmbda=0.013, data_parallel=False, device_noise=False, fuse_chain=True, collective="peer"):
        assert H % 64 == 0 and W % 64 == 0, "latent geometry: h = H/16, hz = H/64"
        self.B, self.H, self.W, self.R, self.M = B, H, W, n_refs, M
        self.h, self.w = H // 16, W // 16
        self.hz, self.wz = H // 64, W // 64
        self.num_slices, self.Cs = num_slices, M // num_slices
        self.train, self.patch, self.k, self.T = train, patch, k, float(temperature)
        self.match_mode, self.fused_slices = match_mode, fused_slices
        self.device_noise = bool(device_noise) and train      # in-kernel Philox noise instead of noise tensors
        self.P = (self.h // patch) * (self.w // patch)
        self.corr_w = self.w - patch + 1
        self.num_pixels = B * H * W
        self.bpp_coef = 1.0 / (-math.log(2.0) * self.num_pixels)  # dL/dlik = bpp_coef / lik
        dev = torch.device(device)
        self.device = dev
        f = dict(dtype=torch.float32, device=dev)
        h, w, R = self.h, self.w, n_refs
        # ---- inputs: ONE flat device buffer (views per tensor), match-chain inputs first so the
        # host->device upload of a step can be split into two contiguous copies (step_host) ----
        shapes = {"y": (B, M, h, w), "refs": (B, R, M, h, w), "att": (B, R, 1, h, w), "g_fused": (B, M, h, w),
                  "z": (B, z_channels, self.hz, self.wz), "mu": (B, M, h, w), "scale": (B, M, h, w),
                  "lrp": (B, M, h, w), "noise_y": (B, M, h, w), "noise_z": (B, z_channels, self.hz, self.wz),
                  "g_y_hat": (B, M, h, w)}
        # (device_noise: the U(-1/2, 1/2) samples are generated inside the kernels -- as the reference draws them
        # on the device -- so no noise tensors exist or are uploaded)
        skip = set() if train else set(self.TRAIN_ONLY)
        if self.device_noise:
            skip |= {"noise_y", "noise_z"}
        self.step_inputs = [n for n in self.MATCH_INPUTS + self.ENTROPY_INPUTS if n not in skip]
        offs, off = {}, 0
        for n in self.MATCH_INPUTS + self.ENTROPY_INPUTS:
            if n == self.ENTROPY_INPUTS[0]:
                self._n_match_in = off
            offs[n] = off
            off += (math.prod(shapes[n]) + 63) // 64 * 64       # 256-byte aligned segments
            if n in skip:
                off = offs[n]                                   # not part of a step's inputs: not uploaded
        self._in = torch.zeros(off, **f)
        self._in_offsets, self._in_shapes = offs, shapes
        for n in self.MATCH_INPUTS + self.ENTROPY_INPUTS:
            if n in self.step_inputs:
                setattr(self, n, self._in[offs[n]:offs[n] + math.prod(shapes[n])].view(shapes[n]))
            elif n in ("noise_y", "noise_z") and self.device_noise:
                setattr(self, n, None)
            else:
                setattr(self, n, torch.empty(shapes[n], **f))   # unused in eval; kept for API symmetry
        # ---- EntropyBottleneck parameters ----
        from .entropy_models import EntropyBottleneck
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(0)                              # EB biases are U(-.5,.5): keep runs comparable
            eb = EntropyBottleneck(z_channels)
        eb = eb.to(dev)
        self.eb_m = [getattr(eb, f"_matrix{i}").detach() for i in range(5)]
        self.eb_b = [getattr(eb, f"_bias{i}").detach() for i in range(5)]
        self.eb_f = [getattr(eb, f"_factor{i}").detach() for i in range(4)]
        self.quantiles = eb.quantiles.detach()
        # ---- outputs ----
        self.val = torch.empty(B * R, self.P, k, **f)
        self.idx = torch.empty(B * R, self.P, k, dtype=torch.int32, device=dev)
        self.weights = torch.empty(B * R, self.P, k, **f)
        # patches whose bf16-screened candidate set could not be certified (tc mode; set by every match call)
        self.n_uncert = torch.zeros(1, dtype=torch.int32, device=dev)
        self.aligned = torch.empty(B, R, M, h, w, **f)
        self.fused = torch.empty(B, M, h, w, **f)
        self.lik_z = torch.empty_like(self.z)
        self.z_hat = torch.empty_like(self.z)
        self.lik_y = torch.empty(B, M, h, w, **f)
        self.y_hat = torch.empty(B, M, h, w, **f)
        # backward outputs
        self.g_y = torch.empty(B, M, h, w, **f)
        self.g_mu = torch.empty(B, M, h, w, **f)
        self.g_scale = torch.empty(B, M, h, w, **f)
        self.g_lrp = torch.empty(B, M, h, w, **f)
        self.g_z = torch.empty_like(self.z)
        self.g_aligned = torch.empty(B, R, M, h, w, **f)
        self.g_att = torch.empty(B, R, 1, h, w, **f)
        self.g_val = torch.empty(B * R, self.P, k, **f)
        # ---- accumulators: one flat buffer; each chain zeroes its own segment with one memset ----
        n_eb = sum(t.numel() for t in self.eb_m + self.eb_b + self.eb_f)
        n_acc = 4 + n_eb + B * M * h * w                       # 2 doubles + EB grads + g_q
        self._acc = torch.zeros(n_acc, **f)
        self.log2 = self._acc[:4].view(torch.float64)          # [sum log2 lik_y, sum log2 lik_z]
        self._acc_y = self._acc[0:2]                           # slice-loop chain
        self._acc_eb = self._acc[2:4 + n_eb]                   # hyper-latent chain (log2[1] + EB grads)
        self._acc_match = self._acc[4 + n_eb:]                 # match chain (g_q; g_refs is overwritten)
        off = 4
        self.g_eb = []
        for t in self.eb_m + self.eb_b + self.eb_f:
            self.g_eb.append(self._acc[off:off + t.numel()].view_as(t))
            off += t.numel()
        self.g_refs = torch.empty(B, R, M, h, w, **f)
        self.g_q = self._acc[off:off + B * M * h * w].view(B, M, h, w)
        # ---- match workspace ----
        self.mask = matching._cached_mask(h, w, patch, patch, dev) if gaussian_mask else None
        self.gaussian_mask = gaussian_mask
        if match_mode == "tc":
            nb = lib().clc_match_topk_tc_workspace_bytes(B * R, R, M, h, w, patch, patch, k)
            self.corr = None
        elif match_mode == "fp32":
            nb = lib().clc_pearson_corr_workspace_bytes(B * R, self.P, M, patch, patch, h, w)
            self.corr = torch.empty(B * R, self.P, (h - patch + 1) * self.corr_w, **f)
        else:
            raise ValueError(f'Invalid match mode "{match_mode}"')
        self.ws = torch.empty(max(int(nb), 16), dtype=torch.uint8, device=dev)
        nbb = lib().clc_match_bwd_workspace_bytes(B * R, M, h, w) if train else 0
        self.ws_bwd = torch.empty(max(int(nbb), 16), dtype=torch.uint8, device=dev)
        # channels-last fp32 copy of the references that the forward leaves in its workspace
        self._r_cl = (lib().clc_match_topk_tc_ref_cl(ptr(self.ws), B * R, R, M, h, w, patch, patch, k)
                      if match_mode == "tc" else None)
        # shapes the CLM-fused match backward kernel covers (clc_match_clm_bwd); others use the two-call sequence
        self._fused_bwd = (bool(fuse_chain) and match_mode == "tc" and patch == 4 and k <= 4 and n_refs <= 8 and
                           M % 4 == 0 and patch * (M // 4) in (128, 192, 256, 320, 384) and w % 4 == 0)
        # forward CLM fusion (clusters of R CTAs per (image, patch)) pays only for small latents: measured on B200
        # (scripts/chain_timing.py [--no-fuse | --fuse-all]) it wins ~2 us at 16 x 16 latents (cfg2) but loses
        # 12 us at 32 x 48 (cfg3: 121 vs 109 us) and 16 us at 80 x 128 (cfg4: 317 vs 301 us) -- with thousands of
        # clusters, each gated on its slowest CTA, the barrier costs more than the 5 x C x H x W floats of HBM
        # traffic the fusion saves
        # (fuse_chain="all" forces it at any size: tests, measurements)
        self._fused_fwd = (bool(fuse_chain) and match_mode == "tc" and n_refs <= 8 and patch * patch <= 64 and
                           (h * w <= 512 or fuse_chain == "all"))
        self._qview = matching._patch_view_from_image(self.y, patch, patch, R)
        self._gqview = self._qview
        self._graph = self._g_match = self._g_entropy = None
        self._streams = None
        self._zs = None
        # data-parallel training: the only collective the path owns is ONE NCCL all-reduce per
        # step (EntropyBottleneck parameter gradients + the 2-double bpp statistic, see _exchange_stats),
        # enqueued on the entropy branch, i.e. it overlaps the (longer) match chain.
        self.data_parallel = bool(data_parallel) and torch.distributed.is_available() and \
            torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        self._eb_grads_flat = self._acc[4:4 + n_eb]
        self._dp_buf = torch.zeros(2 + n_eb, dtype=torch.float64, device=dev) if self.data_parallel else None
        self._world = torch.distributed.get_world_size() if self.data_parallel else 1
        # the exchange itself: one-shot kernel over NVLink peer memory (default on NCCL / one node), or NCCL
        self._peer = None
        if self.data_parallel and collective == "peer" and self._world <= 8:
            from .dist import PeerAllReduce
            self._peer = PeerAllReduce(2, n_eb if train else 0, dev)
        # device-resident noise stream {seed, base offset} (clc_gc_fwd_rng) + the step's bpp (clc_bpp_finalize)
        self.rng_state = torch.tensor([0x5DEECE66D + 7919 * (id(self) & 0xFFFF), 0], dtype=torch.int64, device=dev)
        self._bpp_dev = torch.zeros(1, dtype=torch.float64, device=dev)

    # -------------------------------------------------------------------------------------
    def randomize(self, seed=1):
        """Synthetic inputs of SURVEY.md 8d (operator-level distributions), generated on the
        device: y~N(0,9), mu~N(0,1), scale log-uniform [0.05,300], ref = 0.5*latent + N(0,1)."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        rn = lambda t: t.normal_(generator=g)
        rn(self.y).mul_(3.0)
        rn(self.z).mul_(2.0)
        rn(self.refs).add_((self.y / 6.0).unsqueeze(1))   # ref = 0.5 * (y / 3) + N(0,1)
        rn(self.mu)
        self.scale.uniform_(math.log(0.05), math.log(300.0), generator=g).exp_()
        rn(self.lrp)
        rn(self.att)
        if not self.device_noise:
            self.noise_y.uniform_(-0.5, 0.5, generator=g)
            self.noise_z.uniform_(-0.5, 0.5, generator=g)
        rn(self.g_y_hat).mul_(1e-3)
        rn(self.g_fused).mul_(1e-3)

    def inputs(self):
        return {n: getattr(self, n) for n in self.INPUT_NAMES if getattr(self, n) is not None}

    _RNG_STRIDE = 1 << 32

    def _noise(self, name, call_index, sl=None):
        """Noise argument of one kernel launch: the uploaded tensor (slice), or an rng ticket."""
        if not self.train:
            return None
        if self.device_noise:
            return (self.rng_state, call_index * self._RNG_STRIDE)
        t = getattr(self, name)
        return t if sl is None else sl(t)

    def _zero(self, t):
        call("clc_zero", ptr(t), t.numel() * t.element_size(), ops._stream())

    # ---- host staging (end-to-end path) ------------------------------------------------------
    def host_staging(self):
        """(flat pinned fp32 buffer, {name: view}) laid out like the device input buffer, so a
        whole step's inputs are uploaded with two contiguous copies (see step_host)."""
        flat = torch.empty(self._in.numel(), dtype=torch.float32).pin_memory()
        views = {n: flat[self._in_offsets[n]:self._in_offsets[n] + math.prod(self._in_shapes[n])]
                 .view(self._in_shapes[n]) for n in self.step_inputs}
        return flat, views

    def load_inputs(self, host):
        """Host (pinned) -> device copies of one step's inputs, one per tensor; returns bytes copied."""
        n = 0
        for name in self.step_inputs:
            dst = getattr(self, name)
            dst.copy_(host[name], non_blocking=True)
            n += dst.numel() * 4
        return n

    def h2d_bytes_per_step(self):
        return sum(math.prod(self._in_shapes[n]) for n in self.step_inputs) * 4

    # -------------------------------------------------------------------------------------
    def _slices(self, t):
        if self.fused_slices:
            return [t]
        return list(t.chunk(self.num_slices, 1))

    def match_chain(self):
        """match -> gather/blend -> CLM fusion (and their backward when training), enqueued on the
        current stream (the backward's zero-fills on a side stream / graph branch).  Reads MATCH_INPUTS
        only."""
        st = ops._stream()
        B, R, M, h, w, p, k = self.B, self.R, self.M, self.h, self.w, self.patch, self.k
        S = h * w
        if self.train:
            # zero-fills needed only by the backward (g_q accumulator, gradient scratch) go on a side
            # stream / graph branch: they run under the forward kernels instead of ahead of them
            cur = torch.cuda.current_stream(self.device)
            zs = self._zero_stream()
            zs.wait_stream(cur)
            with torch.cuda.stream(zs):
                self._zero(self._acc_match)
                call("clc_match_bwd_zero_workspace", ptr(self.ws_bwd), self.ws_bwd.numel(), B * R, M, h, w,
                     ops._stream())
        # 1. match: masked Pearson correlation + top-k over all B*R (image, reference) problems
        r = self.refs.view(B * R, M, h, w)
        if self.match_mode == "tc" and self._fused_fwd:
            # screening GEMM -> exact re-scoring + top-k + softmax + gather/blend + CLM fusion, one call
            call("clc_match_clm_fwd", ptr(self.y), ptr(r), B * R, R, M, h, w, p, p, k,
                 1 if self.gaussian_mask else 0, ptr(self.val), ptr(self.idx), ptr(self.n_uncert), self.T,
                 ptr(self.aligned), ptr(self.weights), ptr(self.att), S, R * S, ptr(self.fused), ptr(self.ws),
                 self.ws.numel(), st)
            n = 1
        elif self.match_mode == "tc":
            # screening GEMM -> exact re-scoring + top-k + softmax + gather/blend (one call)
            call("clc_match_topk_tc", ptr(self.y), ptr(r), B * R, R, M, h, w, p, p, k,
                 1 if self.gaussian_mask else 0, ptr(self.val), ptr(self.idx), ptr(self.n_uncert), self.T, ptr(self.aligned),
                 ptr(self.weights), ptr(self.ws), self.ws.numel(), st)
            n = 1
        else:
            call("clc_pearson_corr", C.byref(self._qview), ptr(r), ptr(self.mask), ptr(self.corr), B * R,
                 self.P, M, p, p, h, w, ptr(self.ws), self.ws.numel(), st)
            call("clc_topk_rows", ptr(self.corr), B * R * self.P, self.corr.shape[-1], k, ptr(self.val),
                 ptr(self.idx), st)
            # 2. softmax weights + gather of the k matched patches + blend
            call("clc_gather_blend_fwd", ptr(r), ptr(self.idx), ptr(self.val), self.T, ptr(self.aligned),
                 ptr(self.weights), B * R, M, h, w, p, p, self.corr_w, k, 0, st)
            n = 3
        # 3. CLM fusion over the aligned references ([B,R,C,S] layout, strided -- no transpose)
        if not (self.match_mode == "tc" and self._fused_fwd):
            call("clc_clm_fuse_fwd", ptr(self.aligned), M * S, R * M * S, ptr(self.att), S, R * S, ptr(self.y),
                 ptr(self.fused), R, B, M, S, st)
            n += 1
        if self.train:
            cur.wait_stream(zs)
            if self._fused_bwd:
                # CLM elementwise backward folded into the match backward: g_aligned is never materialised
                call("clc_match_clm_bwd", C.byref(self._qview), self._r_cl, ptr(self.mask), ptr(self.idx),
                     ptr(self.weights), self.T, ptr(self.g_fused), ptr(self.att), S, R * S, ptr(self.aligned),
                     ptr(self.g_refs), ptr(self.g_q), ptr(self.g_val), ptr(self.g_att), B * R, R, self.P, M, p, p,
                     h, w, k, 3, ptr(self.ws_bwd), self.ws_bwd.numel(), st)
                n += 1
            else:
                call("clc_clm_fuse_bwd", ptr(self.aligned), M * S, R * M * S, ptr(self.att), S, R * S,
                     ptr(self.g_fused), ptr(self.g_aligned), ptr(self.g_att), R, B, M, S, st)
                call("clc_match_bwd", C.byref(self._qview), ptr(r), self._r_cl, ptr(self.mask), ptr(self.idx),
                     ptr(self.weights), self.T, ptr(self.g_aligned), ptr(self.g_refs), ptr(self.g_q), ptr(self.g_val),
                     B * R, self.P, M, p, p, h, w, k, 3, ptr(self.ws_bwd), self.ws_bwd.numel(), st)
                n += 2
        return n

    def hyper_chain(self):
        """EntropyBottleneck on z (+ STE round, bpp partial) and its backward.  Reads z / noise_z."""
        self._zero(self._acc_eb)
        nz = self._noise("noise_z", 0)
        ops.eb_fwd_raw(self.z, nz, self.eb_m, self.eb_b, self.eb_f,
                       self.quantiles, self.lik_z, self.z_hat, None, self.log2[1:2])
        if not self.train:
            return 1
        ops.eb_bwd_raw(self.z, nz, self.eb_m, self.eb_b, self.eb_f, self.quantiles, self.lik_z,
                       None, self.bpp_coef, None, self.g_z, self.g_eb[0:5], self.g_eb[5:10], self.g_eb[10:14])
        return 2      # (data-parallel: the EB parameter gradients are all-reduced in _exchange_stats)

    def slice_chain(self):
        """Slice loop: GaussianConditional + STE round (+ bpp partial), LRP add, and their backward."""
        n = 0
        self._zero(self._acc_y)
        if not self.train:
            noise = [None] * self.num_slices
        elif self.device_noise:
            noise = [self._noise("noise_y", 1 + i) for i in range(self.num_slices)]
        else:
            noise = self._slices(self.noise_y)
        for i, (ys, ss, ms, ls, yh, lr) in enumerate(zip(*(self._slices(t) for t in (
                self.y, self.scale, self.mu, self.lik_y, self.y_hat, self.lrp)))):
            ops.gc_fwd_raw(ys, ss, ms, noise[i], ls, yh, None, self.log2[0:1])
            ops.lrp_add_fwd_raw(yh, lr)
            n += 2
        if not self.train:
            return n
        for i, (ys, ss, ms, ls, lr, gyh, gy, gs, gm, gl) in enumerate(zip(*(self._slices(t) for t in (
                self.y, self.scale, self.mu, self.lik_y, self.lrp, self.g_y_hat, self.g_y, self.g_scale,
                self.g_mu, self.g_lrp)))):
            ops.lrp_add_bwd_raw(gyh, lr, gl)
            ops.gc_bwd_raw(ys, ss, ms, noise[i], ls, None, self.bpp_coef, gyh, gy, gs, gm)
            n += 2
        return n

    def entropy_chain(self):
        return self.hyper_chain() + self.slice_chain()

    def forward(self):
        """Back-compat: whole step on the current stream (forward AND backward when training)."""
        return self.step()

    def step(self, fork=False):
        """Enqueue one full pass (forward, and backward when training).  The three chains
        (match, hyper-latent, slice loop) are data-independent; with fork=True the two entropy
        chains run on a side stream next to the match chain (CUDA graph capture turns this into a
        forked graph).  Returns the number of C-ABI calls."""
        if not fork:
            n = self.match_chain() + self.hyper_chain() + self.slice_chain()
            self._exchange_stats()
            self._finalize()
            return n
        cur = torch.cuda.current_stream(self.device)
        s1, s2 = self._side_streams()
        s1.wait_stream(cur)
        if self.data_parallel:
            # multi-GPU: the step's all-reduce sits at the end of the entropy work and its latency (plus rank
            # skew) must stay under the match chain, so the two entropy chains run on TWO branches
            s2.wait_stream(cur)
            with torch.cuda.stream(s2):
                n = self.hyper_chain()
            with torch.cuda.stream(s1):
                n += self.slice_chain()
                s1.wait_stream(s2)
                self._exchange_stats()
        else:
            with torch.cuda.stream(s1):
                # single GPU: both entropy chains on ONE side branch (measured on B200: two branches next to
                # the match chain are 5% faster than three -- scripts/chain_timing.py)
                n = self.hyper_chain()
                n += self.slice_chain()
        n += self.match_chain()
        cur.wait_stream(s1)
        self._finalize()
        return n

    def _exchange_stats(self):
        """The path's only collective, ONE NCCL all-reduce per step: the 2-double bpp
        statistic and, when training, the EntropyBottleneck parameter gradients, packed into one fp64
        buffer (at 8 ranks a second small all-reduce costs more than the four tiny pack / unpack copies).
        Enqueued at the end of the entropy branch, where it overlaps the match chain."""
        if not self.data_parallel:
            return
        if self._peer is not None:
            # one kernel: pack -> publish over NVLink -> wait -> fixed-order sum; statistic summed, EB parameter
            # gradients averaged (the gradient of the GLOBAL-batch bpp, like every other data-parallel gradient)
            self._peer(self.log2, self._eb_grads_flat if self.train else None)
            return
        buf = self._dp_buf if self.train else self._dp_buf[:2]
        buf[:2].copy_(self.log2)
        if self.train:
            # mean over ranks: the EB gradients were formed with the per-rank 1/(B_local H W) normalisation
            torch.mul(self._eb_grads_flat, 1.0 / self._world, out=buf[2:])
        torch.distributed.all_reduce(buf)
        self.log2.copy_(buf[:2])
        if self.train:
            self._eb_grads_flat.copy_(buf[2:])

    def _finalize(self):
        """bpp of the step from the two accumulated log2 sums (train_CLC.py:48-51) and, with in-kernel noise,
        the advance of the noise stream -- one one-thread kernel at the join of the chains."""
        call("clc_bpp_finalize", ptr(self.log2), 2, float(self.num_pixels * self._world), ptr(self._bpp_dev),
             ptr(self.rng_state) if self.device_noise else None, 16 * self._RNG_STRIDE, ops._stream())

    def _zero_stream(self):
        if self._zs is None:
            self._zs = torch.cuda.Stream(device=self.device)
        return self._zs

    def _side_streams(self):
        if self._streams is None:
            self._streams = tuple(torch.cuda.Stream(device=self.device) for _ in range(3))
        return self._streams[0], self._streams[1]

    # -------------------------------------------------------------------------------------
    def _capture(self, fn):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def capture(self, fork=True):
        """Capture step() into ONE CUDA graph (launch-bound at the small configs); fork=True keeps
        the match chain and the entropy chains on parallel branches of the graph."""
        self._graph = self._capture(lambda: self.step(fork=fork))
        return self._graph

    def replay(self):
        self._graph.replay()

    def capture_split(self):
        """Two graphs for the end-to-end path: match chain / entropy chains, so the second half of
        the host->device upload overlaps the match chain (step_host)."""
        self._g_match = self._capture(self.match_chain)
        self._g_entropy = self._capture(lambda: self._entropy_forked())
        self._host_out = torch.zeros(2, dtype=torch.float64).pin_memory()
        self._ev = [torch.cuda.Event() for _ in range(2)]
        self._side_streams()

    def _entropy_forked(self):
        cur = torch.cuda.current_stream(self.device)
        s1, _ = self._side_streams()
        s1.wait_stream(cur)
        with torch.cuda.stream(s1):
            n = self.hyper_chain()
        n += self.slice_chain()
        cur.wait_stream(s1)
        self._exchange_stats()
        self._finalize()
        return n

    def step_host(self, host_flat):
        """One end-to-end step from HOST buffers: upload this step's inputs from the pinned staging
        buffer (host_staging) in two contiguous copies, run the match chain as soon as its inputs
        have landed while the rest is still in flight, run the entropy chains, read the step's
        result (bpp) back.  Returns bpp as a Python float (synchronises)."""
        if self._g_match is None:
            self.capture_split()
        cur = torch.cuda.current_stream(self.device)
        cp, se = self._streams[2], self._streams[1]
        n1 = self._n_match_in
        cp.wait_stream(cur)
        with torch.cuda.stream(cp):
            self._in[:n1].copy_(host_flat[:n1], non_blocking=True)
            self._ev[0].record(cp)
            self._in[n1:].copy_(host_flat[n1:], non_blocking=True)
            self._ev[1].record(cp)
        cur.wait_event(self._ev[0])
        self._g_match.replay()
        se.wait_event(self._ev[1])
        with torch.cuda.stream(se):
            self._g_entropy.replay()
        cur.wait_stream(se)
        self._host_out.copy_(self.log2, non_blocking=True)
        cur.synchronize()
        return -(self._host_out[0].item() + self._host_out[1].item()) / (self.num_pixels * self._world)

    def n_uncertified(self):
        """Patches of the last step whose top-k could not be certified against the bf16 screening error
        (0 = indices provably equal the exact fp32 ranking).  Synchronises."""
        return int(self.n_uncert.item())

    def bpp(self):
        """Device scalar: -(sum log2 lik_y + sum log2 lik_z) / num_pixels (over all ranks when data-parallel),
        written by clc_bpp_finalize at the end of the step."""
        return self._bpp_dev[0]

    # ---- algorithmic work per kernel launch (stated in DESIGN.md) ------------------
    def algorithmic_work(self):
        """{trace label: ("bytes"|"flops", amount per launch)} for every kernel of one step, SURVEY.md 8(d)
        figures only: COMPULSORY traffic (every tensor the operator must read or write, once).  Re-reads the
        implementation causes (candidate windows re-scored from L2, read-modify-write scatters) are NOT
        algorithmic work; they are listed separately in gathered_bytes() and reported as `l2_bytes`.
        Labels are the ones the library's per-kernel tracing reports (clc_trace_get)."""
        B, R, M, S, k, P = self.B, self.R, self.M, self.h * self.w, self.k, self.P
        NP = B * R
        K = M * self.patch ** 2
        n_slice = B * (M if self.fused_slices else self.Cs) * S
        nz = self.z.numel()
        t = self.train
        L = (self.h - self.patch + 1) * (self.w - self.patch + 1)
        noise_in = 4 if (t and not self.device_noise) else 0
        w = {
            "clc_gc_fwd": ("bytes", n_slice * (20 + noise_in)),            # y,mu,scale(+noise) -> lik,y_hat
            "clc_lrp_add_fwd": ("bytes", n_slice * 12),
            "clc_gc_bwd": ("bytes", n_slice * (16 + noise_in + 12)),       # 8d E1 bwd: 28-32 B/elem
            "clc_lrp_add_bwd": ("bytes", n_slice * 12),
            "clc_eb_fwd": ("bytes", nz * (12 + noise_in)),
            "clc_eb_bwd": ("bytes", nz * (8 + noise_in + 4)),              # 8d: 8 B read + 4 B write
            # K6-K7 gather/blend: k windows + indices/values in, blended reference out, per (image, ref)
            "clc_gather_blend_fwd": ("bytes", NP * ((k + 1) * M * S * 4 + P * k * 8)),
            "clc_clm_fuse_fwd": ("bytes", B * ((R * (M + 1) + M) * S * 4 + M * S * 4)),
            "clc_clm_fuse_bwd": ("bytes", B * ((R * (M + 1) + M) * S * 4 + R * (M + 1) * S * 4)),
            # tensor-core match: 2*P*L*C*ph*pw flop per (image, reference), counted once
            "clc_match_topk_tc(gemm)": ("flops", 2.0 * P * L * K * NP),
            # pre-pass: read the fp32 refs + queries once, write the bf16 GEMM operands (the fp32 channels-last
            # copy it also writes is this implementation's own extra traffic, not counted) + channel sums
            "clc_match_topk_tc(prepass)": ("bytes", (NP + B) * M * S * (4 + 2) + NP * S * 8),
            # re-score + top-k + softmax + gather/blend = 8d K6-K7: (k+1)*C*S*4 + P*k*8 per (image, ref)
            "clc_match_topk_tc(rescore)": ("bytes", NP * ((k + 1) * M * S * 4 + P * k * 8)),
            # candidate selection: one pass over the fp16 screened score map, 2*KC (value, id) pairs out per patch
            "clc_match_topk_tc(select)": ("bytes", NP * P * L * 2 + NP * P * (16 if k <= 4 else 32) * 8),
            "patch_stats": ("bytes", B * M * S * 4 + B * P * 8),
            "clc_pearson_corr": ("flops", 2.0 * P * L * K * NP),
            "clc_topk_rows": ("bytes", NP * P * L * 4 + NP * P * k * 8),
            "channel_sums": ("bytes", NP * M * S * 4 + NP * S * 8),
            "clc_match_bwd(nchw_to_cl)": ("bytes", NP * M * S * 8),
            # fused match backward, compulsory tensors: g_aligned + r in, g_r out per (image, ref); q in, g_q out
            "clc_match_bwd(main)": ("bytes", NP * 3 * M * S * 4 + B * 2 * M * S * 4),
            "clc_match_bwd(cl_to_nchw)": ("bytes", NP * M * S * 8),
            "clc_match_bwd": ("bytes", NP * 3 * M * S * 4 + B * 2 * M * S * 4),
            "clc_bpp_finalize": ("bytes", 64),
            # CLM-fused match backward: g_fused, att, aligned, r, q in; g_r, g_q, g_att out (g_aligned never exists)
            "clc_match_clm_bwd(main)": ("bytes", NP * 3 * M * S * 4 + B * 3 * M * S * 4 + 2 * NP * S * 4),
            "clc_match_clm_bwd(cl_to_nchw)": ("bytes", NP * M * S * 8),
        }
        if self.match_mode == "tc" and self._fused_fwd:
            # re-scoring kernel with the CLM fusion folded in: + query latent and attention logits in, fused out
            kind, v = w["clc_match_topk_tc(rescore)"]
            w["clc_match_topk_tc(rescore)"] = (kind, v + B * 2 * M * S * 4 + NP * S * 4)
        return w

    def gathered_bytes(self):
        """{trace label: bytes per launch} counting every window each time it is gathered (what the kernels
        request from L2), for the kernels whose access pattern re-reads data; bench.py reports it as
        `roofline.l2_bytes` next to the compulsory figure."""
        B, R, M, S, k, P = self.B, self.R, self.M, self.h * self.w, self.k, self.P
        NP = B * R
        K = M * self.patch ** 2
        KC = 8 if k <= 4 else 16
        return {
            "clc_match_topk_tc(rescore)": NP * P * K * 4 * (KC + 1 + k) + NP * P * k * 12 + NP * M * S * 4,
            "clc_match_bwd(main)": NP * P * K * 4 * (2 + k + 2 * k) + B * M * S * 8,
            "clc_match_bwd": NP * P * K * 4 * (2 + k + 2 * k) + B * M * S * 8,
            "clc_match_clm_bwd(main)": NP * P * K * 4 * (3 + k + 2 * k) + B * M * S * 8,
            "clc_gather_blend_fwd": NP * (k * P * K * 4 + P * k * 8 + M * S * 4),
        }

    def algorithmic_bytes(self):
        """Back-compat view: {short name: bytes} + match_flops."""
        out = {}
        for name, (kind, v) in self.algorithmic_work().items():
            if kind == "bytes":
                out[name.replace("clc_", "")] = v
        out["match_flops"] = self.algorithmic_work()["clc_match_topk_tc(gemm)"][1]
        return out


class HostPipeline:
    """Double-buffered end-to-end driver over HOST inputs: two LatentPath instances (own device buffers
    and graphs) alternate, so the upload of step i+1 runs on the copy stream while the kernels and the
    result read-back of step i run on the compute stream.  Every step still uploads its own inputs from
    pinned host memory and reads its own result (bpp) back; at cfg2 the step is PCIe-bound, so this hides
    the kernels and the read-back latency behind the next upload.

        pipe = HostPipeline([LatentPath(...), LatentPath(...)])
        for i, flat in enumerate(batches):            # flat: pinned buffer laid out by host_staging()
            pipe.submit(i, flat)
            if i: bpp = pipe.result(i - 1)            # synchronises on step i-1 only
        bpp = pipe.result(len(batches) - 1)
    """

    def __init__(self, paths):
        assert len(paths) == 2 and paths[0]._in.numel() == paths[1]._in.numel()
        self.lp = list(paths)
        for lp in self.lp:
            if lp._g_match is None:
                lp.capture_split()
        self.copy_stream = torch.cuda.Stream(device=self.lp[0].device)
        self.done = [None, None]

    def submit(self, i, host_flat, before_compute=None):
        """Enqueue step i (no host synchronisation).  `before_compute` (optional callable) is enqueued on
        the compute stream ahead of the step's kernels (bench.py: the L2 flush)."""
        lp = self.lp[i % 2]
        cur = torch.cuda.current_stream(lp.device)
        cp, se = self.copy_stream, lp._streams[1]
        n1 = lp._n_match_in
        if self.done[i % 2] is not None:
            cp.wait_event(self.done[i % 2])          # this slot's previous step has consumed its inputs
        else:
            cp.wait_stream(cur)
        with torch.cuda.stream(cp):
            lp._in[:n1].copy_(host_flat[:n1], non_blocking=True)
            ev0 = cp.record_event()
            lp._in[n1:].copy_(host_flat[n1:], non_blocking=True)
            ev1 = cp.record_event()
        if before_compute is not None:
            before_compute()
        prev = self.done[(i - 1) % 2]
        cur.wait_event(ev0)
        lp._g_match.replay()                         # match chain on the compute stream ...
        if prev is not None:
            se.wait_event(prev)                      # (collectives of consecutive steps never overlap)
        se.wait_event(ev1)
        with torch.cuda.stream(se):
            lp._g_entropy.replay()                   # ... entropy chains next to it, once their inputs landed
        cur.wait_stream(se)
        lp._host_out.copy_(lp.log2, non_blocking=True)
        self.done[i % 2] = cur.record_event()

    def result(self, i):
        """bpp of step i as a Python float (waits for that step only)."""
        lp = self.lp[i % 2]
        self.done[i % 2].synchronize()
        return -(lp._host_out[0].item() + lp._host_out[1].item()) / (lp.num_pixels * lp._world)

