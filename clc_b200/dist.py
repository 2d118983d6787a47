"""Data-parallel plumbing for the CLC latent path: one process per GPU, images sharded across
ranks, NCCL (NVLink 5 / NVSwitch) used only where the reference's training needs an exchange:
the gradient all-reduce and one small statistics all-reduce per step.  The
reference itself uses threaded nn.DataParallel (train_CLC.py:74-79,:472-473), replicating all
weights every step; this replaces it.  Works with the `gloo` backend on CPU for tests."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank).  Single-process runs return (0, 1, 0) untouched."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return 0, 1, 0
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


class PeerAllReduce:
    """One-shot all-reduce of {n_stat doubles (sum), n_grads floats (mean)} over NVLink peer memory
    (clc_peer_allreduce: one kernel per rank, no ring latency; capturable into CUDA graphs).  torch.distributed
    is the plumbing only: it carries the 64-byte CUDA IPC handles of the per-rank regions once, at
    construction.  All ranks must be on one node with peer access (one NVSwitch domain), world <= 8."""

    def __init__(self, n_stat, n_grads, device, group=None):
        import ctypes as C
        from ._lib import call, lib
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_stat, self.n_grads = int(n_stat), int(n_grads)
        nbytes = lib().clc_peer_allreduce_bytes(self.n_stat, self.n_grads, self.world)
        if nbytes == 0:
            raise RuntimeError("clc_peer_allreduce supports at most 8 ranks")
        with torch.cuda.device(device):
            mine = C.c_void_p()
            call("clc_peer_alloc", nbytes, C.byref(mine))
            handle = (C.c_uint8 * 64)()
            call("clc_peer_export", mine, handle)
            on_gpu = dist.get_backend(group) == "nccl"      # (gloo, used by the tests, gathers host tensors)
            t = torch.tensor(list(handle), dtype=torch.uint8, device=device if on_gpu else "cpu")
            gathered = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(gathered, t, group=group)
            self._mine, self._opened = mine, []
            self._regions = (C.c_void_p * self.world)()
            for r, g in enumerate(gathered):
                if r == self.rank:
                    self._regions[r] = mine.value
                else:
                    h = (C.c_uint8 * 64)(*g.cpu().tolist())
                    p = C.c_void_p()
                    call("clc_peer_open", h, C.byref(p))
                    self._opened.append(p)
                    self._regions[r] = p.value
            self.state = torch.zeros(3, dtype=torch.int64, device=device)
            dist.barrier(group=group)          # every region is open everywhere before the first kernel

    def __call__(self, stat, grads, stream=None):
        """stat: float64[n_stat] (summed in place); grads: float32[n_grads] (replaced by the cross-rank mean)."""
        from ._lib import call, ptr
        assert stat.dtype == torch.float64 and stat.numel() == self.n_stat and stat.is_contiguous()
        assert grads is None or (grads.dtype == torch.float32 and grads.numel() == self.n_grads and grads.is_contiguous())
        st = torch.cuda.current_stream(stat.device).cuda_stream if stream is None else stream
        call("clc_peer_allreduce", self._regions, self.rank, self.world, ptr(stat), self.n_stat,
             ptr(grads), self.n_grads if grads is not None else 0, 1.0 / self.world, ptr(self.state), st)

    def close(self):
        from ._lib import call
        for p in self._opened:
            call("clc_peer_close", p)
        self._opened = []
        if self._mine is not None:
            call("clc_peer_free", self._mine)
            self._mine = None


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, stop) of `n_items` independent units (images) for `rank`."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradAllReducer:
    """Bucketed, overlapped gradient all-reduce (mean) for a replicated model.

    A first (discovery) backward shows which parameters receive a gradient and in which order; they are packed,
    in that order, into flat fp32 buckets of ~`bucket_mb`.  From then on ONE post-accumulate hook per bucket --
    on the parameter whose gradient arrives last -- gathers the bucket's gradients into the flat buffer with a
    single multi-tensor copy and launches its all-reduce asynchronously (NCCL: ReduceOp.AVG, in place) while the
    backward pass continues; `finish()` waits and points every `.grad` at its slice of the reduced bucket (no
    scatter copy).  A training step of this model is launch-bound on the host (~700 parameter tensors), so what
    the reducer must not do is add per-parameter work: a hook per parameter and gradients accumulated INTO bucket
    views (one extra `add_` launch each) cost 9-20 ms per step on 8 B200s for an exchange that takes 0.85 ms.
    Parameters that receive no gradient (the reference's unused `feature_alignment`, `multi_ref_fusion`, `cc_*`,
    `lrp_transforms` modules: 310 tensors) are in no bucket on any rank -- the set is a property of
    the graph, not of the data.  Use the optimiser's `zero_grad(set_to_none=True)` (or `zero_grad()` here)
    between steps."""

    def __init__(self, module, bucket_mb=32.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_elems = int(bucket_mb * 1024 * 1024 // 4)
        self.params = [p for p in module.parameters() if p.requires_grad]
        self._order = []                 # discovery pass: parameters in the order their gradients became ready
        self._buckets = None             # [(flat, [params], [views])]
        self._launched, self._inflight = [], []
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self._handles = [p.register_post_accumulate_grad_hook(self._discover) for p in self.params]

    # ---- discovery step: plain flatten -> all-reduce -> scatter (runs once) ----
    def _discover(self, p):
        self._order.append(p)

    def _finish_discovery(self):
        for h in self._handles:
            h.remove()
        ps = [p for p in self._order if p.grad is not None]
        self._order = []
        plists, cur, cur_n = [], [], 0
        for p in ps:
            cur.append(p)
            cur_n += p.numel()
            if cur_n >= self.bucket_elems:
                plists.append(cur)
                cur, cur_n = [], 0
        if cur:
            plists.append(cur)
        self._buckets, self._handles = [], []
        for bi, plist in enumerate(plists):
            flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=plist[0].device)
            views, off = [], 0
            for p in plist:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self._buckets.append((flat, plist, views))
            # the bucket is complete when the parameter that was ready LAST in the discovery pass is ready
            self._handles.append(plist[-1].register_post_accumulate_grad_hook(
                lambda _p, bi=bi: self._launch(bi, from_hook=True)))
        self._launched = [False] * len(self._buckets)
        for bi in range(len(self._buckets)):     # this step's gradients: reduce them now, without overlap
            self._launch(bi)
        self._wait()

    def _launch(self, bi, from_hook=False):
        if self._launched[bi]:
            return
        flat, plist, views = self._buckets[bi]
        grads = [p.grad for p in plist]
        if any(g is None for g in grads):
            if from_hook:
                return                   # order differs from the discovery pass: finish() picks the bucket up
            flat.zero_()
            pairs = [(v, g) for v, g in zip(views, grads) if g is not None]
            if pairs:
                torch._foreach_copy_([v for v, _ in pairs], [g for _, g in pairs])
        else:
            torch._foreach_copy_(views, grads)
        self._launched[bi] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._inflight.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat))

    def _wait(self):
        for work, flat in self._inflight:
            work.wait()
            if not self._avg:
                flat.div_(self.world)
        self._inflight = []
        for flat, plist, views in self._buckets:
            for p, v in zip(plist, views):
                if p.grad is not None:
                    p.grad = v
        self._launched = [False] * len(self._buckets)

    def finish(self):
        """Call after backward(): completes all buckets; gradients become the cross-rank mean."""
        if self._buckets is None:
            self._finish_discovery()
            return
        for bi in range(len(self._buckets)):
            self._launch(bi)
        self._wait()

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def allreduce_stats(log2_lik_y, log2_lik_z, sq_err, n_pix, device=None, group=None):
    """One 4-element float64 all-reduce(sum) of {sum log2 lik_y, sum log2 lik_z, sum sq err,
    pixels}; returns global (bpp, mse) as python floats on every rank."""
    vals = []
    for v in (log2_lik_y, log2_lik_z, sq_err, n_pix):
        vals.append(v.detach().to(torch.float64).reshape(1) if isinstance(v, torch.Tensor)
                    else torch.tensor([float(v)], dtype=torch.float64))
    dev = device if device is not None else next((v.device for v in vals if v.is_cuda), vals[0].device)
    t = torch.cat([v.to(dev) for v in vals])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    s_y, s_z, s_e, n = t.tolist()
    return -(s_y + s_z) / n, s_e / (3.0 * n)
