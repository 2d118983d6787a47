"""Noise stream of the in-kernel quantisation noise (clc_gc_fwd_rng / clc_eb_fwd_rng).

The reference draws `U(-1/2, 1/2)` with torch's global generator inside compressai's
`EntropyModel.quantize(mode="noise")` (reached from CLC_run.py:526 and :569).  Here the kernels generate the
sample themselves (Philox4x32-10) from a per-device state {seed, base offset} held in DEVICE memory plus a
per-call offset chosen on the host:

  * eager use (the drop-in modules): every call takes a fresh offset from a host counter, so no extra
    launch and no noise tensor; `manual_seed()` makes runs reproducible.
  * CUDA graphs: host offsets are frozen into the graph, so the owner of the graph advances the DEVICE base
    offset once per replay (`advance()` inside the captured region, or clc_bpp_finalize's rng arguments).
"""
import torch

from ._lib import call, ptr

_STATE = {}      # device index -> int64[2] tensor {seed, base offset}
_CALLS = {}      # device index -> host-side call counter
_STRIDE = 1 << 32   # offsets of successive calls are this far apart; element counters never reach it (2^34 elems)


def state(device):
    """The device-resident {seed, base offset} pair (int64 view of the uint64 pair)."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _STATE:
        _STATE[idx] = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64,
                                   device=torch.device("cuda", idx))
        _CALLS[idx] = 0
    return _STATE[idx]


def manual_seed(seed, device=None):
    """Reseed (and rewind) the noise stream of `device` (default: current CUDA device)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = state(dev)
    st.copy_(torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))
    _CALLS[st.device.index] = 0


def ticket(device):
    """(state tensor, host offset) for one noise-consuming call; consecutive tickets never overlap."""
    st = state(device)
    i = st.device.index
    _CALLS[i] += 1
    return st, _CALLS[i] * _STRIDE


def advance(device, n=1 << 20):
    """Advance the DEVICE base offset (one-thread kernel, capturable): call once per replay of a captured
    region that contains noise-consuming kernels."""
    st = state(device)
    call("clc_rng_advance", ptr(st), int(n) * _STRIDE, torch.cuda.current_stream(st.device).cuda_stream)
