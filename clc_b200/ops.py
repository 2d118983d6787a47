"""torch.autograd wrappers over the C ABI (include/clc_b200.h) -- entropy stage.

PyTorch is plumbing here: it owns device memory, streams and the autograd graph; every
arithmetic step runs in libclc_b200.so.  Nothing in this file computes on the CPU and there is
no fallback path: tensors must be CUDA fp32.
"""
import math

import torch

from . import _lib
from ._lib import call, ptr, ptr_array

SCALE_BOUND = 0.11      # GaussianConditional scale_bound (compressai default; CLC_run.py:726)
LIKELIHOOD_BOUND = 1e-9  # EntropyModel likelihood_bound


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f"clc_b200: `{name}` must be a CUDA float32 tensor "
                        f"(got {type(t).__name__} {getattr(t, 'device', '')} {getattr(t, 'dtype', '')}); "
                        "there is no CPU fallback")


def _rows(t, name):
    """View a [B, ...] tensor as B rows of CS contiguous elements without copying when its
    trailing dims are dense (e.g. a channel-slice of a [B, 320, h, w] tensor).  Returns
    (tensor, batch_stride_in_elements)."""
    _check(t, name)
    if t.dim() < 1:
        raise ValueError(f"{name}: need at least 1 dim")
    shape, stride = t.shape, t.stride()
    dense, acc = True, 1
    for d in range(t.dim() - 1, 0, -1):
        if shape[d] != 1 and stride[d] != acc:
            dense = False
            break
        acc *= shape[d]
    if not dense or (shape[0] > 1 and stride[0] < acc):
        t = t.contiguous()
        return t, acc
    return t, (stride[0] if shape[0] > 1 else acc)


def _bcs(t):
    B = t.shape[0]
    return B, (t.numel() // B if B else 0)


# ---------------------------------------------------------------------------------------------
# raw (no autograd) entry points -- used by the autograd Functions below and by bench.py
# ---------------------------------------------------------------------------------------------
def _is_ticket(noise):
    """In-kernel noise request: (device state tensor int64[2], host offset) from clc_b200.rng.ticket()."""
    return isinstance(noise, tuple)


def gc_fwd_raw(y, scale, mean, noise, lik, y_hat=None, outputs=None, log2_sum=None,
               scale_bound=SCALE_BOUND, lik_bound=LIKELIHOOD_BOUND):
    """`noise`: None (eval), a tensor (explicit sample) or an rng ticket (generated in the kernel)."""
    y, y_bs = _rows(y, "y")
    scale, s_bs = _rows(scale, "scale")
    B, CS = _bcs(y)
    m_bs = n_bs = yh_bs = o_bs = 0
    if mean is not None:
        mean, m_bs = _rows(mean, "mean")
    tk = noise if _is_ticket(noise) else None
    if noise is not None and tk is None:
        noise, n_bs = _rows(noise, "noise")
    lik_, l_bs = _rows(lik, "lik")
    assert lik_ is lik, "lik output must be row-dense"
    if y_hat is not None:
        yh, yh_bs = _rows(y_hat, "y_hat")
        assert yh is y_hat
    if outputs is not None:
        oo, o_bs = _rows(outputs, "outputs")
        assert oo is outputs
    if tk is not None:
        call("clc_gc_fwd_rng", ptr(y), y_bs, ptr(scale), s_bs, ptr(mean), m_bs, ptr(tk[0]), tk[1],
             ptr(lik), l_bs, ptr(y_hat), yh_bs, ptr(outputs), o_bs, ptr(log2_sum), B, CS,
             scale_bound, lik_bound, _stream())
        return
    call("clc_gc_fwd", ptr(y), y_bs, ptr(scale), s_bs, ptr(mean), m_bs, ptr(noise), n_bs,
         ptr(lik), l_bs, ptr(y_hat), yh_bs, ptr(outputs), o_bs, ptr(log2_sum), B, CS,
         scale_bound, lik_bound, _stream())


def gc_bwd_raw(y, scale, mean, noise, lik, g_lik, bpp_coef, g_y_hat, g_y, g_scale, g_mean,
               scale_bound=SCALE_BOUND, lik_bound=LIKELIHOOD_BOUND):
    y, y_bs = _rows(y, "y")
    scale, s_bs = _rows(scale, "scale")
    lik, l_bs = _rows(lik, "lik")
    B, CS = _bcs(y)
    m_bs = n_bs = gl_bs = gy_bs = gm_bs = 0
    if mean is not None:
        mean, m_bs = _rows(mean, "mean")
    tk = noise if _is_ticket(noise) else None
    if noise is not None and tk is None:
        noise, n_bs = _rows(noise, "noise")
    if g_lik is not None:
        g_lik, gl_bs = _rows(g_lik, "g_lik")
    if g_y_hat is not None:
        g_y_hat, gy_bs = _rows(g_y_hat, "g_y_hat")
    _, go_bs = _rows(g_y, "g_y")
    _, gs_bs = _rows(g_scale, "g_scale")
    if g_mean is not None:
        _, gm_bs = _rows(g_mean, "g_mean")
    if tk is not None:
        call("clc_gc_bwd_rng", ptr(y), y_bs, ptr(scale), s_bs, ptr(mean), m_bs, ptr(tk[0]), tk[1],
             ptr(lik), l_bs, ptr(g_lik), gl_bs, float(bpp_coef), ptr(g_y_hat), gy_bs,
             ptr(g_y), go_bs, ptr(g_scale), gs_bs, ptr(g_mean), gm_bs, B, CS, scale_bound, lik_bound,
             _stream())
        return
    call("clc_gc_bwd", ptr(y), y_bs, ptr(scale), s_bs, ptr(mean), m_bs, ptr(noise), n_bs,
         ptr(lik), l_bs, ptr(g_lik), gl_bs, float(bpp_coef), ptr(g_y_hat), gy_bs,
         ptr(g_y), go_bs, ptr(g_scale), gs_bs, ptr(g_mean), gm_bs, B, CS, scale_bound, lik_bound,
         _stream())


def lrp_add_fwd_raw(y_hat, lrp):
    yh, yh_bs = _rows(y_hat, "y_hat")
    assert yh is y_hat, "y_hat must be row-dense for the in-place LRP add"
    lrp, l_bs = _rows(lrp, "lrp")
    B, CS = _bcs(y_hat)
    call("clc_lrp_add_fwd", ptr(y_hat), yh_bs, ptr(lrp), l_bs, B, CS, _stream())


def lrp_add_bwd_raw(g, lrp, g_lrp):
    g, g_bs = _rows(g, "g")
    lrp, l_bs = _rows(lrp, "lrp")
    _, gl_bs = _rows(g_lrp, "g_lrp")
    B, CS = _bcs(lrp)
    call("clc_lrp_add_bwd", ptr(g), g_bs, ptr(lrp), l_bs, ptr(g_lrp), gl_bs, B, CS, _stream())


def eb_param_arrays(matrices, biases, factors):
    return ptr_array(matrices, 5), ptr_array(biases, 5), ptr_array(factors, 4)


def eb_fwd_raw(z, noise, matrices, biases, factors, quantiles, lik, z_hat=None, outputs=None,
               log2_sum=None, lik_bound=LIKELIHOOD_BOUND):
    _check(z, "z")
    assert z.is_contiguous() and z.dim() >= 2
    B, Cc = z.shape[0], z.shape[1]
    S = z.numel() // max(B * Cc, 1)
    pm, pb, pf = eb_param_arrays(matrices, biases, factors)
    if _is_ticket(noise):
        call("clc_eb_fwd_rng", ptr(z), ptr(noise[0]), noise[1], pm, pb, pf, ptr(quantiles), ptr(lik), ptr(z_hat),
             ptr(outputs), ptr(log2_sum), B, Cc, S, lik_bound, _stream())
        return
    call("clc_eb_fwd", ptr(z), ptr(noise), pm, pb, pf, ptr(quantiles), ptr(lik), ptr(z_hat),
         ptr(outputs), ptr(log2_sum), B, Cc, S, lik_bound, _stream())


def eb_bwd_raw(z, noise, matrices, biases, factors, quantiles, lik, g_lik, bpp_coef, g_z_hat, g_z,
               g_matrices=None, g_biases=None, g_factors=None, lik_bound=LIKELIHOOD_BOUND):
    assert z.is_contiguous()
    B, Cc = z.shape[0], z.shape[1]
    S = z.numel() // max(B * Cc, 1)
    pm, pb, pf = eb_param_arrays(matrices, biases, factors)
    if g_matrices is not None:
        gm, gb, gf = eb_param_arrays(g_matrices, g_biases, g_factors)
    else:
        gm = gb = gf = None
    if _is_ticket(noise):
        call("clc_eb_bwd_rng", ptr(z), ptr(noise[0]), noise[1], pm, pb, pf, ptr(quantiles), ptr(lik), ptr(g_lik),
             float(bpp_coef), ptr(g_z_hat), ptr(g_z), gm, gb, gf, B, Cc, S, lik_bound, _stream())
        return
    call("clc_eb_bwd", ptr(z), ptr(noise), pm, pb, pf, ptr(quantiles), ptr(lik), ptr(g_lik),
         float(bpp_coef), ptr(g_z_hat), ptr(g_z), gm, gb, gf, B, Cc, S, lik_bound, _stream())


def log2_sum_fwd_raw(lik, acc):
    _check(lik, "lik")
    lik = lik.contiguous()
    call("clc_log2_sum_fwd", ptr(lik), lik.numel(), ptr(acc), _stream())


# ---------------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------------
class SliceBuffers:
    """Preallocated [B, C, h, w] likelihood / y_hat buffers that the per-slice GaussianConditional launches
    write IN PLACE (channel-slice views, batch-strided kernel arguments), so that the ChARM loop needs no
    `torch.cat` of its outputs (CLC_run.py:587-590).  Deliberately not a tensor: autograd
    must not see the buffers as inputs of the per-slice Functions (their outputs are later modified in place
    by the LRP add)."""

    def __init__(self, B, C, h, w, device):
        self.lik = torch.empty((B, C, h, w), dtype=torch.float32, device=device)
        self.y_hat = torch.empty((B, C, h, w), dtype=torch.float32, device=device)
        self.views = {"lik": [], "y_hat": []}

    @staticmethod
    def _alias(v):
        # same memory as the channel-slice view `v`, but not a view in autograd's eyes (outputs of a custom
        # Function that are views may not be modified in place; y_hat is, by the LRP add)
        return torch.empty(0, dtype=v.dtype, device=v.device).set_(v.untyped_storage(), v.storage_offset(),
                                                                   v.size(), v.stride())

    def take(self, c0, c1):
        return self._alias(self.lik[:, c0:c1]), self._alias(self.y_hat[:, c0:c1])


class _AssembleFn(torch.autograd.Function):
    """The concatenation of slice tensors that already live in one buffer: forward returns the buffer,
    backward hands every slice a channel-slice VIEW of the incoming gradient (no copies either way)."""

    @staticmethod
    def forward(ctx, holder, which, *slices):
        ctx.widths = [t.shape[1] for t in slices]
        return SliceBuffers._alias(getattr(holder, which))

    @staticmethod
    def backward(ctx, g):
        outs, c = [], 0
        for w in ctx.widths:
            outs.append(g[:, c:c + w])
            c += w
        return (None, None, *outs)


def assemble(holder, which, slices):
    return _AssembleFn.apply(holder, which, *slices)


class _GaussianConditionalFn(torch.autograd.Function):
    """(y, scale, mean, noise) -> (lik, y_hat, outputs).  See clc_gc_fwd / clc_gc_bwd."""

    @staticmethod
    def forward(ctx, y, scale, mean, noise, scale_bound, lik_bound, want_outputs, log2_acc, out=None):
        if out is not None:
            lik, y_hat = out                     # channel-slice views of SliceBuffers (written in place)
        else:
            lik = torch.empty(y.shape, dtype=torch.float32, device=y.device)
            y_hat = torch.empty_like(lik)
        outputs = torch.empty_like(lik) if want_outputs else None
        if y.numel():
            gc_fwd_raw(y, scale, mean, noise, lik, y_hat, outputs, log2_acc, scale_bound, lik_bound)
        ctx.ticket = noise if _is_ticket(noise) else None
        ctx.save_for_backward(y, scale, mean, None if ctx.ticket else noise, lik)
        ctx.bounds = (scale_bound, lik_bound)
        ctx.train = noise is not None
        if not want_outputs:
            outputs = y.new_empty(0)
            ctx.mark_non_differentiable(outputs)
        return lik, y_hat, outputs

    @staticmethod
    def backward(ctx, g_lik, g_y_hat, g_outputs):
        y, scale, mean, noise, lik = ctx.saved_tensors
        if ctx.ticket is not None:
            noise = ctx.ticket          # the backward regenerates the forward's sample from the same ticket
        sb, lb = ctx.bounds
        g_y = torch.empty(y.shape, dtype=torch.float32, device=y.device)
        g_scale = torch.empty_like(g_y)
        g_mean = torch.empty_like(g_y) if mean is not None else None
        # g_lik None (likelihood unused downstream) -> NULL pointer + coef 0 = zero gradient.
        if y.numel():
            gc_bwd_raw(y, scale, mean, noise, lik, g_lik, 0.0, g_y_hat, g_y, g_scale, g_mean, sb, lb)
        if g_outputs is not None and g_outputs.numel():
            # `outputs` = y + noise (train: d/dy = 1) or round(y-mean)+mean (eval: d/dmean = 1)
            if ctx.train:
                g_y = g_y + g_outputs
            elif g_mean is not None:
                g_mean = g_mean + g_outputs
        return g_y, g_scale, g_mean, None, None, None, None, None, None


def gaussian_conditional(y, scale, mean=None, noise=None, scale_bound=SCALE_BOUND,
                         lik_bound=LIKELIHOOD_BOUND, want_outputs=False, log2_acc=None, out=None):
    """Fused GaussianConditional.forward + ste_round.  Returns (lik, y_hat, outputs|None).
    `out`: an `_OutViews` pair from SliceBuffers.take() -- the kernel then writes there."""
    lik, y_hat, outputs = _GaussianConditionalFn.apply(y, scale, mean, noise, float(scale_bound),
                                                       float(lik_bound), bool(want_outputs), log2_acc,
                                                       _OutViews(out) if out is not None else None)
    return lik, y_hat, (outputs if want_outputs else None)


class _OutViews(tuple):
    """(lik_view, y_hat_view) wrapped in a non-tensor container so autograd does not treat the views as inputs."""
    def __new__(cls, pair):
        return super().__new__(cls, pair)


class _LrpAddFn(torch.autograd.Function):
    """y_hat += 0.5 * tanh(lrp), in place (CLC_run.py:582-583)."""

    @staticmethod
    def forward(ctx, y_hat, lrp):
        lrp_add_fwd_raw(y_hat, lrp)
        ctx.mark_dirty(y_hat)
        ctx.save_for_backward(lrp)
        return y_hat

    @staticmethod
    def backward(ctx, g):
        (lrp,) = ctx.saved_tensors
        g = g.contiguous()
        g_lrp = torch.empty(lrp.shape, dtype=torch.float32, device=lrp.device)
        lrp_add_bwd_raw(g, lrp, g_lrp)
        return g, g_lrp


def lrp_add_(y_hat, lrp):
    return _LrpAddFn.apply(y_hat, lrp)


def gc_symbols_indexes(y, scale, mean, scale_table, scale_bound=SCALE_BOUND, want_symbols=True,
                       want_indexes=True):
    """int32 symbols = round(y-mean) and scale-table indexes (CLC_run.py:689-690)."""
    ref = y if y is not None else scale
    symbols = indexes = None
    y_bs = s_bs = m_bs = sy_bs = ix_bs = 0
    if want_symbols:
        y, y_bs = _rows(y, "y")
        if mean is not None:
            mean, m_bs = _rows(mean, "mean")
        symbols = torch.empty(y.shape, dtype=torch.int32, device=y.device)
        sy_bs = symbols.numel() // max(symbols.shape[0], 1)
    if want_indexes:
        scale, s_bs = _rows(scale, "scale")
        _check(scale_table, "scale_table")
        scale_table = scale_table.contiguous()
        indexes = torch.empty(scale.shape, dtype=torch.int32, device=scale.device)
        ix_bs = indexes.numel() // max(indexes.shape[0], 1)
    B, CS = _bcs(ref)
    call("clc_gc_symbols_indexes", ptr(y) if want_symbols else None, y_bs,
         ptr(scale) if want_indexes else None, s_bs, ptr(mean) if want_symbols else None, m_bs,
         ptr(scale_table) if want_indexes else None, int(scale_table.numel()) if want_indexes else 0,
         ptr(symbols), sy_bs, ptr(indexes), ix_bs, B, CS, float(scale_bound), _stream())
    return symbols, indexes


class _EntropyBottleneckFn(torch.autograd.Function):
    """(z, noise, 14 params, quantiles) -> (lik, z_hat, outputs).  See clc_eb_fwd / clc_eb_bwd."""

    @staticmethod
    def forward(ctx, z, noise, quantiles, lik_bound, want_outputs, log2_acc, *params):
        assert len(params) == 14
        z = z.contiguous()
        ms, bs, fs = params[0:5], params[5:10], params[10:14]
        for t in (*params, quantiles):
            _check(t, "EntropyBottleneck parameter")
            assert t.is_contiguous()
        lik = torch.empty(z.shape, dtype=torch.float32, device=z.device)
        z_hat = torch.empty_like(lik)
        outputs = torch.empty_like(lik) if want_outputs else None
        eb_fwd_raw(z, noise, ms, bs, fs, quantiles, lik, z_hat, outputs, log2_acc, lik_bound)
        ctx.ticket = noise if _is_ticket(noise) else None
        ctx.save_for_backward(z, None if ctx.ticket else noise, quantiles, lik, *params)
        ctx.lik_bound = lik_bound
        ctx.train = noise is not None
        if not want_outputs:
            outputs = z.new_empty(0)
            ctx.mark_non_differentiable(outputs)
        return lik, z_hat, outputs

    @staticmethod
    def backward(ctx, g_lik, g_z_hat, g_outputs):
        z, noise, quantiles, lik, *params = ctx.saved_tensors
        if ctx.ticket is not None:
            noise = ctx.ticket
        ms, bs, fs = params[0:5], params[5:10], params[10:14]
        g_z = torch.empty(z.shape, dtype=torch.float32, device=z.device)
        need_p = any(ctx.needs_input_grad[6:])
        gp = [torch.zeros_like(p) for p in params] if need_p else None
        if g_lik is not None:
            g_lik = g_lik.contiguous()
        if g_z_hat is not None:
            g_z_hat = g_z_hat.contiguous()
        eb_bwd_raw(z, noise, ms, bs, fs, quantiles, lik, g_lik, 0.0, g_z_hat, g_z,
                   gp[0:5] if need_p else None, gp[5:10] if need_p else None,
                   gp[10:14] if need_p else None, ctx.lik_bound)
        if g_outputs is not None and g_outputs.numel() and ctx.train:
            g_z = g_z + g_outputs
        return (g_z, None, None, None, None, None) + (tuple(gp) if need_p else (None,) * 14)


def entropy_bottleneck(z, noise, matrices, biases, factors, quantiles, lik_bound=LIKELIHOOD_BOUND,
                       want_outputs=False, log2_acc=None):
    """Fused EntropyBottleneck.forward + z STE round.  Returns (lik, z_hat, outputs|None)."""
    lik, z_hat, outputs = _EntropyBottleneckFn.apply(z, noise, quantiles, float(lik_bound),
                                                     bool(want_outputs), log2_acc,
                                                     *matrices, *biases, *factors)
    return lik, z_hat, (outputs if want_outputs else None)


class _Log2SumFn(torch.autograd.Function):
    """sum(log2(lik)) as a float64 scalar; backward g * 1/(lik ln2)."""

    @staticmethod
    def forward(ctx, lik):
        lik = lik.contiguous()
        acc = torch.zeros(1, dtype=torch.float64, device=lik.device)
        log2_sum_fwd_raw(lik, acc)
        ctx.save_for_backward(lik)
        return acc[0]

    @staticmethod
    def backward(ctx, g):
        (lik,) = ctx.saved_tensors
        # d sum(log2 lik) / d lik = g / (lik * ln 2); g stays on the device (no host read).
        g_lik = torch.empty_like(lik)
        g = g.to(torch.float64).contiguous()
        call("clc_log2_sum_bwd", ptr(lik), 1.0 / math.log(2.0), ptr(g), ptr(g_lik), lik.numel(), _stream())
        return g_lik


def log2_sum(lik):
    return _Log2SumFn.apply(lik)
