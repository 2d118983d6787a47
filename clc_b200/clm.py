"""CLM conditional fusion: drop-in `SimpleCLM` (models/CLM.py:130-187) whose elementwise core
(sigmoid gate, softmax over references, weighted sum, + y) is one fused sm_100a kernel.
The 1x1 / 3x3 convolutions stay nn.Conv2d, as in the reference; parameter names are identical
so a reference state_dict loads unchanged."""
import torch
import torch.nn as nn

from ._lib import call, ptr
from .ops import _check, _stream


class _ClmFuseFn(torch.autograd.Function):
    """(ref_t [R,B,C,h,w], att [R,B,1,h,w], y [B,C,h,w]) -> sum_r softmax_r(att) sigmoid(att_r) ref_t[r] + y."""

    @staticmethod
    def forward(ctx, ref_t, att, y):
        for t, n in ((ref_t, "ref_t"), (att, "att"), (y, "y")):
            _check(t, n)
        ref_t, att, y = ref_t.contiguous(), att.contiguous(), y.contiguous()
        R, B, Cc = ref_t.shape[0], ref_t.shape[1], ref_t.shape[2]
        S = ref_t[0, 0, 0].numel()
        out = torch.empty_like(y)
        call("clc_clm_fuse_fwd", ptr(ref_t), B * Cc * S, Cc * S, ptr(att), B * S, S, ptr(y), ptr(out), R, B, Cc, S,
             _stream())
        ctx.save_for_backward(ref_t, att)
        return out

    @staticmethod
    def backward(ctx, g_out):
        ref_t, att = ctx.saved_tensors
        R, B, Cc = ref_t.shape[0], ref_t.shape[1], ref_t.shape[2]
        S = ref_t[0, 0, 0].numel()
        g_out = g_out.contiguous()
        g_ref_t = torch.empty_like(ref_t)
        g_att = torch.empty_like(att)
        call("clc_clm_fuse_bwd", ptr(ref_t), B * Cc * S, Cc * S, ptr(att), B * S, S, ptr(g_out), ptr(g_ref_t),
             ptr(g_att), R, B, Cc, S, _stream())
        return g_ref_t, g_att, g_out


def clm_fuse(ref_t, att, y):
    return _ClmFuseFn.apply(ref_t, att, y)


class SimpleCLM(nn.Module):
    """Simplified Conditional Latent Matching module (reference: models/CLM.py:130-187)."""

    def __init__(self, input_dim, temperature=0.5):
        super().__init__()
        self.temperature = temperature  # kept for signature parity; unused by the reference too
        self.feature_transform = nn.Conv2d(input_dim, input_dim, 1)
        self.attention_conv = nn.Conv2d(input_dim, 1, 1)
        self.fusion_conv = nn.Sequential(nn.Conv2d(input_dim, input_dim, 3, padding=1), nn.ReLU(inplace=True))

    def forward(self, y, y_refs):
        """y [B,C,H,W]; y_refs: list of R tensors [B,C,H,W] -> fused [B,C,H,W]."""
        B, Cc, H, W = y.shape
        R = len(y_refs)
        # (the reference also computes feature_transform(y) and discards it, CLM.py:161)
        # One conv launch over all references: [R*B, C, H, W] is already the [R,B,C,H,W] layout
        # the fused kernel wants, so the reference's torch.stack copies disappear.
        ref_t = self.feature_transform(torch.cat(list(y_refs), dim=0))
        att = self.attention_conv(ref_t)
        fused = clm_fuse(ref_t.view(R, B, Cc, H, W), att.view(R, B, 1, H, W), y)
        return self.fusion_conv(fused)


# ----------------------------------------------------------------------------------------------
# CLM variant (a): similarity-softmax alignment (models/CLM.py:5-128)
# ----------------------------------------------------------------------------------------------
class _ForwardOnly(torch.autograd.Function):
    """Marks a kernel result as non-differentiable LOUDLY: the reference's variant (a) is not trainable in
    practice (Python loops over B*H*W*9 taps) and this port implements its forward pass only."""

    @staticmethod
    def forward(ctx, out, *inputs):
        return out.clone()

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("clc_b200.CLM (variant a, models/CLM.py:62-128) is forward-only; "
                                  "train with SimpleCLM or the match -> clm_fuse path")


def _fwd_only(out, *inputs):
    if torch.is_grad_enabled() and any(t.requires_grad for t in inputs):
        return _ForwardOnly.apply(out, *inputs)
    return out


def clm_sim_colsum(y_t, ref_t, temperature):
    """Column sums of softmax(y_t^T ref_t / T, dim=-1) (CLM.py:104-107 as used by :16-20), never materialising the
    HW x HW map.  y_t [B, C, H, W]; ref_t [R*B, C, H, W] ([R, B] stacking) -> [R*B, H*W]."""
    _check(y_t, "y_t")
    _check(ref_t, "ref_t")
    y_t, ref_t = y_t.detach().contiguous(), ref_t.detach().contiguous()
    NB, Cc = ref_t.shape[0], ref_t.shape[1]
    HW = ref_t.shape[2] * ref_t.shape[3]
    out = torch.empty(NB, HW, dtype=torch.float32, device=ref_t.device)
    from ._lib import lib
    ws = torch.empty(max(1, lib().clc_clm_sim_colsum_workspace_bytes(NB, HW)), dtype=torch.uint8, device=ref_t.device)
    call("clc_clm_sim_colsum", ptr(y_t), ptr(ref_t), NB, y_t.shape[0], Cc, HW, float(temperature), ptr(out), ptr(ws),
         ws.numel(), _stream())
    return out


def clm_weighted_concat(x, colsum):
    """torch.cat([x, x * colsum], 1) (CLM.py:16-22) in one pass.  x [NB, C, H, W], colsum [NB, H*W]."""
    _check(x, "x")
    x = x.detach().contiguous()
    NB, Cc, H, W = x.shape
    out = torch.empty(NB, 2 * Cc, H, W, dtype=torch.float32, device=x.device)
    call("clc_clm_weighted_concat", ptr(x), ptr(colsum), ptr(out), NB, Cc, H * W, _stream())
    return out


def clm_deform_sample(x, offset, modulation, modulation_is_logit=True):
    """DeformableAlignment.deform_conv (CLM.py:35-60).  x [NB, C, H, W], offset [NB, 18, H, W] (conv output),
    modulation [NB, 9, H, W] (conv output when modulation_is_logit: the sigmoid of :25 is fused)."""
    for t, n in ((x, "x"), (offset, "offset"), (modulation, "modulation")):
        _check(t, n)
    x, offset, modulation = x.detach().contiguous(), offset.detach().contiguous(), modulation.detach().contiguous()
    NB, Cc, H, W = x.shape
    out = torch.empty_like(x)
    call("clc_clm_deform_fwd", ptr(x), ptr(offset), ptr(modulation), 1 if modulation_is_logit else 0, ptr(out), NB, Cc,
         H, W, _stream())
    return out


def clm_attention_sum(aligned, att, y):
    """sum_r softmax_r(att) * aligned[r] + y (CLM.py:117-126).  aligned [R, B, C, H, W], att [R, B, 1, H, W]."""
    for t, n in ((aligned, "aligned"), (att, "att"), (y, "y")):
        _check(t, n)
    aligned, att, y = aligned.detach().contiguous(), att.detach().contiguous(), y.detach().contiguous()
    R, B, Cc = aligned.shape[:3]
    S = aligned[0, 0, 0].numel()
    out = torch.empty_like(y)
    call("clc_clm_attention_sum_fwd", ptr(aligned), ptr(att), ptr(y), ptr(out), R, B, Cc, S, _stream())
    return out


class DeformableAlignment(nn.Module):
    """models/CLM.py:5-60 (same parameter names).  `forward` takes the similarity map's column sums instead of the
    HW x HW map itself -- the only part of it the reference's accumulation loop (:16-20) uses."""

    def __init__(self, input_dim):
        super().__init__()
        self.offset_conv = nn.Conv2d(input_dim * 2, 2 * 3 * 3, kernel_size=3, padding=1)
        self.modulation_conv = nn.Conv2d(input_dim * 2, 3 * 3, kernel_size=3, padding=1)

    def forward(self, x, colsum):
        cat = clm_weighted_concat(x, colsum)
        return clm_deform_sample(x, self.offset_conv(cat), self.modulation_conv(cat), modulation_is_logit=True)


class CLM(nn.Module):
    """Conditional Latent Matching module, variant (a) (reference: models/CLM.py:62-128), forward pass.
    Parameter names are the reference's, so its state_dict loads unchanged.  All R references go through every
    convolution / kernel as one [R*B] batch."""

    def __init__(self, input_dim, temperature=0.5):
        super().__init__()
        self.temperature = temperature
        self.feature_transform = nn.Sequential(nn.Conv2d(input_dim, input_dim, 1), nn.ReLU(inplace=True),
                                               nn.Conv2d(input_dim, input_dim, 1))
        self.alignment = DeformableAlignment(input_dim)
        self.attention_conv = nn.Conv2d(input_dim, 1, 1)
        self.fusion_conv = nn.Sequential(nn.Conv2d(input_dim, input_dim, 3, padding=1), nn.ReLU(inplace=True),
                                         nn.Conv2d(input_dim, input_dim, 3, padding=1))

    def forward(self, y, y_refs, return_parts=False):
        """y [B,C,H,W]; y_refs: list of R tensors [B,C,H,W] -> fused [B,C,H,W]."""
        B, Cc, H, W = y.shape
        R = len(y_refs)
        refs = torch.cat(list(y_refs), dim=0)                       # [R*B, C, H, W]
        y_t = self.feature_transform(y)
        colsum = clm_sim_colsum(y_t, self.feature_transform(refs), self.temperature)
        aligned = self.alignment(refs, colsum)
        att = self.attention_conv(aligned)
        fused_in = clm_attention_sum(aligned.view(R, B, Cc, H, W), att.view(R, B, 1, H, W), y)
        fused = self.fusion_conv(_fwd_only(fused_in, y, refs, *self.parameters()))
        if return_parts:
            return fused, colsum.view(R, B, H * W), aligned.view(R, B, Cc, H, W)
        return fused
