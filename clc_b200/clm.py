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
