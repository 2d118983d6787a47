"""Backbone building blocks of CLC / TCM (analysis / synthesis / hyper transforms and the
per-slice parameter networks).  BASELINE.json keeps these as plain PyTorch modules -- they are
outside the latent hot path -- but the drop-in models need them with the reference's exact
parameter names so reference / HF checkpoints load:

  GDN, ResidualBlock*, AttentionBlock ....... CompressAI layers used at CLC_run.py:4-11
  WMSA, Block, ConvTransBlock, SWAtten, SwinBlock ... CLC_run.py:108-266 (same in tcm.py:139-308)
  ReferenceEncoder, CLMAlign ................. CLC_run.py:269-313

Host-side differences from the reference (results identical): the relative-position index table
and the shifted-window masks are built once and cached instead of per forward call
(SURVEY.md 8f-4 notes the per-call numpy / 6-D bool tensor rebuilds).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# CompressAI-style layers
# ---------------------------------------------------------------------------------------------
class _Bound(nn.Module):
    """max(x, bound) with the CompressAI gradient gate (pass iff x >= bound or grad < 0)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _BoundFn.apply(x, self.bound)


class _BoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b):
        ctx.save_for_backward(x, b)
        return torch.max(x, b)

    @staticmethod
    def backward(ctx, g):
        x, b = ctx.saved_tensors
        return g * ((x >= b) | (g < 0)), None


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        ped = float(reparam_offset) ** 2
        self.register_buffer("pedestal", torch.Tensor([ped]))
        self.lower_bound = _Bound((float(minimum) + ped) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(in_channels)))

    def forward(self, x):
        C = x.shape[1]
        norm = F.conv2d(x * x, self.gamma_reparam(self.gamma).view(C, C, 1, 1), self.beta_reparam(self.beta))
        return x * (torch.sqrt(norm) if self.inverse else torch.rsqrt(norm))


def conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride)


def subpel_conv3x3(cin, cout, r=1):
    return nn.Sequential(nn.Conv2d(cin, cout * r * r, 3, padding=1), nn.PixelShuffle(r))


def conv(cin, cout, kernel_size=5, stride=2):
    return nn.Conv2d(cin, cout, kernel_size, stride=stride, padding=kernel_size // 2)


class ResidualBlockWithStride(nn.Module):
    def __init__(self, cin, cout, stride=2):
        super().__init__()
        self.conv1 = conv3x3(cin, cout, stride)
        self.conv2 = conv3x3(cout, cout)
        self.gdn = GDN(cout)
        self.skip = conv1x1(cin, cout, stride) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        out = self.gdn(self.conv2(F.leaky_relu(self.conv1(x))))
        return out + (x if self.skip is None else self.skip(x))


class ResidualBlockUpsample(nn.Module):
    def __init__(self, cin, cout, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(cin, cout, upsample)
        self.conv = conv3x3(cout, cout)
        self.igdn = GDN(cout, inverse=True)
        self.upsample = subpel_conv3x3(cin, cout, upsample)

    def forward(self, x):
        return self.igdn(self.conv(F.leaky_relu(self.subpel_conv(x)))) + self.upsample(x)


class ResidualBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = conv3x3(cin, cout)
        self.conv2 = conv3x3(cout, cout)
        self.skip = conv1x1(cin, cout) if cin != cout else None

    def forward(self, x):
        out = F.leaky_relu(self.conv2(F.leaky_relu(self.conv1(x))))
        return out + (x if self.skip is None else self.skip(x))


class _ResidualUnit(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(), conv3x3(N // 2, N // 2), nn.ReLU(),
                                  conv1x1(N // 2, N))

    def forward(self, x):
        return F.relu(self.conv(x) + x)


class AttentionBlock(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv_a = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N))
        self.conv_b = nn.Sequential(_ResidualUnit(N), _ResidualUnit(N), _ResidualUnit(N), conv1x1(N, N))

    def forward(self, x):
        return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x


# ---------------------------------------------------------------------------------------------
# Swin blocks
# ---------------------------------------------------------------------------------------------
class WMSA(nn.Module):
    """(Shifted-)window multi-head self-attention with learned relative position bias."""

    def __init__(self, input_dim, output_dim, head_dim, window_size, type):
        super().__init__()
        self.input_dim, self.output_dim, self.head_dim = input_dim, output_dim, head_dim
        self.scale = head_dim ** -0.5
        self.n_heads = input_dim // head_dim
        self.window_size = window_size
        self.type = type
        self.embedding_layer = nn.Linear(input_dim, 3 * input_dim, bias=True)
        ws = window_size
        table = torch.zeros((2 * ws - 1) * (2 * ws - 1), self.n_heads)
        nn.init.trunc_normal_(table, std=.02)
        # stored as [heads, 2ws-1, 2ws-1], the layout the reference keeps in its state_dict
        self.relative_position_params = nn.Parameter(
            table.view(2 * ws - 1, 2 * ws - 1, self.n_heads).permute(2, 0, 1).contiguous())
        self.linear = nn.Linear(input_dim, output_dim)
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij"), -1).view(-1, 2)
        rel = coords[:, None, :] - coords[None, :, :] + ws - 1
        self.register_buffer("_rel_index", rel[..., 0] * (2 * ws - 1) + rel[..., 1], persistent=False)
        self._mask_cache = {}

    def _bias(self):
        return self.relative_position_params.flatten(1)[:, self._rel_index]  # [heads, p*p, p*p]

    def _shift_mask(self, nh, nw, device):
        key = (nh, nw, str(device))
        m = self._mask_cache.get(key)
        if m is None:
            p, s = self.window_size, self.window_size - self.window_size // 2
            m = torch.zeros(nh, nw, p, p, p, p, dtype=torch.bool, device=device)
            m[-1, :, :s, :, s:, :] = True
            m[-1, :, s:, :, :s, :] = True
            m[:, -1, :, :s, :, s:] = True
            m[:, -1, :, s:, :, :s] = True
            m = m.view(1, 1, nh * nw, p * p, p * p)
            self._mask_cache[key] = m
        return m

    def _fused_ok(self, x):
        """Inference on the GPU: everything between the two Linear layers is one kernel (clc_window_attention_fwd)."""
        return (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and self.window_size == 8
                and self.head_dim in (8, 16, 32) and x.shape[1] % 8 == 0 and x.shape[2] % 8 == 0
                and getattr(self, "fused_attention", True))

    def forward(self, x):  # x: [B, H, W, C]
        p, shifted = self.window_size, self.type != "W"
        if self._fused_ok(x):
            from ._lib import call, ptr
            from .ops import _stream
            B, H, W, C = x.shape
            qkv = self.embedding_layer(x.contiguous())          # per-token: no roll / window partition needed first
            out = torch.empty(B, H, W, C, dtype=torch.float32, device=x.device)
            call("clc_window_attention_fwd", ptr(qkv), ptr(self.relative_position_params.detach().contiguous()),
                 ptr(out), B, H, W, C, self.head_dim, p, 1 if shifted else 0, float(self.scale), _stream())
            return self.linear(out)
        if shifted:
            x = torch.roll(x, shifts=(-(p // 2), -(p // 2)), dims=(1, 2))
        B, H, W, C = x.shape
        nh, nw = H // p, W // p
        x = x.view(B, nh, p, nw, p, C).permute(0, 1, 3, 2, 4, 5).reshape(B, nh * nw, p * p, C)
        qkv = self.embedding_layer(x)
        # channel layout of the 3C projection is (three * heads, head_dim)
        qkv = qkv.view(B, nh * nw, p * p, 3 * self.n_heads, self.head_dim).permute(3, 0, 1, 2, 4)
        q, k, v = qkv[:self.n_heads], qkv[self.n_heads:2 * self.n_heads], qkv[2 * self.n_heads:]
        sim = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        sim = sim + self._bias()[:, None, None]
        if shifted:
            sim = sim.masked_fill(self._shift_mask(nh, nw, x.device), float("-inf"))
        out = torch.matmul(F.softmax(sim, dim=-1), v)  # [heads, B, windows, p*p, head_dim]
        out = out.permute(1, 2, 3, 0, 4).reshape(B, nh * nw, p * p, C)
        out = self.linear(out)
        out = out.view(B, nh, nw, p, p, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)
        if shifted:
            out = torch.roll(out, shifts=(p // 2, p // 2), dims=(1, 2))
        return out


class Block(nn.Module):
    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, type="W", input_resolution=None):
        super().__init__()
        assert type in ("W", "SW")
        if drop_path and drop_path > 0:
            raise NotImplementedError("drop_path > 0 is never used by the reference configs")
        self.type = type
        self.ln1 = nn.LayerNorm(input_dim)
        self.msa = WMSA(input_dim, input_dim, head_dim, window_size, type)
        self.ln2 = nn.LayerNorm(input_dim)
        self.mlp = nn.Sequential(nn.Linear(input_dim, 4 * input_dim), nn.GELU(), nn.Linear(4 * input_dim, output_dim))

    def forward(self, x):
        x = x + self.msa(self.ln1(x))
        return x + self.mlp(self.ln2(x))


class ConvTransBlock(nn.Module):
    """Parallel conv / Swin branches on a channel split, merged by 1x1 convs, residual."""

    def __init__(self, conv_dim, trans_dim, head_dim, window_size, drop_path, type="W"):
        super().__init__()
        self.conv_dim, self.trans_dim = conv_dim, trans_dim
        self.trans_block = Block(trans_dim, trans_dim, head_dim, window_size, drop_path, type)
        self.conv1_1 = nn.Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1)
        self.conv1_2 = nn.Conv2d(conv_dim + trans_dim, conv_dim + trans_dim, 1)
        self.conv_block = ResidualBlock(conv_dim, conv_dim)

    def forward(self, x):
        cx, tx = torch.split(self.conv1_1(x), (self.conv_dim, self.trans_dim), dim=1)
        cx = self.conv_block(cx) + cx
        tx = self.trans_block(tx.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return x + self.conv1_2(torch.cat((cx, tx), dim=1))


class SwinBlock(nn.Module):
    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path):
        super().__init__()
        self.block_1 = Block(input_dim, output_dim, head_dim, window_size, drop_path, "W")
        self.block_2 = Block(input_dim, output_dim, head_dim, window_size, drop_path, "SW")
        self.window_size = window_size

    def forward(self, x):
        ws = self.window_size
        if x.size(-1) <= ws or x.size(-2) <= ws:  # reference pads tiny maps (and never crops back)
            pr, pc = (ws - x.size(-2)) // 2, (ws - x.size(-1)) // 2
            x = F.pad(x, (pc, pc + 1, pr, pr + 1))
        t = self.block_2(self.block_1(x.permute(0, 2, 3, 1)))
        return t.permute(0, 3, 1, 2)


class SWAtten(AttentionBlock):
    def __init__(self, input_dim, output_dim, head_dim, window_size, drop_path, inter_dim=192):
        width = inter_dim if inter_dim is not None else input_dim
        super().__init__(N=width)
        self.non_local_block = SwinBlock(width, width, head_dim, window_size, drop_path)
        self._project = inter_dim is not None
        if self._project:
            self.in_conv = conv1x1(input_dim, inter_dim)
            self.out_conv = conv1x1(inter_dim, output_dim)

    def forward(self, x):
        x = self.in_conv(x)
        z = self.non_local_block(x)
        out = self.conv_a(x) * torch.sigmoid(self.conv_b(z)) + x
        return self.out_conv(out)


class ReferenceEncoder(nn.Module):
    def __init__(self, N=128, M=320):
        super().__init__()
        self.encoder = nn.Sequential(ResidualBlockWithStride(3, N, 2), ResidualBlockWithStride(N, N, 2),
                                     ResidualBlockWithStride(N, M, 2), conv3x3(M, M, stride=2))

    def forward(self, x):
        return self.encoder(x)


class CLMAlign(nn.Module):
    """CLC_run.py:284-313 `CLM` (SWAtten alignment + 1x1 fusion).  Instantiated by the reference as
    `feature_alignment` and never called; kept so the state_dict keys exist."""

    def __init__(self, channels, head_dim=8, window_size=8):
        super().__init__()
        self.channels = channels
        self.alignment = SWAtten(channels * 2, channels, head_dim, window_size, 0, inter_dim=channels)
        self.fusion = nn.Sequential(conv1x1(channels * 2, channels), nn.GELU(), conv1x1(channels, channels))

    def forward(self, x, ref_feat):
        aligned = self.alignment(torch.cat([x, ref_feat], dim=1))
        return self.fusion(torch.cat([x, aligned], dim=1))
