"""Drop-in `GaussianConditional` / `EntropyBottleneck` (CompressAI-style operator API, as used
at models/CLC_run.py:483-491,:526-528,:569,:689-690) backed by the fused sm_100a kernels.

Same constructor arguments, parameter / buffer names (so reference and HF checkpoints load),
method names and error behaviour (`ValueError` for a bad quantisation mode or scale table).
CUDA fp32 only: there is no CPU path.
"""
import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops


class LowerBound(nn.Module):
    """Holds the `bound` buffer for state_dict compatibility (the gate itself is fused into the
    kernels: forward max(x, bound), backward passes iff x >= bound or grad < 0)."""
    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._value = float(bound)

    def value(self) -> float:
        return self._value


def _uniform_noise_like(x):
    return torch.empty_like(x).uniform_(-0.5, 0.5)


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder=None, entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self._lik_bound = float(likelihood_bound) if self.use_likelihood_bound else 0.0
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    def quantize(self, inputs: Tensor, mode: str, means=None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return _NoiseAddFn.apply(inputs, _uniform_noise_like(inputs))
        if mode == "symbols":
            symbols, _ = ops.gc_symbols_indexes(inputs, None, means, None, want_indexes=False)
            return symbols
        return _DequantizeFn.apply(inputs, means)

    @staticmethod
    def dequantize(inputs: Tensor, means=None, dtype=torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs


class _NoiseAddFn(torch.autograd.Function):
    """inputs + noise through the GC kernel's `outputs` path is overkill; the add is plumbing."""

    @staticmethod
    def forward(ctx, x, noise):
        return x + noise

    @staticmethod
    def backward(ctx, g):
        return g, None


class _DequantizeFn(torch.autograd.Function):
    """round(x - means) + means via the symbols kernel (eval-mode quantise, zero gradient to x)."""

    @staticmethod
    def forward(ctx, x, means):
        sym, _ = ops.gc_symbols_indexes(x, None, means, None, want_indexes=False)
        out = sym.to(torch.float32)
        if means is not None:
            out += means
        ctx.has_means = means is not None
        return out

    @staticmethod
    def backward(ctx, g):
        return torch.zeros_like(g), (g if ctx.has_means else None)


class GaussianConditional(EntropyModel):
    """Mean-scale Gaussian conditional.  forward(inputs, scales, means=None, training=None)
    -> (outputs, likelihood).  Extra keyword `noise=` injects the U(-1/2,1/2) sample explicitly
    (bit-reproducible parity runs); `ste=True` additionally returns round(inputs-means)+means
    with straight-through gradient from the same kernel launch."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table",
                             torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self._scale_bound = float(scale_bound)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = torch.as_tensor(scale_table, dtype=torch.float32).to(self.scale_table.device).clone()
        return True

    def forward(self, inputs, scales, means=None, training=None, *, noise=None, ste=False,
                want_outputs=True, log2_acc=None):
        if training is None:
            training = self.training
        if training and noise is None:
            noise = _uniform_noise_like(inputs)
        if not training:
            noise = None
        lik, y_hat, outputs = ops.gaussian_conditional(
            inputs, scales, means, noise, self._scale_bound, self._lik_bound, want_outputs=want_outputs,
            log2_acc=log2_acc)
        if ste:
            return outputs, lik, y_hat
        return outputs, lik

    def build_indexes(self, scales: Tensor) -> Tensor:
        if self.scale_table.numel() == 0:
            raise ValueError("scale_table is empty: call update_scale_table / model.update() first")
        _, idx = ops.gc_symbols_indexes(None, scales, None, self.scale_table.to(scales.device),
                                        self._scale_bound, want_symbols=False)
        return idx


class EntropyBottleneck(EntropyModel):
    """Factorised prior with filters (3,3,3,3).  forward(x, training=None) -> (outputs, likelihood)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise ValueError("clc_b200 implements the reference configuration filters=(3,3,3,3) only")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        Cc = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(Cc, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(Cc, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(Cc, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(Cc, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2].detach()

    def _params(self):
        ms = [getattr(self, f"_matrix{i}") for i in range(5)]
        bs = [getattr(self, f"_bias{i}") for i in range(5)]
        fs = [getattr(self, f"_factor{i}") for i in range(4)]
        return ms, bs, fs

    def loss(self) -> Tensor:
        """Auxiliary quantile loss (3 values per channel; stays in torch)."""
        logits = self.quantiles
        for i in range(5):
            logits = torch.matmul(torch.nn.functional.softplus(getattr(self, f"_matrix{i}").detach()), logits)
            logits = logits + getattr(self, f"_bias{i}").detach()
            if i < 4:
                logits = logits + torch.tanh(getattr(self, f"_factor{i}").detach()) * torch.tanh(logits)
        return torch.abs(logits - self.target).sum()

    def forward(self, x, training=None, *, noise=None, ste=False, want_outputs=True, log2_acc=None):
        if training is None:
            training = self.training
        if training and noise is None:
            noise = _uniform_noise_like(x)
        if not training:
            noise = None
        ms, bs, fs = self._params()
        lik, z_hat, outputs = ops.entropy_bottleneck(x, noise, ms, bs, fs, self.quantiles.detach(),
                                                     self._lik_bound, want_outputs=want_outputs,
                                                     log2_acc=log2_acc)
        if ste:
            return outputs, lik, z_hat
        return outputs, lik

    def update(self, force=False):
        return False  # CDF tables belong to the rANS path (out of scope this round)
