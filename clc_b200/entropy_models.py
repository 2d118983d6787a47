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

from . import ops, rng


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

    # ---- range-coder side (EntropyModel._pmf_to_cdf / compress / decompress [upstream], reached from
    # CLC_run.py:486-491, :643-644, :749).  Tables are built once per checkpoint on the host (they
    # are saved in the state_dict and must be identical on the encoding and the decoding machine, so
    # they do not depend on the device); the quantiser and the coder are C entry points. ----
    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        from .ans import pmf_to_quantized_cdf
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = torch.IntTensor(pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision))
            cdf[i, : _cdf.size(0)] = _cdf
        return cdf

    def _check_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def coder_tables(self):
        """Host copy of (quantized_cdf, cdf_length, offset) in the coder's layout, cached until the
        buffers change (the reference converts the whole table with `.tolist()` on every call)."""
        from .ans import _Tables
        self._check_tables()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, tuple(self._quantized_cdf.shape))
        if getattr(self, "_tables_key", None) != key:
            self._tables = _Tables(self._quantized_cdf, self._cdf_length.reshape(-1), self._offset.reshape(-1))
            self._tables_key = key
        return self._tables

    def compress(self, inputs, indexes, means=None):
        """One rANS string per batch element (EntropyModel.compress [upstream])."""
        from .ans import RansEncoder
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        tables = self.coder_tables()
        symbols = self.quantize(inputs, "symbols", means)
        sym = symbols.reshape(symbols.size(0), -1).to(torch.int32).cpu()          # ONE device->host copy
        idx = indexes.reshape(indexes.size(0), -1).to(torch.int32).cpu()
        coder = RansEncoder()
        return [coder.encode_with_indexes(sym[i], idx[i], tables, None, None) for i in range(sym.size(0))]

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        from .ans import RansDecoder
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        if means is not None and means.size()[:2] != indexes.size()[:2]:
            raise ValueError("Invalid means or indexes parameters")
        tables = self.coder_tables()
        idx = indexes.reshape(indexes.size(0), -1).to(torch.int32).cpu()
        coder = RansDecoder()
        out = torch.stack([coder.decode_with_indexes(s, idx[i], tables, None, None, as_tensor=True)
                           for i, s in enumerate(strings)])
        outputs = out.reshape(indexes.size()).to(indexes.device)                  # ONE host->device copy
        return self.dequantize(outputs, means, dtype)


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
        self.update()
        return True

    @staticmethod
    def _standardized_cumulative(inputs: Tensor) -> Tensor:
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile: float) -> float:
        # inverse normal CDF in float64 (upstream: scipy.stats.norm.ppf)
        return float(torch.special.ndtri(torch.tensor(quantile, dtype=torch.float64)))

    def update(self):
        """Quantised CDF table per scale-table entry (GaussianConditional.update [upstream]): symmetric
        integer support of half-width ceil(scale * |ppf(tail_mass / 2)|), bin masses from the same
        erfc form as the likelihood."""
        dev = self.scale_table.device
        table = self.scale_table.detach().float().cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = table.unsqueeze(1)
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)

    def forward(self, inputs, scales, means=None, training=None, *, noise=None, ste=False,
                want_outputs=True, log2_acc=None, out=None):
        if training is None:
            training = self.training
        if training and noise is None:
            noise = rng.ticket(inputs.device)        # U(-1/2, 1/2) generated inside the kernel (Philox)
        if not training:
            noise = None
        lik, y_hat, outputs = ops.gaussian_conditional(
            inputs, scales, means, noise, self._scale_bound, self._lik_bound, want_outputs=want_outputs,
            log2_acc=log2_acc, out=out)
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

    def _logits_cumulative(self, inputs: Tensor) -> Tensor:
        """Cumulative logits with detached parameters (torch; used by the aux loss and by update())."""
        logits = inputs
        for i in range(5):
            logits = torch.matmul(torch.nn.functional.softplus(getattr(self, f"_matrix{i}").detach()), logits)
            logits = logits + getattr(self, f"_bias{i}").detach()
            if i < 4:
                logits = logits + torch.tanh(getattr(self, f"_factor{i}").detach()) * torch.tanh(logits)
        return logits

    def loss(self) -> Tensor:
        """Auxiliary quantile loss (3 values per channel; stays in torch)."""
        return torch.abs(self._logits_cumulative(self.quantiles) - self.target).sum()

    def forward(self, x, training=None, *, noise=None, ste=False, want_outputs=True, log2_acc=None):
        if training is None:
            training = self.training
        if training and noise is None:
            noise = rng.ticket(x.device)             # U(-1/2, 1/2) generated inside the kernel (Philox)
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
        """Quantised CDF table per channel (EntropyBottleneck.update [upstream]): integer support from
        the learned quantiles, bin masses from the factorised density (sign-stabilised form)."""
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.quantiles.device
        with torch.no_grad():
            host = {n: t.detach().float().cpu() for n, t in self.named_parameters()}
            q = host["quantiles"]
            medians = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]

            def logits(x):
                for i in range(5):
                    x = torch.matmul(torch.nn.functional.softplus(host[f"_matrix{i}"]), x) + host[f"_bias{i}"]
                    if i < 4:
                        x = x + torch.tanh(host[f"_factor{i}"]) * torch.tanh(x)
                return x

            lower, upper = logits(samples - 0.5), logits(samples + 0.5)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
            self._offset = (-minima).to(dev)
            self._cdf_length = (pmf_length + 2).to(dev)
        return True

    def _build_indexes(self, size):
        N, C = size[0], size[1]
        view = [1] * len(size)
        view[1] = -1
        return torch.arange(C, dtype=torch.int32).view(*view).repeat(N, 1, *size[2:])

    def _medians_like(self, n, spatial_dims):
        med = self._get_medians().detach().reshape(-1, *([1] * spatial_dims))
        return med.expand(n, *([-1] * (spatial_dims + 1)))

    def compress(self, x):
        """One string per image (EntropyBottleneck.compress [upstream], CLC_run.py:643)."""
        from .ans import RansEncoder
        tables = self.coder_tables()
        sym = torch.round(x - self._medians_like(x.size(0), x.dim() - 2)).to(torch.int32)
        sym = sym.reshape(x.size(0), -1).cpu()                                      # one device->host copy
        idx = self._build_indexes(x.size()).reshape(x.size(0), -1)
        coder = RansEncoder()
        return [coder.encode_with_indexes(sym[i], idx[i], tables, None, None) for i in range(sym.size(0))]

    def decompress(self, strings, size):
        """-> z_hat [len(strings), C, *size] on the module's device (CLC_run.py:644, :749)."""
        from .ans import RansDecoder
        tables = self.coder_tables()
        out_size = (len(strings), self._quantized_cdf.size(0), *size)
        idx = self._build_indexes(out_size).reshape(len(strings), -1)
        coder = RansDecoder()
        sym = torch.stack([coder.decode_with_indexes(s, idx[i], tables, None, None, as_tensor=True)
                           for i, s in enumerate(strings)]).reshape(out_size)
        med = self._medians_like(len(strings), len(size))
        return self.dequantize(sym.to(med.device), med, med.dtype)
