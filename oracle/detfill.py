"""TEST INFRASTRUCTURE ONLY -- deterministic, construction-order-independent weight fill.

Random init depends on module construction order and RNG consumption, so it cannot be shared
between the reference model (built in this container) and the drop-in model (built on the GPU
box).  `fill_(model, seed)` instead assigns every floating-point parameter from a CPU generator
seeded by crc32(parameter name), so any two models with the same state_dict keys get identical
weights.  GDN beta/gamma, EntropyBottleneck `_matrix*`/quantiles and all buffers keep their
(RNG-free) default initialisation.  The scale-transform output biases are spread over
[0.05, 3] so the Gaussian scales are not all below the 0.11 floor."""
import math
import zlib

import torch

_SKIP_SUFFIX = ("beta", "gamma", "quantiles")


def fill_(model, seed=0):
    with torch.no_grad():
        for name, p in model.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            if leaf in _SKIP_SUFFIX or leaf.startswith("_matrix"):
                continue
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
            if leaf.startswith("_bias"):
                v = torch.rand(p.shape, generator=g) - 0.5
            elif leaf.startswith("_factor"):
                v = 0.2 * (torch.rand(p.shape, generator=g) - 0.5)
            elif leaf == "relative_position_params":
                v = 0.02 * torch.randn(p.shape, generator=g)
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                a = 1.0 / math.sqrt(fan_in)
                v = (2 * torch.rand(p.shape, generator=g) - 1) * a
            elif leaf == "weight":  # LayerNorm weight
                v = 1.0 + 0.1 * (torch.rand(p.shape, generator=g) - 0.5)
            else:  # biases
                v = 0.1 * (torch.rand(p.shape, generator=g) - 0.5)
            if "cc_scale_transforms" in name and name.endswith(".4.bias"):
                v = torch.exp(torch.linspace(math.log(0.05), math.log(3.0), p.numel()))
            p.copy_(v.to(p.dtype))
    return model


def det_image(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)
