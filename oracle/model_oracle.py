"""TEST INFRASTRUCTURE ONLY -- run a drop-in clc_b200 model with the ORACLE's entropy arithmetic.

`to_oracle_mode(model)` swaps the model's fused-kernel entropy modules for adapters around the
pure-PyTorch shim classes (same state) and the LRP add for the oracle's, so the very same
backbone weights can be evaluated on CPU (or with eager torch ops on a GPU) as the model-level
oracle for the CUDA path.  Lives under oracle/ and is imported by tests only."""
import copy

import torch
import torch.nn as nn

from . import clc_oracle as O


class _OracleGC(nn.Module):
    def __init__(self, src):
        super().__init__()
        self.inner = O.GaussianConditional(None)
        self.inner.load_state_dict(src.state_dict(), strict=False)

    def forward(self, inputs, scales, means=None, training=None, *, noise=None, ste=False, want_outputs=True,
                log2_acc=None):
        training = self.training if training is None else training
        self.inner.train(training)
        if training and noise is not None:
            with O._NoiseInjected(self.inner, noise):
                outputs, lik = self.inner(inputs, scales, means)
        else:
            outputs, lik = self.inner(inputs, scales, means)
        if ste:
            y_hat = O.ste_round(inputs - means) + means if means is not None else O.ste_round(inputs)
            return outputs, lik, y_hat
        return outputs, lik

    def update_scale_table(self, *a, **k):
        return self.inner.update_scale_table(*a, **k)


class _OracleEB(nn.Module):
    def __init__(self, src):
        super().__init__()
        self.inner = O.EntropyBottleneck(src.channels)
        self.inner.load_state_dict(src.state_dict(), strict=False)

    def forward(self, x, training=None, *, noise=None, ste=False, want_outputs=True, log2_acc=None):
        training = self.training if training is None else training
        outputs, lik, z_hat = O.eb_forward(self.inner, x, noise=noise if training else None) \
            if (noise is not None or not training) else (*self._sampled(x), None)
        if z_hat is None:
            med = self.inner._get_medians().reshape(1, -1, 1, 1)
            z_hat = O.ste_round(x - med) + med
        if ste:
            return outputs, lik, z_hat
        return outputs, lik

    def _sampled(self, x):
        self.inner.train(True)
        return self.inner(x)

    def loss(self):
        return self.inner.loss()

    def _get_medians(self):
        return self.inner._get_medians()


def to_oracle_mode(model):
    """Deep-copies `model` and replaces its hot-path ops with the oracle's.  Returns the copy."""
    m = copy.deepcopy(model)
    dev = next(model.parameters()).device
    m.gaussian_conditional = _OracleGC(model.gaussian_conditional).to(dev)
    m.entropy_bottleneck = _OracleEB(model.entropy_bottleneck).to(dev)
    m._lrp_add = lambda y_hat, lrp: O.lrp_add(y_hat, lrp)
    if getattr(m, "match_refs", False):
        raise NotImplementedError("oracle mode covers the shipped forward (match_refs=False)")
    return m
