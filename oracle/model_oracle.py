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
                log2_acc=None, out=None):
        assert out is None, "oracle mode runs the slice loop with plain tensors (model._inplace_slices = False)"
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
    m._inplace_slices = False
    if getattr(m, "match_refs", False):
        # Level-C wiring: the oracle's SI_Finder (Patch_Matching.py:157-216 restated) per reference, on the
        # CPU in fp32 exactly as the reference would run it; indices are kept for the parity test
        p, k, T = m.match_patch, m.match_k, m.match_temperature
        m.oracle_match_idx = []

        def align(y, feats):
            yc = y.detach().float().cpu()
            fc = feats.detach().float().cpu()
            mask = O.gaussian_masks(fc.shape[-2], fc.shape[-1], p, p)
            outs, idxs = [], []
            for r in range(fc.shape[1]):
                (o,), _, idx = O.si_finder(yc, fc[:, r], p, p, fc[:, r], k, T, mask=mask, return_index=True)
                outs.append(o)
                idxs.append(idx)
            m.oracle_match_idx.append(torch.stack(idxs, 1))          # [B, R, P, k]
            return torch.stack(outs, 1).to(feats.device)

        m._align_refs = align
    return m
