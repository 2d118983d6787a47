import torch.nn as nn
from torch.nn.init import trunc_normal_  # same algorithm as timm's


class DropPath(nn.Module):
    """Stochastic depth; the reference only ever builds it with p=0 -> nn.Identity (CLC_run.py:180)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep
