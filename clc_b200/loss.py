"""Drop-in `RateDistortionLoss` (train_CLC.py:36-59) and `compute_bpp` / `compute_psnr`
(eval.py:20-31).  The rate term sum(log(lik)) / (-ln2 * N*H*W) runs in the fused log2-sum kernel
(float32 per-thread partials, one float64 atomic per CTA); the distortion term stays torch
(not part of the latent path)."""
import math

import torch
import torch.nn as nn

from . import ops


class RateDistortionLoss(nn.Module):
    def __init__(self, lmbda=1e-2, type="mse"):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = lmbda
        self.type = type

    def forward(self, output, target):
        N, _, H, W = target.size()
        out = {}
        num_pixels = N * H * W
        # sum_k log(lik_k).sum() / (-ln2 * num_pixels) == -sum_k log2(lik_k).sum() / num_pixels
        total = None
        for lik in output["likelihoods"].values():
            s = ops.log2_sum(lik)
            total = s if total is None else total + s
        out["bpp_loss"] = (total / (-float(num_pixels))).to(torch.float32)
        if self.type == "mse":
            out["mse_loss"] = self.mse(output["x_hat"], target)
            out["loss"] = self.lmbda * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
        else:
            raise NotImplementedError("ms-ssim distortion needs pytorch_msssim, which is outside the latent path")
        return out


def compute_bpp(out_net):
    size = out_net["x_hat"].size()
    num_pixels = size[0] * size[2] * size[3]
    total = sum(ops.log2_sum(l) for l in out_net["likelihoods"].values())
    return (total / (-float(num_pixels))).item()


def compute_psnr(a, b):
    mse = torch.mean((a - b) ** 2).item()
    return -10 * math.log10(mse)
