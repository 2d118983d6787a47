"""TEST INFRASTRUCTURE ONLY -- CPU oracle port of clc_b200.latent_path.LatentPath.step():
the same operator sequence written with the oracle restatements (plain PyTorch autograd on
CPU).  Used by tests/test_latent_path_gpu.py as the checker and by bench.py as the timed
`cpu_baseline` / `--impl reference` arm (kind "port": the reference itself cannot travel to
the GPU box, and its three path components are not wired together upstream)."""
import math

import torch

from . import clc_oracle as O


def make_eb(state):
    """Shim EntropyBottleneck holding the given parameter tensors."""
    eb = O.EntropyBottleneck(state["quantiles"].shape[0])
    with torch.no_grad():
        for i in range(5):
            getattr(eb, f"_matrix{i}").copy_(state[f"_matrix{i}"])
            getattr(eb, f"_bias{i}").copy_(state[f"_bias{i}"])
            if i < 4:
                getattr(eb, f"_factor{i}").copy_(state[f"_factor{i}"])
        eb.quantiles.copy_(state["quantiles"])
    return eb


def step(inp, eb, train=True, patch=4, k=4, temperature=15.0, gaussian_mask=True, num_slices=5,
         backward=True):
    """inp: dict of CPU tensors named as LatentPath.INPUT_NAMES.  Returns a dict of outputs and
    (when training) gradients."""
    y = inp["y"].clone().requires_grad_(train)
    z = inp["z"].clone().requires_grad_(train)
    refs = inp["refs"].clone().requires_grad_(train)
    mu = inp["mu"].clone().requires_grad_(train)
    scale = inp["scale"].clone().requires_grad_(train)
    lrp = inp["lrp"].clone().requires_grad_(train)
    att = inp["att"].clone().requires_grad_(train)
    B, R, M, h, w = refs.shape
    mask = O.gaussian_masks(h, w, patch, patch) if gaussian_mask else None
    aligned, vals, idxs = [], [], []
    for r in range(R):
        out, v, i = O.si_finder(y, refs[:, r], patch, patch, refs[:, r], k, temperature, mask=mask,
                                return_index=True)
        aligned.append(out[0])
        vals.append(v)
        idxs.append(i)
    aligned = torch.stack(aligned, 0)                       # [R,B,M,h,w]
    fused = O.clm_fuse(aligned, att.transpose(0, 1), y)
    _, lik_z, z_hat = O.eb_forward(eb, z, noise=inp["noise_z"] if train else None)
    liks, y_hats = [], []
    Cs = M // num_slices
    for s in range(num_slices):
        sl = slice(s * Cs, (s + 1) * Cs)
        _, lik, y_hat = O.gc_forward(y[:, sl], scale[:, sl], mu[:, sl], noise=inp["noise_y"][:, sl] if train else None)
        y_hats.append(O.lrp_add(y_hat, lrp[:, sl]))
        liks.append(lik)
    lik_y, y_hat = torch.cat(liks, 1), torch.cat(y_hats, 1)
    num_pixels = B * h * 16 * w * 16
    bpp = O.bpp_loss([lik_y, lik_z], num_pixels)
    res = {"val": torch.stack(vals, 1), "idx": torch.stack(idxs, 1), "aligned": aligned.transpose(0, 1),
           "fused": fused, "lik_z": lik_z, "z_hat": z_hat, "lik_y": lik_y, "y_hat": y_hat, "bpp": bpp}
    if train and backward:
        loss = bpp + (y_hat * inp["g_y_hat"]).sum() + (fused * inp["g_fused"]).sum()
        loss.backward()
        res.update(g_y=y.grad, g_z=z.grad, g_refs=refs.grad, g_mu=mu.grad, g_scale=scale.grad,
                   g_lrp=lrp.grad, g_att=att.grad,
                   g_eb={n: p.grad for n, p in eb.named_parameters() if p.grad is not None})
    return {k_: (v.detach() if isinstance(v, torch.Tensor) else v) for k_, v in res.items()}
