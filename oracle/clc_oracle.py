"""TEST INFRASTRUCTURE ONLY -- standalone CPU restatement of the CLC latent hot path.

Plain PyTorch on CPU (fp32 to follow the reference's arithmetic, fp64 variants where a test
needs to adjudicate fp32 noise).  Each function cites the reference lines it restates.  The
restatements are pinned against the reference's own code by tests/test_oracle_golden.py using
fixtures produced by oracle/make_golden.py (which executes the reference source itself).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import enable_shim

enable_shim()
from compressai.entropy_models import EntropyBottleneck, GaussianConditional  # noqa: E402  (shim)
from compressai.ops import LowerBound  # noqa: E402,F401


# ----------------------------------------------------------------------------------------------
# entropy stage
# ----------------------------------------------------------------------------------------------
def get_scale_table(min=0.11, max=256, levels=64):
    """CLC_run.py:32-33."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def ste_round(x):
    """CLC_run.py:35-36."""
    return torch.round(x) - x.detach() + x


def gaussian_likelihood(inputs, scales, means=None, dtype=None):
    """CLC_run.py:718-736 (`_likelihood` + `_standardized_cumulative`), without the 1e-9 floor."""
    if dtype is not None:
        inputs, scales = inputs.to(dtype), scales.to(dtype)
        means = means.to(dtype) if means is not None else None
    half = float(0.5)
    values = inputs - means if means is not None else inputs
    scales = torch.max(scales, torch.tensor(0.11, dtype=scales.dtype))
    values = torch.abs(values)
    const = float(-(2 ** -0.5))
    upper = half * torch.erfc(const * ((half - values) / scales))
    lower = half * torch.erfc(const * ((-half - values) / scales))
    return upper - lower


class _NoiseInjected:
    """Context that makes EntropyModel.quantize('noise') use a caller-supplied noise tensor, so the
    training-mode path can be compared bit-for-bit."""

    def __init__(self, module, noise):
        self.module, self.noise = module, noise

    def __enter__(self):
        mod, noise = self.module, self.noise
        orig = mod.quantize

        def quantize(inputs, mode, means=None):
            if mode == "noise":
                return inputs + noise.reshape(inputs.shape)
            return orig(inputs, mode, means)

        mod.quantize = quantize
        return mod

    def __exit__(self, *a):
        del self.module.quantize


def gc_forward(y, scale, mean=None, noise=None):
    """GaussianConditional.forward as called at CLC_run.py:569 + ste_round :571.
    noise None -> eval ("dequantize").  Returns (outputs, likelihood, y_hat)."""
    gc = GaussianConditional(None)
    gc.train(noise is not None)
    if noise is not None:
        with _NoiseInjected(gc, noise):
            outputs, lik = gc(y, scale, mean)
    else:
        outputs, lik = gc(y, scale, mean)
    y_hat = ste_round(y - mean) + mean if mean is not None else ste_round(y)
    return outputs, lik, y_hat


def lrp_add(y_hat, lrp):
    """CLC_run.py:582-583."""
    return y_hat + 0.5 * torch.tanh(lrp)


def gc_symbols_indexes(y, scale, mean, scale_table):
    """CLC_run.py:689-690."""
    gc = GaussianConditional(None)
    gc.update_scale_table(scale_table)
    return gc.quantize(y, "symbols", mean), gc.build_indexes(scale)


def eb_forward(eb, z, noise=None):
    """EntropyBottleneck.forward as called at CLC_run.py:526 + z STE :528-530, on a shim
    EntropyBottleneck `eb` holding the parameters.  Returns (outputs, likelihood, z_hat)."""
    eb.train(noise is not None)
    if noise is not None:
        # the module permutes to channel-major [C,1,B*S] before quantising
        perm = noise.transpose(0, 1).reshape(noise.shape[1], 1, -1)
        with _NoiseInjected(eb, perm):
            outputs, lik = eb(z)
    else:
        outputs, lik = eb(z)
    med = eb._get_medians().reshape(1, -1, *([1] * (z.dim() - 2)))
    z_hat = ste_round(z - med) + med
    return outputs, lik, z_hat


def bpp_loss(likelihoods, num_pixels):
    """train_CLC.py:48-51 / eval.py:27-31."""
    return sum((torch.log(l).sum() / (-math.log(2) * num_pixels)) for l in likelihoods)


def rate_distortion_loss(output, target, lmbda):
    """train_CLC.py:43-54 (type='mse')."""
    N, _, H, W = target.size()
    out = {"bpp_loss": bpp_loss(output["likelihoods"].values(), N * H * W)}
    out["mse_loss"] = F.mse_loss(output["x_hat"], target)
    out["loss"] = lmbda * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
    return out


def psnr(a, b):
    """eval.py:20-22."""
    return -10 * math.log10(torch.mean((a - b) ** 2).item())


# ----------------------------------------------------------------------------------------------
# reference matching (models/Patch_Matching.py), `.cuda()` removed
# ----------------------------------------------------------------------------------------------
def pearson_corr(x, y, dtype=torch.float32):
    """L2_or_pearson_corr, Patch_Matching.py:854-910.  x [P,C,ph,pw] patches, y [1,C,fh,fw]
    -> [1,P,fh-ph+1,fw-pw+1].  The conv2d weights are detached query patches (:869)."""
    x, y = x.to(dtype), y.to(dtype)
    P, C, H, W = x.shape
    patch_size = int(H * W * C)
    xy = F.conv2d(y, x.detach())
    y_mean = F.conv2d(y, torch.ones(1, C, H, W, dtype=dtype) / patch_size)
    x_sum = torch.sum(x, dim=[1, 2, 3])
    numerator = xy - y_mean * x_sum[None, :, None, None]
    sum_x_square = torch.sum(torch.square(x), dim=[1, 2, 3])
    x_mean = torch.mean(x, dim=[1, 2, 3])
    denominator_x = sum_x_square - x_mean * x_sum
    sum_y_square = F.conv2d(torch.square(y), torch.ones(1, C, H, W, dtype=dtype))
    denominator_y = sum_y_square - y_mean * y_mean * patch_size
    denominator = denominator_y * denominator_x[None, :, None, None]
    return numerator / torch.sqrt(denominator)


def gaussian_masks(img_h, img_w, patch_h, patch_w):
    """create_gaussian_masks, Patch_Matching.py:779-807 (numpy float64 -> float32)."""
    num_patches = np.arange(0, (img_h * img_w) // (patch_h * patch_w))
    patch_img_w = img_w / patch_w
    w = np.arange(1, img_w + 1, 1, float) - (patch_w % 2) / 2
    h = (np.arange(1, img_h + 1, 1, float) - (patch_h % 2) / 2)[:, np.newaxis]
    center_h = (num_patches // patch_img_w + 0.5) * patch_h
    center_w = ((num_patches % patch_img_w) + 0.5) * patch_w
    sigma_h, sigma_w = 0.5 * img_h, 0.5 * img_w
    cols = (w - center_w[:, np.newaxis])[:, np.newaxis, :] ** 2 / sigma_w ** 2
    rows = np.transpose(h - center_h)[:, :, np.newaxis] ** 2 / sigma_h ** 2
    g = np.exp(-4 * np.log(2) * (rows + cols))
    g = g[:, (patch_h + 1) // 2 - 1:img_h - patch_h // 2, (patch_w + 1) // 2 - 1:img_w - patch_w // 2]
    return torch.from_numpy(g.astype(np.float32)[np.newaxis])


def topk_lowest_index(corr2d, k):
    """torch.topk with a deterministic tie rule (lowest index first) -- torch's own tie order is
    implementation-defined."""
    order = torch.sort(corr2d, dim=1, descending=True, stable=True)
    return order.values[:, :k], order.indices[:, :k]


def si_wrapper(cross_corr, patch_h, patch_w, patches_num, y, k=1, temperature=15, is_stack=False,
               return_index=False):
    """SI_Wraper, Patch_Matching.py:218-240."""
    _, _, corr_h, corr_w = cross_corr.shape
    _, C, fh, fw = y.shape
    cc = cross_corr.reshape(1, -1, corr_h * corr_w)
    value, index = topk_lowest_index(cc[0], k)
    weight = F.softmax(value * temperature, dim=1)
    index_h, index_w = torch.div(index, corr_w, rounding_mode="floor"), index % corr_w
    ph_i, pw_i = torch.meshgrid(torch.arange(0, patch_h), torch.arange(0, patch_w), indexing="ij")
    ih = index_h[:, :, None, None] + ph_i
    iw = index_w[:, :, None, None] + pw_i
    pixel_index = (ih * fw + iw).reshape(-1)
    y_patches = torch.index_select(y.reshape(-1, C, fh * fw), 2, pixel_index).reshape(
        -1, C, patches_num, k, patch_h, patch_w)
    if is_stack:
        out = y_patches.reshape(-1, C, fh // patch_h, fw // patch_w, k, patch_h, patch_w).permute(
            0, 4, 1, 2, 5, 3, 6).reshape(-1, k * C, fh, fw)
    else:
        y_patches = torch.sum(y_patches * weight[None, None, :, :, None, None], 3)
        out = y_patches.reshape(-1, C, fh // patch_h, fw // patch_w, patch_h, patch_w).permute(
            0, 1, 2, 4, 3, 5).reshape(-1, C, fh, fw)
    if return_index:
        return out, value, index, weight
    return out


def extract_patches(x, patch_h, patch_w):
    """Patch_Matching.py:172: [1,C,H,W] -> [P,C,ph,pw], patch = py*(W/pw)+px."""
    _, C, H, W = x.shape
    return x.reshape(1, C, H // patch_h, patch_h, W // patch_w, patch_w).permute(0, 2, 4, 1, 3, 5).reshape(
        -1, C, patch_h, patch_w)


def si_finder(x_decs, ys, patch_h, patch_w, y_decs, k, temperature, is_stack=False, mask=None,
              other_ys=None, return_index=False, dtype=torch.float32):
    """SI_Finder_at_Decoder_Feature_Domain, Patch_Matching.py:157-216 (single_layer=0 branch,
    feature-domain matching).  Returns a list of per-scale outputs [N,C,fh_i,fw_i] (+ indices)."""
    N = x_decs.shape[0]
    outs, vals, idxs = [], [], []
    for n in range(N):
        q = extract_patches(x_decs[n:n + 1], patch_h, patch_w)
        P = q.shape[0]
        cross = pearson_corr(q, y_decs[n:n + 1], dtype=dtype).to(torch.float32)
        if mask is not None:
            cross = cross * mask
        o, v, i, _ = si_wrapper(cross, patch_h, patch_w, P, ys[n:n + 1], k, temperature, is_stack, True)
        per = [o]
        if other_ys is not None:
            for s_i, oy in enumerate(other_ys):
                s = 2 ** (s_i + 1)
                per.append(si_wrapper(cross[:, :, ::s, ::s], patch_h // s, patch_w // s, P, oy[n:n + 1], k,
                                      temperature, is_stack))
        outs.append(per)
        vals.append(v)
        idxs.append(i)
    res = [torch.cat([o[j] for o in outs], 0) for j in range(len(outs[0]))]
    if return_index:
        return res, torch.stack(vals), torch.stack(idxs)
    return res


# ----------------------------------------------------------------------------------------------
# CLM fusion (models/CLM.py)
# ----------------------------------------------------------------------------------------------
def clm_fuse(ref_t, att, y):
    """Elementwise core of SimpleCLM.forward, CLM.py:170-182.
    ref_t [R,B,C,H,W], att [R,B,1,H,W], y [B,C,H,W]."""
    feats = [ref_t[r] * torch.sigmoid(att[r]) for r in range(ref_t.shape[0])]
    w = F.softmax(torch.stack([att[r] for r in range(att.shape[0])], dim=1), dim=1)
    stack = torch.stack(feats, dim=1)
    return (stack * w).sum(dim=1) + y


class SimpleCLM(torch.nn.Module):
    """SimpleCLM, CLM.py:130-187 (same parameter names)."""

    def __init__(self, input_dim, temperature=0.5):
        super().__init__()
        self.temperature = temperature
        self.feature_transform = torch.nn.Conv2d(input_dim, input_dim, 1)
        self.attention_conv = torch.nn.Conv2d(input_dim, 1, 1)
        self.fusion_conv = torch.nn.Sequential(torch.nn.Conv2d(input_dim, input_dim, 3, padding=1),
                                               torch.nn.ReLU(inplace=True))

    def forward(self, y, y_refs):
        ref_t = torch.stack([self.feature_transform(r) for r in y_refs], 0)
        att = torch.stack([self.attention_conv(t) for t in ref_t], 0)
        return self.fusion_conv(clm_fuse(ref_t, att, y))


# ----------------------------------------------------------------------------------------------
# CLM variant (a): similarity-softmax alignment (models/CLM.py:5-128)
# ----------------------------------------------------------------------------------------------
def clm_sim_colsum(y_t, ref_t, temperature):
    """Column sums of the row-softmax of the HW x HW similarity (CLM.py:104-107 feeding :16-20).

    DeformableAlignment.forward accumulates `weighted_x += sim[:, i, j, :, :] * x` over every query position
    (i, j) (:17-20), i.e. weighted_x[b, c, p] = x[b, c, p] * sum_q softmax_p(sim[b, q, :])[p]: only the column
    sums of the similarity map are ever used.  y_t, ref_t [B, C, H, W] -> [B, H*W]."""
    B, C = y_t.shape[:2]
    sim = torch.bmm(y_t.reshape(B, C, -1).transpose(1, 2), ref_t.reshape(B, C, -1)) / temperature
    return F.softmax(sim, dim=-1).sum(dim=1)


def clm_deform_sample(x, offset, modulation):
    """DeformableAlignment.deform_conv (CLM.py:35-60), vectorised; same arithmetic order per element.
    x [B, C, H, W]; offset [B, 9, 2, H, W]; modulation [B, 9, 1, H, W] (already sigmoided) -> [B, C, H, W].
    Quirks kept: no kernel-tap base offsets (every tap samples around (h, w) itself), taps outside
    [0, H-1] x [0, W-1] contribute nothing, int() truncation == floor on the valid range, (h1, w1) clamped."""
    B, C, H, W = x.shape
    hh = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    ww = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    xf = x.reshape(B, C, H * W)
    result = torch.zeros_like(x)
    for k in range(9):
        off_h = hh + offset[:, k, 0]
        off_w = ww + offset[:, k, 1]
        valid = (off_h >= 0) & (off_h <= H - 1) & (off_w >= 0) & (off_w <= W - 1)
        h0 = off_h.nan_to_num(0.0).clamp(0, H - 1).long()       # (invalid / NaN taps are masked out below)
        w0 = off_w.nan_to_num(0.0).clamp(0, W - 1).long()
        h1 = (h0 + 1).clamp_max(H - 1)
        w1 = (w0 + 1).clamp_max(W - 1)
        lh = (off_h - h0).unsqueeze(1)
        lw = (off_w - w0).unsqueeze(1)

        def at(hi, wi):
            return torch.gather(xf, 2, (hi * W + wi).view(B, 1, H * W).expand(B, C, H * W)).view(B, C, H, W)

        val = (1 - lh) * (1 - lw) * at(h0, w0) + lh * (1 - lw) * at(h1, w0) + (1 - lh) * lw * at(h0, w1) \
            + lh * lw * at(h1, w1)
        result = result + torch.where(valid.unsqueeze(1), val * modulation[:, k], torch.zeros_like(val))
    return result


def clm_attention_sum(aligned, att, y):
    """CLM.py:117-126: softmax over the references of the 1-channel attention, weighted sum, + y.
    aligned [R, B, C, H, W], att [R, B, 1, H, W]."""
    w = F.softmax(att.permute(1, 0, 2, 3, 4), dim=1)
    return (aligned.permute(1, 0, 2, 3, 4) * w).sum(dim=1) + y


class DeformableAlignment(torch.nn.Module):
    """CLM.py:5-33 (same parameter names)."""

    def __init__(self, input_dim):
        super().__init__()
        self.offset_conv = torch.nn.Conv2d(input_dim * 2, 2 * 3 * 3, kernel_size=3, padding=1)
        self.modulation_conv = torch.nn.Conv2d(input_dim * 2, 3 * 3, kernel_size=3, padding=1)

    def forward(self, x, colsum):
        B, C, H, W = x.shape
        weighted_x = x * colsum.view(B, 1, H, W)
        cat = torch.cat([x, weighted_x], dim=1)
        offset = self.offset_conv(cat).view(B, 9, 2, H, W)
        modulation = torch.sigmoid(self.modulation_conv(cat)).view(B, 9, 1, H, W)
        return clm_deform_sample(x, offset, modulation)


class CLM(torch.nn.Module):
    """CLM, CLM.py:62-128 (same parameter names)."""

    def __init__(self, input_dim, temperature=0.5):
        super().__init__()
        self.temperature = temperature
        self.feature_transform = torch.nn.Sequential(torch.nn.Conv2d(input_dim, input_dim, 1),
                                                     torch.nn.ReLU(inplace=True),
                                                     torch.nn.Conv2d(input_dim, input_dim, 1))
        self.alignment = DeformableAlignment(input_dim)
        self.attention_conv = torch.nn.Conv2d(input_dim, 1, 1)
        self.fusion_conv = torch.nn.Sequential(torch.nn.Conv2d(input_dim, input_dim, 3, padding=1),
                                               torch.nn.ReLU(inplace=True),
                                               torch.nn.Conv2d(input_dim, input_dim, 3, padding=1))

    def forward(self, y, y_refs, return_parts=False):
        y_t = self.feature_transform(y)
        colsums, aligned, att = [], [], []
        for y_ref in y_refs:
            cs = clm_sim_colsum(y_t, self.feature_transform(y_ref), self.temperature)
            al = self.alignment(y_ref, cs)
            colsums.append(cs)
            aligned.append(al)
            att.append(self.attention_conv(al))
        fused = self.fusion_conv(clm_attention_sum(torch.stack(aligned), torch.stack(att), y))
        if return_parts:
            return fused, torch.stack(colsums), torch.stack(aligned)
        return fused
