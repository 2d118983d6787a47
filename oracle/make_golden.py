"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every array written here is produced by reference source (imported unmodified through the
CompressAI/timm shim, or exec'd from source slices with `.cuda()` stripped -- see ref_loader.py).
The committed fixtures let the oracle restatements and the CUDA path be checked against the
reference on the GPU box, where /root/reference does not exist.
"""
import math
import os
import types

import numpy as np
import torch

from . import detfill, ref_loader

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", name, {k: tuple(np.asarray(v.detach() if isinstance(v, torch.Tensor) else v).shape)
                          for k, v in arrs.items()})


def golden_gaussian(CLC):
    """CLC._likelihood (CLC_run.py:718-736), ste_round (:35-36), get_scale_table (:32-33)."""
    import models.CLC_run as R
    self = types.SimpleNamespace()
    self._standardized_cumulative = lambda x: CLC._standardized_cumulative(self, x)
    lik = lambda i, s, m: CLC._likelihood(self, i, s, m)
    # known-answer vectors of SURVEY.md 8c
    y = torch.tensor([0, 0.3, 1, -2, 5, 12, 0.3, 40.0])
    mu = torch.tensor([0, 0.1, 0, 0.5, 0, 0, 0.1, 0.0])
    sc = torch.tensor([1, 0.5, 0.05, 2, 0.11, 1, 64, 1.0])
    y_hat = R.ste_round(y - mu) + mu
    kat = lik(y_hat, sc, mu)
    tail_v = torch.arange(0, 14, dtype=torch.float32)
    tail = lik(tail_v, torch.ones(14), torch.zeros(14))
    # broad random sweep: scales log-uniform over [0.05, 300], exact-half residues, outliers
    g = torch.Generator().manual_seed(1234)
    n = 4096
    ry = 3 * torch.randn(n, generator=g)
    rmu = torch.randn(n, generator=g)
    rsc = torch.exp(torch.empty(n).uniform_(math.log(0.05), math.log(300.0), generator=g))
    ry[:40] = rmu[:40] + torch.arange(-20, 20, dtype=torch.float32) + 0.5  # exact halves
    ry[40:43] = rmu[40:43] + torch.tensor([8.0, 12.0, 40.0])
    noise = torch.rand(n, generator=g) - 0.5
    r_yhat = R.ste_round(ry - rmu) + rmu
    lik_eval = lik(r_yhat, rsc, rmu)
    lik_train = lik(ry + noise, rsc, rmu)
    lik_eval64 = lik(r_yhat.double(), rsc.double(), rmu.double())
    rounds = torch.round(torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 2.5000002]))
    _save("gaussian.npz", kat_y=y, kat_mu=mu, kat_scale=sc, kat_y_hat=y_hat, kat_lik=kat, tail_v=tail_v,
          tail_lik=tail, y=ry, mu=rmu, scale=rsc, noise=noise, y_hat=r_yhat, lik_eval=lik_eval,
          lik_train=lik_train, lik_eval64=lik_eval64, scale_table=R.get_scale_table(), rounds=rounds)


def golden_match():
    """L2_or_pearson_corr / create_gaussian_masks / SI_Wraper / SI_Finder (Patch_Matching.py)."""
    pm = ref_loader.load_patch_matching()
    g = torch.Generator().manual_seed(7)
    N, C, h, w, p, k = 2, 40, 12, 16, 4, 4
    y = torch.randn(N, C, h, w, generator=g)
    r = 0.5 * y + torch.randn(N, C, h, w, generator=g)
    r_half = torch.randn(N, C, h // 2, w // 2, generator=g)
    mask = pm.create_gaussian_masks(h, w, p, p)
    q0 = y[:1].reshape(1, C, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5).reshape(-1, C, p, p)
    corr0 = pm.L2_or_pearson_corr(q0, r[:1], p, p)
    P = q0.shape[0]
    wr_k1 = pm.SI_Wraper(corr0, p, p, P, r[:1], 1, 15, False)
    wr_stack = pm.SI_Wraper(corr0 * mask, p, p, P, r[:1], k, 15, True)
    args = types.SimpleNamespace(num_k=k, temperature=15, is_stack=False, single_layer=0)
    f_mask = pm.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1"], args, mask=mask)["1"]
    f_nomask = pm.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1"], args)["1"]
    f_multi = pm.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1", "2"], args, mask=mask, other_ys=[r_half])
    val, idx = torch.topk((corr0 * mask).reshape(P, -1), k, dim=1)
    mask_odd = pm.create_gaussian_masks(9, 6, 3, 3)
    _save("match.npz", y=y, r=r, r_half=r_half, mask=mask, corr0=corr0, wr_k1=wr_k1, wr_stack=wr_stack,
          f_mask=f_mask, f_nomask=f_nomask, f_multi_1=f_multi["1"], f_multi_2=f_multi["2"], topk_val=val,
          topk_idx=idx, mask_odd=mask_odd, geom=np.array([N, C, h, w, p, k]))


def golden_clm():
    """SimpleCLM / CLM forward (models/CLM.py), seed 42 as in its __main__."""
    clm = ref_loader.load_clm()
    torch.manual_seed(42)
    B, C, H, W, M = 2, 16, 8, 12, 3
    m = clm.SimpleCLM(C)
    y = torch.randn(B, C, H, W)
    refs = [torch.randn(B, C, H, W) for _ in range(M)]
    out = m(y, refs)
    sd = {"sd_" + k.replace(".", "__"): v for k, v in m.state_dict().items()}
    _save("clm.npz", y=y, refs=torch.stack(refs), out=out, **sd)


def golden_clm_full():
    """CLM variant (a) forward (models/CLM.py:62-128, DeformableAlignment :5-60), seed 42 as in its __main__.
    The offset convolution is scaled up (weights are data) so that the sampled taps leave their cell, cross
    integer boundaries and fall outside the image; intermediates are taken with forward hooks."""
    clm = ref_loader.load_clm()
    torch.manual_seed(42)
    B, C, H, W, M = 2, 16, 8, 12, 3
    m = clm.CLM(C, temperature=0.5)
    with torch.no_grad():
        m.alignment.offset_conv.weight.mul_(8.0)
        m.alignment.offset_conv.bias.add_(0.7)
    y = torch.randn(B, C, H, W)
    refs = [torch.randn(B, C, H, W) for _ in range(M)]
    colsums, aligned = [], []

    def tap(mod, inp, out):
        colsums.append(inp[1].sum(dim=1))
        aligned.append(out)

    hook = m.alignment.register_forward_hook(tap)
    with torch.no_grad():
        out = m(y, refs)
    hook.remove()
    sd = {"sd_" + k.replace(".", "__"): v for k, v in m.state_dict().items()}
    _save("clm_full.npz", y=y, refs=torch.stack(refs), out=out, colsum=torch.stack(colsums),
          aligned=torch.stack(aligned), **sd)


def golden_rd_loss():
    RD = ref_loader.load_rd_loss()
    g = torch.Generator().manual_seed(3)
    lik_y = torch.rand(2, 32, 4, 4, generator=g).clamp_min(1e-9)
    lik_z = torch.rand(2, 8, 1, 1, generator=g).clamp_min(1e-9)
    x_hat = torch.rand(2, 3, 64, 64, generator=g)
    x = torch.rand(2, 3, 64, 64, generator=g)
    out = RD(lmbda=0.013)({"likelihoods": {"y": lik_y, "z": lik_z}, "x_hat": x_hat}, x)
    _save("rd_loss.npz", lik_y=lik_y, lik_z=lik_z, x_hat=x_hat, x=x, bpp_loss=out["bpp_loss"],
          mse_loss=out["mse_loss"], loss=out["loss"])


def golden_model(CLC, TCM):
    """Full reference CLC.forward / TCM.forward (cfg1: N=64, 1x3x256x256, 3 refs, eval), weights
    from detfill.fill_ (name-seeded, so the drop-in model can rebuild them on the GPU box)."""
    torch.manual_seed(0)
    x = detfill.det_image((1, 3, 256, 256), 11)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i) for i in range(3)]
    m = detfill.fill_(CLC(N=64).eval(), seed=0)
    with torch.no_grad():
        out = m(x, refs)
        out_noref = m(x, None)
    bpp = sum(torch.log(l).sum() / (-math.log(2) * 256 * 256) for l in out["likelihoods"].values())
    _save("clc_cfg1.npz", x_hat=out["x_hat"].half(), lik_y=out["likelihoods"]["y"], lik_z=out["likelihoods"]["z"],
          means=out["para"]["means"], scales=out["para"]["scales"], y=out["para"]["y"], bpp=bpp,
          x_hat_mean=out["x_hat"].double().mean(), noref_lik_y=out_noref["likelihoods"]["y"],
          noref_x_hat=out_noref["x_hat"].half(), n_params=sum(p.numel() for p in m.parameters()))
    t = detfill.fill_(TCM(N=64).eval(), seed=0)
    with torch.no_grad():
        o = t(x)
    bpp_t = sum(torch.log(l).sum() / (-math.log(2) * 256 * 256) for l in o["likelihoods"].values())
    _save("tcm_cfg1.npz", x_hat=o["x_hat"].half(), lik_y=o["likelihoods"]["y"], lik_z=o["likelihoods"]["z"],
          bpp=bpp_t, n_params=sum(p.numel() for p in t.parameters()))


def golden_coder(CLC, TCM):
    """Bitstream path: the reference's own CLC.compress / decompress (CLC_run.py:629-716, :738-814)
    and TCM.compress (tcm.py), run unmodified through the shim's range coder, cfg1 inputs and
    detfill weights; plus the CDF tables its update() builds and a raw coder vector with bypass symbols."""
    import random

    from compressai import ans
    x = detfill.det_image((1, 3, 256, 256), 11)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i) for i in range(3)]
    m = detfill.fill_(CLC(N=64).eval(), seed=0)
    m.update()
    captured = {}
    enc_cls = ans.BufferedRansEncoder

    class Spy(enc_cls):                      # records what the reference hands to the coder
        def encode_with_indexes(self, symbols, indexes, *a):
            captured["symbols"], captured["indexes"] = list(symbols), list(indexes)
            return super().encode_with_indexes(symbols, indexes, *a)

    import models.CLC_run as R
    R.BufferedRansEncoder = Spy
    try:
        with torch.no_grad():
            out = m.compress(x, refs)
            rec = m.decompress(out["strings"], out["shape"], refs)
    finally:
        R.BufferedRansEncoder = enc_cls
    gc, eb = m.gaussian_conditional, m.entropy_bottleneck
    t = detfill.fill_(TCM(N=64).eval(), seed=0)
    t.update()
    with torch.no_grad():
        out_t = t.compress(x)
    # raw coder vector: random tables, out-of-range symbols on both sides (bypass coding)
    rnd = random.Random(5)
    cdfs, sizes, offsets = [], [], []
    for _ in range(6):
        n = rnd.randint(3, 30)
        pmf = [rnd.random() ** 3 + 1e-6 for _ in range(n)]
        tot = sum(pmf)
        c = ans.pmf_to_quantized_cdf([p / tot for p in pmf] + [1e-9], 16)
        cdfs.append(c + [0] * (34 - len(c)))
        sizes.append(len(c))
        offsets.append(-(n // 2))
    idx = [rnd.randrange(6) for _ in range(3000)]
    sym = [rnd.randint(offsets[i] - 9, offsets[i] + sizes[i] + (70000 if rnd.random() < 0.02 else 3)) for i in idx]
    raw = ans.RansEncoder().encode_with_indexes(sym, idx, cdfs, sizes, offsets)
    _save("coder.npz",
          y_string=np.frombuffer(out["strings"][0][0], dtype=np.uint8), z_string=np.frombuffer(out["strings"][1][0], dtype=np.uint8),
          shape=np.array(list(out["shape"])), symbols=np.array(captured["symbols"], dtype=np.int16),
          indexes=np.array(captured["indexes"], dtype=np.int8), x_hat=rec["x_hat"].half(),
          gc_cdf=gc.quantized_cdf.numpy().astype(np.int32), gc_cdf_length=gc.cdf_length.numpy().astype(np.int32),
          gc_offset=gc.offset.numpy().astype(np.int32), eb_cdf=eb.quantized_cdf.numpy().astype(np.int32),
          eb_cdf_length=eb.cdf_length.numpy().astype(np.int32), eb_offset=eb.offset.numpy().astype(np.int32),
          tcm_y_bytes=np.array([len(out_t["strings"][0][0])]), tcm_z_bytes=np.array([len(out_t["strings"][1][0])]),
          raw_cdfs=np.array(cdfs, dtype=np.int32), raw_sizes=np.array(sizes, dtype=np.int32),
          raw_offsets=np.array(offsets, dtype=np.int32), raw_idx=np.array(idx, dtype=np.int32),
          raw_sym=np.array(sym, dtype=np.int32), raw_string=np.frombuffer(raw, dtype=np.uint8))


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    CLC, TCM = ref_loader.import_models()
    golden_gaussian(CLC)
    golden_match()
    golden_clm()
    golden_clm_full()
    golden_rd_loss()
    golden_model(CLC, TCM)
    golden_coder(CLC, TCM)


if __name__ == "__main__":
    main()
