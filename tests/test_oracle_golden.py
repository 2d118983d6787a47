"""CPU: pin the oracle restatements against the golden vectors produced by the reference's own
code (oracle/make_golden.py), and against the known-answer values recorded in SURVEY.md 8c."""
import math

import pytest
import torch

from conftest import load_golden
from oracle import clc_oracle as O
from oracle import ref_loader


def test_gaussian_known_answers():
    g = load_golden("gaussian.npz")
    # SURVEY.md 8c KATs (reference lines CLC_run.py:718-736, torch CPU fp32)
    kat = torch.tensor([3.8292491436e-01, 6.8268942833e-01, 2.7408441383e-06, 1.2097758800e-01, 0,
                        6.5957138488e-31, 6.2334239483e-03, 0])
    assert torch.allclose(g["kat_lik"], kat, rtol=2e-6, atol=0)
    assert torch.equal(g["kat_y_hat"], torch.tensor([0, 0.1, 1, -1.5, 5, 12, 0.1, 40.0]))
    lik = O.gaussian_likelihood(g["kat_y_hat"], g["kat_scale"], g["kat_mu"])
    assert torch.equal(lik, g["kat_lik"])
    m_log2 = -torch.log2(lik.clamp_min(1e-9))
    exp = torch.tensor([1.3848665953, 0.5506986976, 18.4769477844, 3.0471882820, 29.8973522186,
                        29.8973522186, 7.3257594109, 29.8973522186])
    assert torch.allclose(m_log2, exp, rtol=1e-6)
    assert torch.equal(g["rounds"], torch.tensor([0, 2, 2, -0.0, -2, 0, 3.0]))
    st = O.get_scale_table()
    assert torch.equal(st, g["scale_table"]) and st.numel() == 64
    assert abs(st[0].item() - 0.11) < 1e-7 and abs(st[63].item() - 256.0) < 1e-3


def test_gaussian_conditional_shim_matches_reference_likelihood():
    g = load_golden("gaussian.npz")
    y, mu, sc, noise = g["y"], g["mu"], g["scale"], g["noise"]
    out_e, lik_e, y_hat = O.gc_forward(y, sc, mu)
    assert torch.equal(y_hat, g["y_hat"]) and torch.equal(out_e, g["y_hat"])
    assert torch.equal(lik_e, g["lik_eval"].clamp_min(1e-9))
    out_t, lik_t, _ = O.gc_forward(y, sc, mu, noise=noise)
    assert torch.equal(out_t, y + noise)
    assert torch.equal(lik_t, g["lik_train"].clamp_min(1e-9))
    # fp32 self-noise of the reference arithmetic vs fp64 (what a 1e-4 bar has to absorb)
    big = g["lik_eval64"] > 1e-9
    rel = ((g["lik_eval"].double() - g["lik_eval64"]).abs() / g["lik_eval64"])[big].max().item()
    assert rel < 2e-4


def test_lower_bound_gradient_gate():
    x = torch.tensor([0.05, 0.2, 0.05, 0.2], requires_grad=True)
    lb = O.LowerBound(0.11)
    lb(x).backward(torch.tensor([1.0, 1.0, -1.0, -1.0]))
    assert torch.equal(x.grad, torch.tensor([0.0, 1.0, -1.0, -1.0]))


def test_build_indexes_and_symbols():
    st = O.get_scale_table()
    scale = torch.tensor([0.01, 0.11, 0.12, 0.125, 5.0, 255.0, 256.0, 1e4, -3.0])
    y = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, 2.5000002, 7.2, -7.7])
    sym, idx = O.gc_symbols_indexes(y, scale, torch.zeros_like(y), st)
    assert sym.dtype == torch.int32 and idx.dtype == torch.int32
    assert sym.tolist() == [0, 2, 2, 0, -2, 0, 3, 7, -8]
    assert idx.tolist()[0] == 0 and idx.tolist()[1] == 0 and idx.tolist()[-3:] == [63, 63, 0]
    import numpy as np
    tab = st.numpy()
    brute = [min(63, next((i for i, t in enumerate(tab) if max(np.float32(s), np.float32(0.11)) <= t), 63))
             for s in scale.numpy()]
    assert idx.tolist() == brute


def test_match_oracle_vs_reference_golden():
    g = load_golden("match.npz")
    N, C, h, w, p, k = [int(v) for v in g["geom"]]
    y, r = g["y"], g["r"]
    assert torch.equal(O.gaussian_masks(h, w, p, p), g["mask"])
    assert torch.equal(O.gaussian_masks(9, 6, 3, 3), g["mask_odd"])
    q0 = O.extract_patches(y[:1], p, p)
    corr0 = O.pearson_corr(q0, r[:1])
    assert torch.allclose(corr0, g["corr0"], rtol=0, atol=1e-6)
    # fp64 Pearson: fp32 map is the textbook coefficient to ~1e-6
    assert (O.pearson_corr(q0, r[:1], torch.float64) - g["corr0"].double()).abs().max() < 2e-6
    P = q0.shape[0]
    assert torch.allclose(O.si_wrapper(g["corr0"], p, p, P, r[:1], 1, 15, False), g["wr_k1"], atol=1e-6)
    assert torch.equal(O.si_wrapper(g["corr0"] * g["mask"], p, p, P, r[:1], k, 15, True), g["wr_stack"])
    val, idx = O.topk_lowest_index((g["corr0"] * g["mask"]).reshape(P, -1), k)
    assert torch.equal(idx, g["topk_idx"]) and torch.equal(val, g["topk_val"])
    assert torch.allclose(O.si_finder(y, r, p, p, r, k, 15, mask=g["mask"])[0], g["f_mask"], atol=1e-5)
    assert torch.allclose(O.si_finder(y, r, p, p, r, k, 15)[0], g["f_nomask"], atol=1e-5)
    multi = O.si_finder(y, r, p, p, r, k, 15, mask=g["mask"], other_ys=[g["r_half"]])
    assert torch.allclose(multi[0], g["f_multi_1"], atol=1e-5) and torch.allclose(multi[1], g["f_multi_2"], atol=1e-5)


def test_self_match_reconstructs_input():
    torch.manual_seed(0)
    y = torch.randn(1, 8, 8, 8)
    out, val, idx = O.si_finder(y, y, 4, 4, y, 1, 15, return_index=True)
    assert torch.allclose(out[0], y, atol=1e-6)
    assert torch.allclose(val, torch.ones_like(val), atol=1e-5)


def test_clm_oracle_vs_reference_golden():
    g = load_golden("clm.npz")
    m = O.SimpleCLM(16)
    m.load_state_dict({k[3:].replace("__", "."): v for k, v in g.items() if k.startswith("sd_")})
    out = m(g["y"], list(g["refs"]))
    assert torch.allclose(out, g["out"], atol=1e-6)


def test_clm_variant_a_oracle_vs_reference_golden():
    """CLM.forward (models/CLM.py:84-128) incl. the Python-loop DeformableAlignment (:5-60): vectorised
    restatement vs the reference's own output and hooked intermediates."""
    g = load_golden("clm_full.npz")
    m = O.CLM(16, temperature=0.5)
    m.load_state_dict({k[3:].replace("__", "."): v for k, v in g.items() if k.startswith("sd_")})
    with torch.no_grad():
        out, colsum, aligned = m(g["y"], list(g["refs"]), return_parts=True)
    assert torch.allclose(colsum, g["colsum"], rtol=1e-5, atol=1e-6)
    # (the x8 offsets amplify the 1e-7 summation-order difference of weighted_x into ~1e-5 in the samples)
    assert torch.allclose(aligned, g["aligned"], atol=2e-5)
    assert torch.allclose(out, g["out"], atol=5e-6)
    # the taps really exercise the quirks: some fall outside the image, some leave their own cell
    B, C, H, W = g["y"].shape
    ws = g["refs"][0] * g["colsum"][0].view(B, 1, H, W)
    off = m.alignment.offset_conv(torch.cat([g["refs"][0], ws], 1)).view(B, 9, 2, H, W)
    oh = off[:, :, 0] + torch.arange(H).view(1, 1, H, 1)
    assert (oh < 0).any() and (oh > H - 1).any() and (off.abs() > 1).any()


def test_rd_loss_oracle_vs_reference_golden():
    g = load_golden("rd_loss.npz")
    out = O.rate_distortion_loss({"likelihoods": {"y": g["lik_y"], "z": g["lik_z"]}, "x_hat": g["x_hat"]},
                                 g["x"], 0.013)
    for key in ("bpp_loss", "mse_loss", "loss"):
        assert torch.allclose(out[key], g[key], rtol=1e-6), key


def test_entropy_bottleneck_shim_properties():
    """No reference vector exists for EntropyBottleneck (parity unpinned): check the published
    algorithm's invariants instead -- pmf sums to one over the integers, likelihood floor, medians."""
    torch.manual_seed(0)
    eb = O.EntropyBottleneck(4)
    ks = torch.arange(-60, 61, dtype=torch.float32).reshape(1, 1, -1, 1).repeat(1, 4, 1, 1)
    _, lik, z_hat = O.eb_forward(eb, ks)
    # the pmf telescopes: sum_k lik(k) = cdf(60.5) - cdf(-60.5), cdf = sigmoid(logits)
    ends = torch.tensor([-60.5, 60.5]).reshape(1, 1, 2).repeat(4, 1, 1)
    cdf = torch.sigmoid(eb._logits_cumulative(ends, stop_gradient=True))
    tele = (cdf[:, 0, 1] - cdf[:, 0, 0])
    assert torch.allclose(lik.sum(dim=2).flatten(), tele, atol=1e-4) and tele.min() > 0.99
    assert lik.min() >= 1e-9 and torch.equal(z_hat, ks)
    noise = torch.rand(1, 4, 121, 1) - 0.5
    out, lik_t, _ = O.eb_forward(eb, ks, noise=noise)
    assert torch.equal(out, ks + noise)
    t = math.log(2 / 1e-9 - 1)
    assert torch.allclose(eb.target, torch.tensor([-t, 0, t]))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
def test_reference_itself_reproduces_committed_golden():
    """In the build container: re-run the reference's own functions and compare with the fixtures."""
    pm = ref_loader.load_patch_matching()
    g = load_golden("match.npz")
    N, C, h, w, p, k = [int(v) for v in g["geom"]]
    q0 = O.extract_patches(g["y"][:1], p, p)
    assert torch.equal(pm.L2_or_pearson_corr(q0, g["r"][:1], p, p), g["corr0"])
    assert torch.equal(pm.create_gaussian_masks(h, w, p, p), g["mask"])
