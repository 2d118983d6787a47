"""GPU parity: reference matching (fp32 mode) and CLM fusion vs oracle + reference golden.
Bars: top-k match indices bit-exact in fp32 mode; values / gathered features to fp32 round-off."""
import types

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _args(k, stack=False):
    return types.SimpleNamespace(num_k=k, temperature=15, is_stack=stack, single_layer=0)


def test_match_reference_golden():
    import clc_b200
    from oracle import clc_oracle as O
    g = load_golden("match.npz")
    N, C, h, w, p, k = [int(v) for v in g["geom"]]
    d = _dev()
    y, r = g["y"].to(d), g["r"].to(d)
    mask = clc_b200.create_gaussian_masks(h, w, p, p)
    assert torch.equal(mask.cpu(), g["mask"])
    assert torch.equal(clc_b200.create_gaussian_masks(9, 6, 3, 3).cpu(), g["mask_odd"])
    q0 = O.extract_patches(g["y"][:1], p, p).to(d)
    corr0 = clc_b200.L2_or_pearson_corr(q0, r[:1], p, p)
    assert corr0.shape == g["corr0"].shape
    assert (corr0.cpu() - g["corr0"]).abs().max().item() < 2e-6
    P = q0.shape[0]
    val, idx = clc_b200.topk_rows((corr0 * mask).reshape(P, -1), k)
    assert torch.equal(idx.cpu().long(), g["topk_idx"]), "top-k indices must be bit-exact"
    assert torch.allclose(val.cpu(), g["topk_val"], atol=2e-6)
    assert torch.allclose(clc_b200.SI_Wraper(corr0, p, p, P, r[:1], 1, 15, False).cpu(), g["wr_k1"], atol=1e-5)
    assert torch.equal(clc_b200.SI_Wraper(corr0 * mask, p, p, P, r[:1], k, 15, True).cpu(), g["wr_stack"])
    f = clc_b200.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1"], _args(k), mask=mask)["1"]
    assert torch.allclose(f.cpu(), g["f_mask"], atol=2e-5)
    f = clc_b200.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1"], _args(k))["1"]
    assert torch.allclose(f.cpu(), g["f_nomask"], atol=2e-5)
    fm = clc_b200.SI_Finder_at_Decoder_Feature_Domain(y, r, p, p, r, ["1", "2"], _args(k), mask=mask,
                                                      other_ys=[g["r_half"].to(d)])
    assert torch.allclose(fm["1"].cpu(), g["f_multi_1"], atol=2e-5)
    assert torch.allclose(fm["2"].cpu(), g["f_multi_2"], atol=2e-5)


@pytest.mark.parametrize("geom", [(8, 320, 16, 16, 4), (3, 320, 32, 48, 4), (2, 24, 6, 10, 2), (1, 7, 9, 6, 3)])
def test_match_topk_indices_bit_exact_vs_oracle(geom):
    """SURVEY 8d match distribution: y~N(0,1), ref = 0.5*y + N(0,1); k=4, T=15, mask on."""
    import clc_b200
    from oracle import clc_oracle as O
    B, C, h, w, p = geom
    k = 4
    g = torch.Generator().manual_seed(B * 1000 + C)
    y = torch.randn(B, C, h, w, generator=g)
    r = 0.5 * y + torch.randn(B, C, h, w, generator=g)
    mask = O.gaussian_masks(h, w, p, p)
    outs, val_o, idx_o = O.si_finder(y, r, p, p, r, k, 15.0, mask=mask, return_index=True)
    d = _dev()
    val, idx, _ = clc_b200.match_topk(y.to(d), [r.to(d)], p, p, k, gaussian_mask=True, mode="fp32")
    assert torch.equal(idx[:, 0].cpu().long(), idx_o), "top-k match indices must be bit-exact (fp32 mode)"
    assert torch.allclose(val[:, 0].cpu(), val_o, atol=3e-6)
    out = clc_b200.match_and_gather(y.to(d), [r.to(d)], p, p, k, 15.0, True, False, mode="fp32")
    assert torch.allclose(out[:, 0].cpu(), outs[0], atol=3e-5)
    # margin audit: the smallest gap between ranked values must dwarf fp32 summation noise
    full, _ = O.topk_lowest_index(torch.cat([(O.pearson_corr(O.extract_patches(y[n:n + 1], p, p), r[n:n + 1]) * mask)
                                             .reshape(-1, mask.shape[2] * mask.shape[3]) for n in range(B)]), k + 1)
    assert (full[:, :-1] - full[:, 1:]).min().item() > 5e-7


def test_match_multi_ref_and_self_match():
    import clc_b200
    from oracle import clc_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(5)
    B, R, C, h, w, p, k = 2, 3, 32, 8, 12, 4, 2
    y = torch.randn(B, C, h, w, generator=g)
    refs = [0.5 * y + torch.randn(B, C, h, w, generator=g) for _ in range(R)]
    val, idx, _ = clc_b200.match_topk(y.to(d), [t.to(d) for t in refs], p, p, k, gaussian_mask=False, mode="fp32")
    for j in range(R):
        _, v_o, i_o = O.si_finder(y, refs[j], p, p, refs[j], k, 15.0, return_index=True)
        assert torch.equal(idx[:, j].cpu().long(), i_o)
        assert torch.allclose(val[:, j].cpu(), v_o, atol=3e-6)
    # self-match with k=1 reconstructs the input exactly (SURVEY 8c probe)
    out = clc_b200.match_and_gather(y.to(d), [y.to(d)], p, p, 1, 15.0, False, False, mode="fp32")
    assert torch.equal(out[:, 0].cpu(), y)


def test_topk_rows_ties_and_order():
    import clc_b200
    d = _dev()
    x = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0], [5.0, 5.0, 5.0, 5.0, 5.0], [-1.0, -2.0, float("nan"), 0.0, -0.0]])
    val, idx = clc_b200.topk_rows(x.to(d), 3)
    assert idx.cpu().tolist()[0] == [1, 2, 4] and idx.cpu().tolist()[1] == [0, 1, 2]
    assert idx.cpu().tolist()[2][0] == 2  # NaN ranks first, as torch.topk
    g = torch.Generator().manual_seed(0)
    big = torch.randn(37, 9625, generator=g)
    val, idx = clc_b200.topk_rows(big.to(d), 8)
    tv, ti = torch.topk(big, 8, dim=1)
    assert torch.equal(val.cpu(), tv) and torch.equal(idx.cpu().long(), ti)


def test_match_backward_vs_oracle_autograd():
    """Gradients through gather/blend, softmax weights and the masked Pearson values, incl. the
    reference's detached-conv-weights semantics for the query (Patch_Matching.py:869)."""
    import clc_b200
    from oracle import clc_oracle as O
    d = _dev()
    g = torch.Generator().manual_seed(11)
    B, C, h, w, p, k = 2, 16, 8, 12, 4, 3
    y = torch.randn(B, C, h, w, generator=g)
    r = 0.5 * y + torch.randn(B, C, h, w, generator=g)
    wgt = torch.randn(B, C, h, w, generator=g)
    mask = O.gaussian_masks(h, w, p, p)
    yo, ro = y.clone().requires_grad_(True), r.clone().requires_grad_(True)
    out_o = O.si_finder(yo, ro, p, p, ro, k, 15.0, mask=mask)[0]
    (out_o * wgt).sum().backward()
    yc, rc = y.to(d).requires_grad_(True), r.to(d).requires_grad_(True)
    out_c = clc_b200.match_and_gather(yc, [rc], p, p, k, 15.0, True, False, mode="fp32")[:, 0]
    assert torch.allclose(out_c.detach().cpu(), out_o.detach(), atol=3e-5)
    (out_c * wgt.to(d)).sum().backward()
    for name, a, b in (("ref", rc.grad.cpu(), ro.grad), ("query", yc.grad.cpu(), yo.grad)):
        s = b.abs().max().item()
        assert s > 0, name
        assert (a - b).abs().max().item() / s < 2e-3, (name, (a - b).abs().max().item() / s)


def test_clm_fuse_and_simple_clm_vs_reference():
    import clc_b200
    from oracle import clc_oracle as O
    d = _dev()
    g = load_golden("clm.npz")
    m = clc_b200.SimpleCLM(16)
    m.load_state_dict({k[3:].replace("__", "."): v for k, v in g.items() if k.startswith("sd_")})
    m = m.to(d)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = m(g["y"].to(d), [t.to(d) for t in g["refs"]])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert torch.allclose(out.cpu(), g["out"], atol=2e-5)
    # elementwise core, forward + backward, several shapes incl. non-multiple-of-4 spatial size
    for (R, B, C, H, W) in [(3, 8, 320, 16, 16), (1, 2, 5, 3, 7), (5, 1, 64, 32, 48), (8, 2, 16, 4, 4)]:
        gg = torch.Generator().manual_seed(R * 100 + C)
        ref_t = torch.randn(R, B, C, H, W, generator=gg)
        att = 2 * torch.randn(R, B, 1, H, W, generator=gg)
        y = torch.randn(B, C, H, W, generator=gg)
        wgt = torch.randn(B, C, H, W, generator=gg)
        ro, ao, yo = (t.clone().requires_grad_(True) for t in (ref_t, att, y))
        oo = O.clm_fuse(ro, ao, yo)
        (oo * wgt).sum().backward()
        rc, ac, yc = (t.to(d).requires_grad_(True) for t in (ref_t, att, y))
        oc = clc_b200.clm_fuse(rc, ac, yc)
        (oc * wgt.to(d)).sum().backward()
        assert torch.allclose(oc.detach().cpu(), oo.detach(), atol=2e-6, rtol=1e-5)
        assert torch.allclose(rc.grad.cpu(), ro.grad, atol=1e-6, rtol=1e-5)
        assert torch.allclose(yc.grad.cpu(), yo.grad)
        s = ao.grad.abs().max().item()
        assert (ac.grad.cpu() - ao.grad).abs().max().item() / s < 1e-4
