"""GPU model-level parity (SURVEY.md 7.1 level B): the drop-in CLC / TCM running the fused
sm_100a kernels vs (a) the committed outputs of the reference's own forward and (b) the same
backbone evaluated with the oracle's entropy arithmetic on the same device.
Bars: quantised symbols bit-exact, likelihood 1e-4 rel, bpp 1e-3, PSNR 0.01 dB."""
import math

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_convs():
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


def _psnr(a, b):
    return -10 * math.log10(torch.mean((a - b) ** 2).item())


def _bpp(out, npix):
    return sum(torch.log(l.double()).sum().item() for l in out["likelihoods"].values()) / (-math.log(2) * npix)


def test_clc_cfg1_vs_reference_golden_and_oracle_mode():
    import clc_b200
    from clc_b200.models import CLC
    from oracle import detfill
    from oracle.model_oracle import to_oracle_mode
    g = load_golden("clc_cfg1.npz")
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    x = detfill.det_image((1, 3, 256, 256), 11).to(d)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i).to(d) for i in range(3)]
    with torch.no_grad():
        out = m(x, refs)
        out_o = to_oracle_mode(m).eval()(x, refs)
    # (b) same backbone, oracle entropy arithmetic, same device: everything upstream of the
    # entropy stage is bit-identical, so the symbols must be too.
    assert torch.equal(out["para"]["y"], out_o["para"]["y"])
    assert torch.equal(out["para"]["means"], out_o["para"]["means"])
    assert torch.equal(out["para"]["scales"], out_o["para"]["scales"])
    assert torch.equal(out["x_hat"], out_o["x_hat"]), "x_hat differs => y_hat symbols differ"
    for key in ("y", "z"):
        a, b = out["likelihoods"][key], out_o["likelihoods"][key]
        big = b > 1e-9
        assert ((a.double() - b.double()).abs() / b.double())[big].max().item() < 1e-4
    assert abs(_bpp(out, 65536) - _bpp(out_o, 65536)) < 1e-3
    # (a) reference's own forward on the CPU (golden): conv round-off differs between CPU and GPU,
    # so compare the judged aggregates.
    assert abs(_bpp(out, 65536) - g["bpp"].item()) < 1e-3
    ref_x_hat = g["x_hat"].float().to(d)
    assert abs(_psnr(x, out["x_hat"]) - _psnr(x, ref_x_hat)) < 0.01
    lik, ref = out["likelihoods"]["y"].cpu(), g["lik_y"]
    assert ((lik - ref).abs() <= 1e-3 * ref + 1e-9).float().mean().item() > 0.99
    # rate term through the drop-in loss
    crit = clc_b200.RateDistortionLoss(lmbda=0.013)
    loss = crit(out, x)
    assert abs(loss["bpp_loss"].item() - g["bpp"].item()) < 1e-3


def test_tcm_cfg1_vs_reference_golden():
    from clc_b200.models import TCM
    from oracle import detfill
    g = load_golden("tcm_cfg1.npz")
    d = torch.device("cuda:0")
    m = detfill.fill_(TCM(N=64), seed=0).eval().to(d)
    x = detfill.det_image((1, 3, 256, 256), 11).to(d)
    with torch.no_grad():
        out = m(x)
    assert abs(_bpp(out, 65536) - g["bpp"].item()) < 1e-3
    assert abs(_psnr(x, out["x_hat"]) - _psnr(x, g["x_hat"].float().to(d))) < 0.01


def test_clc_training_step_gradients_vs_oracle_mode():
    """Train mode with the uniform noise injected as explicit tensors: loss and
    parameter gradients of the fused path vs autograd through the oracle arithmetic."""
    import clc_b200
    from clc_b200.models import CLC
    from oracle import detfill
    from oracle.model_oracle import to_oracle_mode
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=1).train().to(d)
    mo = to_oracle_mode(m).train()
    B = 2
    x = detfill.det_image((B, 3, 256, 256), 21).to(d)
    refs = [detfill.det_image((B, 3, 256, 256), 22 + i).to(d) for i in range(3)]
    gen = torch.Generator().manual_seed(9)
    noise = {"y": (torch.rand(B, 320, 16, 16, generator=gen) - 0.5).to(d),
             "z": (torch.rand(B, 192, 4, 4, generator=gen) - 0.5).to(d)}
    crit = clc_b200.RateDistortionLoss(lmbda=0.013)
    out = m(x, refs, noise=noise)
    loss = crit(out, x)["loss"]
    loss.backward()
    out_o = mo(x, refs, noise=noise)
    npix = B * 256 * 256
    bpp_o = sum(torch.log(l).sum() / (-math.log(2) * npix) for l in out_o["likelihoods"].values())
    loss_o = 0.013 * 255 ** 2 * torch.nn.functional.mse_loss(out_o["x_hat"], x) + bpp_o
    loss_o.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-3 * max(1.0, abs(loss_o.item()))
    checked = 0
    po = dict(mo.named_parameters())
    for n, p in m.named_parameters():
        if n.startswith(("gaussian_conditional", "entropy_bottleneck")):
            n_o = n.replace("entropy_bottleneck.", "entropy_bottleneck.inner.")
        else:
            n_o = n
        q = po.get(n_o)
        if p.grad is None or q is None or q.grad is None:
            continue
        s = q.grad.abs().max().item()
        if s < 1e-12:
            continue
        err = (p.grad - q.grad).abs().max().item() / s
        assert err < 5e-3, (n, err)
        checked += 1
    assert checked > 300
    unused = [n for n, p in m.named_parameters() if p.grad is None]
    assert any(n.startswith("feature_alignment") for n in unused)  # as in the reference


def test_clc_match_refs_extended_wiring_and_symbols():
    """Level C wiring (match_refs=True) runs end to end, and the coder-input tensors
    (int32 symbols / scale-table indexes, CLC_run.py:689-690) stay on the device."""
    from clc_b200.models import CLC
    from oracle import detfill
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64, match_refs=True, match_mode="fp32"), seed=2).eval().to(d)
    x = detfill.det_image((1, 3, 256, 256), 31).to(d)
    refs = [detfill.det_image((1, 3, 256, 256), 32 + i).to(d) for i in range(3)]
    with torch.no_grad():
        out = m(x, refs)
        sym = m.symbols_and_indexes(x, refs)
    assert out["x_hat"].shape == x.shape and torch.isfinite(out["x_hat"]).all()
    assert sym["symbols"].dtype == torch.int32 and sym["symbols"].shape == (1, 320, 16, 16)
    assert sym["indexes"].min().item() >= 0 and sym["indexes"].max().item() <= 63
    y, mu = out["para"]["y"], out["para"]["means"]
    assert torch.equal(sym["symbols"], torch.round(y - mu).int())


@pytest.mark.parametrize("mode", ["tc", "fp32"])
def test_clc_match_refs_model_parity_vs_oracle_si_finder(mode):
    """Level-C wiring at MODEL level: CLC(match_refs=True) with the tensor-core match path (the default) vs the
    same weights in oracle mode, where the alignment is the oracle's SI_Finder_at_Decoder_Feature_Domain
    (Patch_Matching.py:157-216 restated, CPU fp32).  Match indices bit-exact, aligned context / likelihoods /
    x_hat to conv round-off."""
    import clc_b200.models as M
    from clc_b200 import matching
    from oracle import detfill
    from oracle.model_oracle import to_oracle_mode
    d = torch.device("cuda:0")
    m = detfill.fill_(M.CLC(N=64, match_refs=True, match_mode=mode), seed=2).eval().to(d)
    mo = to_oracle_mode(m).eval()
    x = detfill.det_image((2, 3, 256, 256), 41).to(d)
    refs = [detfill.det_image((2, 3, 256, 256), 42 + i).to(d) for i in range(3)]
    with torch.no_grad():
        out = m(x, refs)
        out_o = mo(x, refs)
        # the indices the product path selected, on the same latents the oracle saw
        y = m.g_a(x)
        feats = m.ref_encoder(torch.cat(refs, 0)).view(3, 2, 320, 16, 16).transpose(0, 1).contiguous()
        _, idx, _ = matching.match_topk(y, feats, 4, 4, 4, gaussian_mask=True, mode=mode)
    assert torch.equal(idx.cpu().long(), mo.oracle_match_idx[0]), "model-level match indices differ from the oracle"
    lik, lik_o = out["likelihoods"]["y"], out_o["likelihoods"]["y"]
    close = ((lik - lik_o).abs() <= 1e-3 * lik_o + 1e-9).float().mean().item()
    assert close > 0.995, close                                    # conv round-off may flip a few symbols
    assert torch.allclose(out["likelihoods"]["z"], out_o["likelihoods"]["z"], rtol=1e-3, atol=1e-9)
    mse = torch.mean((out["x_hat"] - out_o["x_hat"]) ** 2).item()
    psnr_gap = abs(10 * math.log10(1.0 / max(torch.mean((out["x_hat"] - x) ** 2).item(), 1e-12)) -
                   10 * math.log10(1.0 / max(torch.mean((out_o["x_hat"] - x) ** 2).item(), 1e-12)))
    assert mse < 1e-6 and psnr_gap < 0.01


def test_graphed_forward_equals_eager_forward():
    """make_graphed_forward (whole CLC.forward in one CUDA graph, mean / scale branches forked) returns what the
    eager forward returns, for new inputs too (static buffers are refilled), with and without references."""
    from clc_b200.models import CLC
    from oracle import detfill
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    x = detfill.det_image((1, 3, 256, 256), 11).to(d)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i).to(d) for i in range(3)]
    run = m.make_graphed_forward(x, refs)
    x2 = detfill.det_image((1, 3, 256, 256), 41).to(d)
    refs2 = [detfill.det_image((1, 3, 256, 256), 42 + i).to(d) for i in range(3)]
    for xi, ri in ((x, refs), (x2, refs2)):
        with torch.no_grad():
            want = m(xi, ri)
        got = run(xi, ri)
        torch.cuda.synchronize()
        assert torch.equal(got["para"]["y"], want["para"]["y"])
        assert torch.equal(got["x_hat"], want["x_hat"])
        for key in ("y", "z"):
            assert torch.equal(got["likelihoods"][key], want["likelihoods"][key])
    with pytest.raises(ValueError):
        run(x, refs[:2])
    m.train()
    with pytest.raises(RuntimeError):
        m.make_graphed_forward(x, refs)


@pytest.mark.parametrize("tf32_matmul", [False, True])
def test_graphed_forward_channels_last_within_conv_roundoff(tf32_matmul):
    """channels_last weights / inputs (cuDNN NHWC kernels, no layout-conversion kernels), optionally TF32 tensor
    cores for the Linear layers: same rate and distortion as the NCHW fp32-matmul forward up to round-off
    (measured: bpp 4e-4, PSNR 2e-4 dB, 0.007 % of the symbols with TF32 matmuls)."""
    from clc_b200.models import CLC
    from oracle import detfill
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    x = detfill.det_image((1, 3, 256, 256), 11).to(d)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i).to(d) for i in range(3)]
    with torch.no_grad():
        want = m(x, refs)
    bpp_want, psnr_want = _bpp(want, 65536), _psnr(x, want["x_hat"])
    run = m.make_graphed_forward(x, refs, channels_last=True, tf32_matmul=tf32_matmul)
    assert torch.backends.cuda.matmul.allow_tf32 is False          # restored after the capture
    got = run(x, refs)
    torch.cuda.synchronize()
    assert abs(_bpp(got, 65536) - bpp_want) < (5e-3 if tf32_matmul else 1e-3)
    assert abs(_psnr(x, got["x_hat"]) - psnr_want) < 0.01
    moved = (got["para"]["y"].round() != want["para"]["y"].round()).float().mean().item()
    assert moved < 2e-3, moved
