"""GPU: compress() / decompress() of the drop-in CLC / TCM (CLC_run.py:629-716, :738-814) --
device kernels produce the symbols and scale-table indexes, the host C coder the strings."""
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
    return -10 * math.log10(torch.mean((a.float() - b.float()) ** 2).item())


def _inputs(d, B=1):
    from oracle import detfill
    x = torch.cat([detfill.det_image((1, 3, 256, 256), 11 + 100 * b) for b in range(B)]).to(d)
    refs = [torch.cat([detfill.det_image((1, 3, 256, 256), 12 + i + 100 * b) for b in range(B)]).to(d) for i in range(3)]
    return x, refs


def test_clc_compress_decompress_round_trip_and_reference_golden():
    import oracle
    from clc_b200.models import CLC
    from oracle import detfill
    oracle.enable_shim()
    from compressai import ans as O
    g = load_golden("coder.npz")
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    with pytest.raises(ValueError, match="Uninitialized CDFs"):
        m.compress(*_inputs(d))
    m.update()
    x, refs = _inputs(d)
    with torch.no_grad():
        out = m.compress(x, refs)
        rec = m.decompress(out["strings"], out["shape"], refs)
        fwd = m(x, refs)
        si = m.symbols_and_indexes(x, refs)
    assert tuple(out["shape"]) == (4, 4) and len(out["strings"][0]) == 1 and len(out["strings"][1]) == 1
    # decoding reproduces the eval forward exactly (same symbols => same y_hat => same x_hat)
    assert torch.equal(rec["x_hat"], fwd["x_hat"].clamp(0, 1))
    # the y string is the oracle coder's encoding of the device symbols / indexes (slice-major order)
    tab = m.gaussian_conditional
    sym = torch.cat([s.reshape(-1) for s in si["symbols"].chunk(5, 1)]).cpu()
    idx = torch.cat([s.reshape(-1) for s in si["indexes"].chunk(5, 1)]).cpu()
    n = 6000
    want_prefix_free = O.RansEncoder().encode_with_indexes(
        sym[-n:].tolist(), idx[-n:].tolist(), tab.quantized_cdf.cpu().tolist(), tab.cdf_length.cpu().tolist(),
        tab.offset.cpu().tolist())
    from clc_b200 import ans as A
    assert A.RansEncoder().encode_with_indexes(sym[-n:], idx[-n:], tab.quantized_cdf, tab.cdf_length, tab.offset) \
        == want_prefix_free
    dec = A.RansDecoder().decode_with_indexes(out["strings"][0][0], idx, tab.quantized_cdf, tab.cdf_length,
                                              tab.offset, as_tensor=True)
    assert torch.equal(dec, sym)
    # against the reference's own compress() on the CPU (golden): conv round-off moves a few symbols, so
    # compare stream sizes and reconstruction quality
    ref_bytes = g["y_string"].numel() + g["z_string"].numel()
    got_bytes = len(out["strings"][0][0]) + len(out["strings"][1][0])
    assert abs(got_bytes - ref_bytes) <= 0.005 * ref_bytes, (got_bytes, ref_bytes)
    flips = (sym != g["symbols"].int()).float().mean().item()
    assert flips < 2e-3, flips
    assert abs(_psnr(x.cpu(), rec["x_hat"].cpu()) - _psnr(x.cpu(), g["x_hat"])) < 0.01
    # real rate vs the likelihood estimate
    bpp_est = sum(torch.log(l.double()).sum().item() for l in fwd["likelihoods"].values()) / (-math.log(2) * 65536)
    bpp_real = 8.0 * got_bytes / 65536
    assert abs(bpp_real - bpp_est) < 0.05 * bpp_est + 0.02, (bpp_real, bpp_est)


def test_batched_compress_and_tcm():
    from clc_b200.models import CLC, TCM
    from oracle import detfill
    g = load_golden("coder.npz")
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    m.update()
    x, refs = _inputs(d, B=2)
    with torch.no_grad():
        out = m.compress(x, refs)
        rec = m.decompress(out["strings"], out["shape"], refs)
        fwd = m(x, refs)
    assert len(out["strings"][1]) == 2 and len(out["strings"][0]) == 1      # z: one per image, y: one stream
    assert torch.allclose(rec["x_hat"], fwd["x_hat"].clamp(0, 1), atol=1e-6)
    # without references: the non-ref transforms (CLC_run.py:558-561)
    with torch.no_grad():
        out_n = m.compress(x[:1], None)
        rec_n = m.decompress(out_n["strings"], out_n["shape"], None)
        assert torch.equal(rec_n["x_hat"], m(x[:1], None)["x_hat"].clamp(0, 1))
    t = detfill.fill_(TCM(N=64), seed=0).eval().to(d)
    t.update()
    with torch.no_grad():
        out_t = t.compress(x[:1])
        rec_t = t.decompress(out_t["strings"], out_t["shape"])
        assert torch.equal(rec_t["x_hat"], t(x[:1])["x_hat"].clamp(0, 1))
    ref_bytes = int(g["tcm_y_bytes"][0]) + int(g["tcm_z_bytes"][0])
    got = len(out_t["strings"][0][0]) + len(out_t["strings"][1][0])
    assert abs(got - ref_bytes) <= 0.005 * ref_bytes, (got, ref_bytes)
    mm = CLC(N=64, match_refs=True)
    with pytest.raises(NotImplementedError):
        mm.compress(x, refs)
