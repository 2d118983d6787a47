"""CPU: the drop-in CLC / TCM backbone is a faithful restatement of the reference model.
 (1) in the build container: identical state_dict keys and shapes as the reference classes;
 (2) everywhere: with the oracle's entropy arithmetic plugged in, the drop-in model reproduces the
     committed outputs of the reference's own CLC.forward / TCM.forward (tests/golden/*_cfg1.npz)
     on the name-seeded weights of oracle/detfill.py."""
import math

import pytest
import torch

from conftest import load_golden
from oracle import detfill, ref_loader
from oracle.model_oracle import to_oracle_mode


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("name", ["CLC", "TCM"])
def test_state_dict_keys_match_reference(name):
    import clc_b200.models as M
    RefCLC, RefTCM = ref_loader.import_models()
    ref = (RefCLC if name == "CLC" else RefTCM)(N=64)
    mine = getattr(M, name)(N=64)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert sorted(set(a) - set(b)) == [] and sorted(set(b) - set(a)) == []
    assert a == b
    assert sum(p.numel() for p in ref.parameters()) == sum(p.numel() for p in mine.parameters())


def _check_against_golden(out, g, H=256, W=256):
    bpp = sum(torch.log(l).sum() / (-math.log(2) * H * W) for l in out["likelihoods"].values())
    assert abs(bpp.item() - g["bpp"].item()) < 1e-3                    # north_star: bpp within 1e-3
    mse = torch.mean((out["x_hat"] - g["x_hat"].float()) ** 2).item()
    assert mse < 1e-6                                                  # x_hat stored as fp16
    lik, ref = out["likelihoods"]["y"], g["lik_y"]
    close = ((lik - ref).abs() <= 1e-3 * ref + 1e-9).float().mean().item()
    assert close > 0.995, close                                        # conv round-off may flip a few symbols
    assert torch.allclose(out["likelihoods"]["z"], g["lik_z"], rtol=1e-3, atol=1e-9)


def test_clc_backbone_reproduces_reference_forward():
    import clc_b200.models as M
    g = load_golden("clc_cfg1.npz")
    torch.manual_seed(0)
    m = detfill.fill_(M.CLC(N=64), seed=0).eval()
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    mo = to_oracle_mode(m).eval()
    x = detfill.det_image((1, 3, 256, 256), 11)
    refs = [detfill.det_image((1, 3, 256, 256), 12 + i) for i in range(3)]
    with torch.no_grad():
        out = mo(x, refs)
        out_noref = mo(x, None)
    assert torch.allclose(out["para"]["y"], g["y"], atol=1e-4)
    assert torch.allclose(out["para"]["means"], g["means"], atol=1e-3)
    assert torch.allclose(out["para"]["scales"], g["scales"], atol=1e-3)
    assert (out["para"]["scales"] > 0.11).float().mean() > 0.5          # not only the clamp branch
    _check_against_golden(out, g)
    lik, ref = out_noref["likelihoods"]["y"], g["noref_lik_y"]
    assert ((lik - ref).abs() <= 1e-3 * ref + 1e-9).float().mean().item() > 0.995


def test_tcm_backbone_reproduces_reference_forward():
    import clc_b200.models as M
    g = load_golden("tcm_cfg1.npz")
    m = detfill.fill_(M.TCM(N=64), seed=0).eval()
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    with torch.no_grad():
        out = to_oracle_mode(m).eval()(detfill.det_image((1, 3, 256, 256), 11))
    _check_against_golden(out, g)


def test_clc_loads_checkpoint_saved_after_update():
    """A checkpoint saved after update() carries non-empty CDF tables for BOTH entropy models; the
    reference (compressai CompressionModel.load_state_dict, CLC_run.py:599-618) resizes the empty buffers so
    that it loads.  Round trip through a fresh model."""
    import clc_b200.models as M
    src = M.CLC(N=64)
    src.update()
    sd = src.state_dict()
    assert sd["entropy_bottleneck._quantized_cdf"].numel() > 0
    assert sd["gaussian_conditional._quantized_cdf"].numel() > 0
    dst = M.CLC(N=64)
    dst.load_state_dict(sd)
    for k in ("entropy_bottleneck._quantized_cdf", "entropy_bottleneck._offset", "entropy_bottleneck._cdf_length",
              "gaussian_conditional._quantized_cdf", "gaussian_conditional._cdf_length",
              "gaussian_conditional.scale_table"):
        assert torch.equal(dst.state_dict()[k], sd[k]), k
