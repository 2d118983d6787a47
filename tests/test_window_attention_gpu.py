"""Fused Swin window attention (clc_window_attention_fwd) against the module's own torch path -- the
restatement of the reference's WMSA.forward (CLC_run.py:142-193: roll, window partition, q k^T * scale + relative
position bias, shifted-window mask, softmax, . v, window reverse, roll back)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_matmul():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("typ", ["W", "SW"])
@pytest.mark.parametrize("geom", [(2, 16, 24, 64, 16), (1, 8, 8, 32, 8), (1, 32, 8, 128, 32), (3, 8, 40, 64, 32)])
def test_fused_window_attention_equals_torch_path(typ, geom):
    from clc_b200.layers import WMSA
    B, H, W, C, hd = geom
    torch.manual_seed(sum(geom))
    m = WMSA(C, C, hd, 8, typ).cuda().eval()
    with torch.no_grad():
        m.relative_position_params.normal_(0.0, 1.0)       # a bias that matters (the init is 0.02)
        x = torch.randn(B, H, W, C, device="cuda")
        assert m._fused_ok(x)
        got = m(x)
        m.fused_attention = False
        want = m(x)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-6), (got - want).abs().max().item()


def test_fused_path_only_without_grad():
    from clc_b200.layers import WMSA
    m = WMSA(32, 32, 16, 8, "SW").cuda()
    x = torch.randn(1, 8, 8, 32, device="cuda", requires_grad=True)
    assert not m._fused_ok(x)                               # autograd on: the differentiable torch path
    m(x).sum().backward()
    assert x.grad is not None and m.relative_position_params.grad is not None
