"""Authorities for the EntropyBottleneck / LowerBound arithmetic that do NOT depend on the in-tree CompressAI shim
(CompressAI is an un-vendored, un-pinned dependency of the reference: README.md:41,:70; SURVEY 8c "parity
unpinned").  Three independent checks, applied to the oracle (CPU) and to the CUDA kernels (GPU):

 1. closed form: with zero biases and zero factors the factorised density's cumulative logits are LINEAR,
    logits(x) = x * prod_i [f_in / (scale * f_out)] = x / init_scale, because every layer's weight is
    softplus(log(expm1(1/(scale f_out)))) = 1/(scale f_out) (the published initialisation), so
    likelihood(x) = sigmoid((x + .5)/10) - sigmoid((x - .5)/10) exactly;
 2. fp64 evaluation of the published algorithm (softplus-matrices, tanh gates, sign-stabilised sigmoid
    difference) written out independently below -- the fp32 results must be within 1e-4 of fp64, not merely
    of another fp32 implementation;
 3. the known constants: target = +-log(2/1e-9 - 1), init matrices, quantiles, LowerBound's gradient gate
    (gradient passes iff x >= bound or grad < 0).
"""
import math

import numpy as np
import pytest
import torch

LIK_RTOL = 1e-4


def _fp64_likelihood(x, state, bound=1e-9):
    """Independent float64 restatement of the factorised prior (filters (3,3,3,3)): numpy, per element."""
    M = [state[f"_matrix{i}"].double().numpy() for i in range(5)]
    b = [state[f"_bias{i}"].double().numpy() for i in range(5)]
    f = [state[f"_factor{i}"].double().numpy() for i in range(4)]
    xs = x.double().numpy()                                   # [B, C, ...]
    B, C = xs.shape[:2]
    flat = xs.reshape(B, C, -1)
    out = np.empty_like(flat)

    def logits(c, v):                                         # v: [n] values of channel c
        h = v[None, :]                                        # [1, n]
        for i in range(5):
            W = np.log1p(np.exp(M[i][c]))                     # softplus, [f_out, f_in]
            h = W @ h + b[i][c]
            if i < 4:
                h = h + np.tanh(f[i][c]) * np.tanh(h)
        return h[0]

    for c in range(C):
        v = flat[:, c, :].reshape(-1)
        lo, up = logits(c, v - 0.5), logits(c, v + 0.5)
        s = -np.sign(lo + up)
        lik = np.abs(1 / (1 + np.exp(-s * up)) - 1 / (1 + np.exp(-s * lo)))
        out[:, c, :] = np.maximum(lik, bound).reshape(B, -1)
    return torch.from_numpy(out.reshape(xs.shape))


def _state(C, seed, perturb):
    """Published initialisation (+ optional perturbation of every parameter)."""
    g = torch.Generator().manual_seed(seed)
    filters = (1, 3, 3, 3, 3, 1)
    scale = 10 ** (1 / 5)
    st = {}
    for i in range(5):
        init = math.log(math.expm1(1 / scale / filters[i + 1]))
        st[f"_matrix{i}"] = torch.full((C, filters[i + 1], filters[i]), init)
        st[f"_bias{i}"] = torch.rand(C, filters[i + 1], 1, generator=g) - 0.5
        if i < 4:
            st[f"_factor{i}"] = torch.zeros(C, filters[i + 1], 1)
    st["quantiles"] = torch.tensor([-10.0, 0.0, 10.0]).repeat(C, 1, 1)
    if perturb:
        for k in st:
            if k != "quantiles":
                st[k] = st[k] + perturb * torch.randn(st[k].shape, generator=g)
        st["quantiles"][:, 0, 1] = torch.randn(C, generator=g)
    return st


def _zero_bias(st):
    for i in range(5):
        st[f"_bias{i}"].zero_()
    return st


def _closed_form(x):
    xd = x.double()
    return torch.sigmoid((xd + 0.5) / 10) - torch.sigmoid((xd - 0.5) / 10)


def _relerr(a, ref, where=None):
    ref = ref.double()
    big = ref > 1e-9
    if where is not None:
        big = big & where
    return ((a.double() - ref).abs() / ref)[big].max().item()


def _check_closed_form(lik, q):
    """lik vs the closed form, except where lower + upper == 0 exactly (q == 0 with zero biases): the published
    sign-stabilised form uses sign = -sign(lower + upper) = 0 there, the difference of sigmoids collapses to 0
    and the likelihood is the 1e-9 floor -- a quirk of the algorithm both sides must reproduce."""
    assert _relerr(lik, _closed_form(q), q != 0) < LIK_RTOL
    assert torch.equal(lik[q == 0].double(), torch.full_like(lik[q == 0], 1e-9).double())


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_constants_match_published_closed_forms():
    from oracle import clc_oracle as O
    eb = O.EntropyBottleneck(4)
    t = math.log(2 / 1e-9 - 1)
    assert torch.allclose(eb.target, torch.tensor([-t, 0.0, t]))
    assert torch.equal(eb.quantiles.data[0, 0], torch.tensor([-10.0, 0.0, 10.0]))
    scale = 10 ** (1 / 5)
    for i, fo in enumerate((3, 3, 3, 3, 1)):
        m = getattr(eb, f"_matrix{i}")
        assert torch.allclose(m, torch.full_like(m, math.log(math.expm1(1 / scale / fo))))
        assert torch.allclose(torch.nn.functional.softplus(m), torch.full_like(m, 1 / (scale * fo)))
    for i in range(4):
        assert torch.count_nonzero(getattr(eb, f"_factor{i}")) == 0
    assert eb._get_medians().requires_grad is False


def test_oracle_lower_bound_gradient_gate():
    """compressai.ops.LowerBound: forward max(x, b); backward passes iff x >= b or grad < 0."""
    from oracle import enable_shim
    enable_shim()
    from compressai.ops import LowerBound
    lb = LowerBound(0.11)
    x = torch.tensor([0.05, 0.05, 0.11, 0.5], requires_grad=True)
    g = torch.tensor([1.0, -1.0, 1.0, 1.0])
    y = lb(x)
    assert torch.equal(y.detach(), torch.tensor([0.11, 0.11, 0.11, 0.5]))
    y.backward(g)
    assert torch.equal(x.grad, torch.tensor([0.0, -1.0, 1.0, 1.0]))


def test_oracle_eb_equals_closed_form_and_fp64():
    from oracle import clc_oracle as O
    from oracle import latent_path_oracle as LO
    g = torch.Generator().manual_seed(1)
    z = 6 * torch.randn(3, 8, 5, 7, generator=g)
    st = _zero_bias(_state(8, 2, 0.0))
    _, lik, z_hat = O.eb_forward(LO.make_eb(st), z)
    assert torch.equal(z_hat, torch.round(z))                            # medians 0
    _check_closed_form(lik.detach(), torch.round(z))
    for perturb in (0.0, 0.1, 0.3):
        st = _state(8, 3, perturb)
        _, lik, z_hat = O.eb_forward(LO.make_eb(st), z)
        med = st["quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        q = torch.round(z - med) + med
        assert torch.equal(z_hat, q)
        assert _relerr(lik, _fp64_likelihood(q, st)) < LIK_RTOL, perturb
        noise = torch.rand(z.shape, generator=g) - 0.5
        _, lik_n, _ = O.eb_forward(LO.make_eb(st), z, noise=noise)
        assert _relerr(lik_n, _fp64_likelihood(z + noise, st)) < LIK_RTOL, perturb


# ------------------------------------------------------------------------------------------------ GPU: the kernels
@pytest.mark.gpu
def test_cuda_eb_equals_closed_form_and_fp64():
    import clc_b200
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    z = 6 * torch.randn(3, 8, 5, 7, generator=g)

    def run(st, noise=None):
        eb = clc_b200.EntropyBottleneck(8).to(d)
        with torch.no_grad():
            for k, v in st.items():
                getattr(eb, k).copy_(v)
        eb.train(noise is not None)
        out, lik, z_hat = eb(z.to(d), noise=None if noise is None else noise.to(d), ste=True)
        return lik.cpu(), z_hat.cpu()

    st = _zero_bias(_state(8, 2, 0.0))
    lik, z_hat = run(st)
    assert torch.equal(z_hat, torch.round(z))
    _check_closed_form(lik, torch.round(z))                                 # closed form x / init_scale
    for perturb in (0.0, 0.1, 0.3):
        st = _state(8, 3, perturb)
        lik, z_hat = run(st)
        med = st["quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        q = torch.round(z - med) + med
        assert torch.equal(z_hat, q)                                       # symbols bit-exact
        assert _relerr(lik, _fp64_likelihood(q, st)) < LIK_RTOL, perturb    # vs FLOAT64, not vs the shim
        noise = torch.rand(z.shape, generator=g) - 0.5
        lik_n, _ = run(st, noise)
        assert _relerr(lik_n, _fp64_likelihood(z + noise, st)) < LIK_RTOL, perturb


@pytest.mark.gpu
def test_cuda_lower_bound_gates():
    """Both LowerBound gates of GaussianConditional (scale floor 0.11, likelihood floor 1e-9) in the analytic
    backward: gradient passes iff x >= bound or grad < 0."""
    import clc_b200
    d = torch.device("cuda:0")
    gc = clc_b200.GaussianConditional(None).to(d).eval()
    y = torch.zeros(1, 4, device=d)
    mu = torch.zeros(1, 4, device=d)
    sc = torch.tensor([[0.05, 0.05, 0.5, 0.5]], device=d, requires_grad=True)
    _, lik = gc(y, sc, mu)
    # d lik / d scale < 0 at v = 0 (a wider Gaussian puts less mass in the centre bin)
    lik.backward(torch.tensor([[1.0, -1.0, 1.0, -1.0]], device=d))
    g = sc.grad.cpu()[0]
    assert g[0].item() < 0.0            # upstream +1 -> gradient to the clamped scale is negative -> passes
    assert g[1].item() == 0.0           # upstream -1 -> positive gradient on a clamped scale -> blocked
    assert g[2].item() < 0.0 and g[3].item() > 0.0
    # likelihood floor: y far in the tail -> lik clamps to 1e-9; positive upstream gradient is blocked
    y2 = torch.full((1, 2), 40.0, device=d, requires_grad=True)
    sc2 = torch.ones(1, 2, device=d)
    gct = clc_b200.GaussianConditional(None).to(d).train()
    _, lik2 = gct(y2, sc2, torch.zeros(1, 2, device=d), noise=torch.zeros(1, 2, device=d))
    assert torch.equal(lik2.detach().cpu(), torch.full((1, 2), 1e-9))
    lik2.backward(torch.tensor([[1.0, -1.0]], device=d))
    assert y2.grad[0, 0].item() == 0.0
