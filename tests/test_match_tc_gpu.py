"""GPU parity of the tcgen05 match path (clc_match_topk_tc): raw bf16 GEMM accumulators vs a
float64 evaluation of the same bf16-rounded operands, final top-k indices bit-exact vs the fp32
path and vs the CPU oracle, values to fp32 round-off, every patch certified."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _inputs(NQ, R, Cc, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(NQ, Cc, h, w, generator=g)
    refs = 0.5 * y.unsqueeze(1) + torch.randn(NQ, R, Cc, h, w, generator=g)
    return y, refs


def _run_debug(y, refs, p, k, gauss):
    """Calls the bring-up hook: returns (val, idx, xy) with xy the raw accumulators [NP, P, h*w]."""
    from clc_b200 import _lib
    h_ = _lib.lib()
    fn = h_.clc_debug_match_tc_xy
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 8 + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]
    NQ, R, Cc, h, w = refs.shape
    d = _dev()
    yq = y.to(d).contiguous()
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    P = (h // p) * (w // p)
    val = torch.empty(NQ * R, P, k, device=d)
    idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
    xy = torch.zeros(NQ * R, P, h * w, device=d)
    nb = h_.clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
    assert nb > 0
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    rc = fn(yq.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, int(gauss), val.data_ptr(), idx.data_ptr(),
            xy.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, (rc, h_.clc_last_cuda_error())
    torch.cuda.synchronize()
    return val, idx, xy


def _xy_ref(y, refs, p):
    """float64 correlation of the bf16-rounded operands at every in-range linear origin."""
    NQ, R, Cc, h, w = refs.shape
    yb = y.bfloat16().double()
    rb = refs.bfloat16().double()
    q = yb.reshape(NQ, Cc, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5).reshape(NQ, -1, Cc, p, p)  # [NQ,P,C,p,p]
    out = torch.zeros(NQ, R, q.shape[1], h - p + 1, w - p + 1, dtype=torch.float64)
    for n in range(NQ):
        for r in range(R):
            out[n, r] = torch.nn.functional.conv2d(rb[n, r:r + 1], q[n])[0]
    return out


@pytest.mark.parametrize("geom", [
    # NQ, R, C, h, w, p
    (1, 1, 64, 8, 8, 4),        # single chunk, single tile
    (2, 3, 320, 16, 16, 4),     # cfg2 latent: several narrow tiles per problem
    (1, 2, 320, 32, 48, 4),     # cfg3 latent
    (1, 1, 128, 12, 20, 2),     # 2x2 patches, P = 60
])
def test_tc_gemm_accumulators(geom):
    NQ, R, Cc, h, w, p = geom
    y, refs = _inputs(NQ, R, Cc, h, w, seed=sum(geom))
    _, _, xy = _run_debug(y, refs, p, min(4, (h - p + 1) * (w - p + 1)), gauss=True)
    ref = _xy_ref(y, refs, p)                                     # [NQ,R,P,ch,cw]
    P = ref.shape[2]
    got = xy.cpu().double().reshape(NQ, R, P, h, w)[:, :, :, :h - p + 1, :w - p + 1]
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err < 2e-4 * scale + 1e-3, f"bf16 GEMM accumulators off: max err {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("geom", [
    (2, 3, 320, 16, 16, 4, 4, True),
    (2, 1, 320, 16, 16, 4, 1, False),
    (1, 2, 320, 32, 48, 4, 4, True),
    (1, 3, 64, 8, 12, 4, 2, True),
    (1, 1, 128, 12, 20, 2, 8, True),
    (3, 2, 192, 20, 16, 4, 3, False),
])
def test_tc_topk_equals_fp32_mode_and_oracle(geom):
    import clc_b200
    from oracle import clc_oracle as O
    NQ, R, Cc, h, w, p, k, gauss = geom
    y, refs = _inputs(NQ, R, Cc, h, w, seed=7 + sum(geom[:6]))
    d = _dev()
    cnt = torch.zeros(1, dtype=torch.int32, device=d)
    from clc_b200 import _lib
    from clc_b200.ops import _stream
    yq = y.to(d)
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    P = (h // p) * (w // p)
    val = torch.empty(NQ * R, P, k, device=d)
    idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
    nb = _lib.lib().clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    _lib.call("clc_match_topk_tc", yq.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, int(gauss),
              val.data_ptr(), idx.data_ptr(), cnt.data_ptr(), 0.0, None, None, ws.data_ptr(), ws.numel(), _stream())
    v32, i32, _ = clc_b200.match_topk(yq, refs.to(d), p, p, k, gaussian_mask=gauss, mode="fp32")
    assert torch.equal(idx.view(NQ, R, P, k), i32), "tc-mode indices differ from fp32 mode"
    assert torch.allclose(val.view(NQ, R, P, k), v32, atol=2e-6, rtol=0)
    # public API + oracle
    vt, it, _ = clc_b200.match_topk(yq, refs.to(d), p, p, k, gaussian_mask=gauss, mode="tc")
    assert torch.equal(it, i32)
    mask = O.gaussian_masks(h, w, p, p) if gauss else None
    for rr in range(R):
        _, val_o, idx_o = O.si_finder(y, refs[:, rr], p, p, refs[:, rr], k, 15.0, mask=mask, return_index=True)
        assert torch.equal(it[:, rr].cpu().long(), idx_o), "tc-mode indices differ from the oracle"
        assert torch.allclose(vt[:, rr].cpu(), val_o, atol=3e-6)
    assert cnt.item() == 0, f"{cnt.item()} patches could not be certified"


def test_tc_unsupported_shapes_fail_loudly():
    from clc_b200 import _lib
    h_ = _lib.lib()
    assert h_.clc_match_topk_tc_workspace_bytes(1, 1, 100, 8, 8, 4, 4, 4) == 0     # C not a multiple of 64
    assert h_.clc_match_topk_tc_workspace_bytes(1, 1, 64, 8, 8, 4, 4, 9) == 0      # k > 8
    d = _dev()
    t = torch.zeros(1, 100, 8, 8, device=d)
    o = torch.zeros(1, 4, 4, device=d)
    with pytest.raises(RuntimeError, match="unsupported"):
        _lib.call("clc_match_topk_tc", t.data_ptr(), t.data_ptr(), 1, 1, 100, 8, 8, 4, 4, 4, 0, o.data_ptr(),
                  o.data_ptr(), None, 0.0, None, None, o.data_ptr(), 16, None)
