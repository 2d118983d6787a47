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
    h_ = _lib.debug_lib()          # bring-up build (-DCLC_DEBUG_ABI) of the same sources
    fn = h_.clc_debug_match_tc_xy
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
    (1, 2, 64, 12, 128, 4),     # wide latent (CLIC-shaped rows): large halo, several tiles per problem
    (1, 1, 64, 80, 128, 4),     # cfg4 latent geometry: two-unit tiles, 32-channel (64-byte swizzle) K chunks
    (1, 2, 128, 24, 128, 4),    # unit ranges that end inside a group: mixed one- and two-unit tiles
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
    (1, 2, 64, 8, 80, 4, 4, True),       # wide rows: three column segments per query block row (32+32+16)
    (1, 1, 64, 9, 36, 3, 2, True),       # 3x3 patches: 30-column segments (scalar staging path), 30 + 6
    (1, 2, 64, 12, 128, 4, 4, True),     # wide latent: many candidate lists per patch (2 per tile)
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
    assert torch.equal(vt.reshape(val.shape), val), "public API and raw C call must agree bit for bit"
    mask = O.gaussian_masks(h, w, p, p) if gauss else None
    for rr in range(R):
        _, val_o, idx_o = O.si_finder(y, refs[:, rr], p, p, refs[:, rr], k, 15.0, mask=mask, return_index=True)
        assert torch.equal(it[:, rr].cpu().long(), idx_o), "tc-mode indices differ from the oracle"
        assert torch.allclose(vt[:, rr].cpu(), val_o, atol=3e-6)
    assert cnt.item() == 0, f"{cnt.item()} patches could not be certified"


def test_tc_cfg4_indices_equal_fp32_mode_and_oracle():
    """BASELINE configs[3] geometry (1 x 3 refs x 320 ch x 80 x 128 latent, P = 640, L = 9625): SURVEY 7.3-1
    measured that this is where reduced-precision screening flips indices.  The tensor-core path must give
    idx_tc == idx_fp32 == oracle SI_Finder (reference lines Patch_Matching.py:181-185, :224), values to fp32
    round-off and every patch certified."""
    import clc_b200
    from oracle import clc_oracle as O
    NQ, R, Cc, h, w, p, k = 1, 3, 320, 80, 128, 4, 4
    y, refs = _inputs(NQ, R, Cc, h, w, seed=404)
    d = _dev()
    yq = y.to(d)
    cnt = torch.full((1,), 7, dtype=torch.int32, device=d)           # the call resets it
    from clc_b200 import _lib
    from clc_b200.ops import _stream
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    P = (h // p) * (w // p)
    val = torch.empty(NQ * R, P, k, device=d)
    idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
    nb = _lib.lib().clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    _lib.call("clc_match_topk_tc", yq.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, 1,
              val.data_ptr(), idx.data_ptr(), cnt.data_ptr(), 0.0, None, None, ws.data_ptr(), ws.numel(), _stream())
    v32, i32, _ = clc_b200.match_topk(yq, refs.to(d), p, p, k, gaussian_mask=True, mode="fp32")
    assert torch.equal(idx.view(NQ, R, P, k), i32), "tc-mode indices differ from fp32 mode at cfg4"
    assert torch.allclose(val.view(NQ, R, P, k), v32, atol=3e-6, rtol=0)
    assert cnt.item() == 0, f"{cnt.item()} patches could not be certified"
    mask = O.gaussian_masks(h, w, p, p)
    for rr in range(R):
        _, val_o, idx_o = O.si_finder(y, refs[:, rr], p, p, refs[:, rr], k, 15.0, mask=mask, return_index=True)
        assert torch.equal(idx.view(NQ, R, P, k)[:, rr].cpu().long(), idx_o), "tc-mode indices differ from the oracle"
        assert torch.allclose(val.view(NQ, R, P, k)[:, rr].cpu(), val_o, atol=3e-6)


def test_tc_many_exact_ties_take_the_exact_selection_path():
    """A reference that is periodic with the patch size makes hundreds of windows IDENTICAL, so their screened
    scores tie exactly and the threshold scan of select_kernel overflows its list: the exact arg-max fallback
    must still return the fp32 path's indices (ties -> lowest window index, as torch.topk does on the
    reference's map) -- nothing can be certified, and the counter must say so."""
    import clc_b200
    NQ, R, Cc, h, w, p, k = 1, 1, 64, 48, 96, 4, 4
    g = torch.Generator().manual_seed(5)
    y = torch.randn(NQ, Cc, h, w, generator=g)
    tile = torch.randn(NQ, R, Cc, p, p, generator=g)
    refs = tile.repeat(1, 1, 1, h // p, w // p)                   # period p in both directions
    d = _dev()
    yq = y.to(d)
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    cnt = torch.zeros(1, dtype=torch.int32, device=d)
    from clc_b200 import _lib
    from clc_b200.ops import _stream
    P = (h // p) * (w // p)
    val = torch.empty(NQ * R, P, k, device=d)
    idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
    nb = _lib.lib().clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    _lib.call("clc_match_topk_tc", yq.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, 0,
              val.data_ptr(), idx.data_ptr(), cnt.data_ptr(), 0.0, None, None, ws.data_ptr(), ws.numel(), _stream())
    v32, i32, _ = clc_b200.match_topk(yq, refs.to(d), p, p, k, gaussian_mask=False, mode="fp32")
    # identical windows: the 16 distinct shifts of the tile repeat (h/p-1)*(w/p-1)+ times each
    assert torch.equal(idx.view(NQ, R, P, k), i32)
    assert torch.allclose(val.view(NQ, R, P, k), v32, atol=3e-6, rtol=0)
    assert cnt.item() > 0           # exact ties cannot be certified against the screening error


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


def _dcall(name, *args):
    """Status-checked call into the bring-up library (stage masks / experiment bits live only there)."""
    from clc_b200 import _lib
    h_ = _lib.debug_lib()
    rc = getattr(h_, name)(*args)
    assert rc == 0, (name, rc, h_.clc_last_cuda_error())


def _tc_call(yq, r, R, p, k, gauss, want_aligned, debug_build=False):
    from clc_b200 import _lib
    from clc_b200.ops import _stream
    NP, Cc, h, w = r.shape
    P = (h // p) * (w // p)
    d = r.device
    val = torch.empty(NP, P, k, device=d)
    idx = torch.empty(NP, P, k, dtype=torch.int32, device=d)
    aligned = torch.empty_like(r) if want_aligned else None
    weights = torch.empty(NP, P, k, device=d) if want_aligned else None
    nb = _lib.lib().clc_match_topk_tc_workspace_bytes(NP, R, Cc, h, w, p, p, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=d)
    if debug_build:
        nb = _lib.debug_lib().clc_match_topk_tc_workspace_bytes(NP, R, Cc, h, w, p, p, k)
        ws = torch.empty(nb, dtype=torch.uint8, device=d)
    (_dcall if debug_build else _lib.call)(
        "clc_match_topk_tc", yq.data_ptr(), r.data_ptr(), NP, R, Cc, h, w, p, p, k, int(gauss), val.data_ptr(),
        idx.data_ptr(), None, 15.0, _lib.ptr(aligned), _lib.ptr(weights), ws.data_ptr(), ws.numel(), _stream())
    return val, idx, aligned, weights, ws


@pytest.mark.parametrize("geom", [(2, 3, 320, 16, 16, 4, 4), (1, 2, 128, 32, 48, 4, 3), (2, 1, 64, 8, 12, 4, 2)])
def test_tc_fused_gather_equals_gather_kernel(geom):
    """The gather/blend fused into the re-scoring kernel == clc_gather_blend_fwd on the same (idx, val)."""
    from clc_b200 import _lib
    from clc_b200.ops import _stream
    NQ, R, Cc, h, w, p, k = geom
    y, refs = _inputs(NQ, R, Cc, h, w, seed=11 + sum(geom))
    d = _dev()
    yq = y.to(d)
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    val, idx, aligned, weights, _ = _tc_call(yq, r, R, p, k, True, True)
    val2, idx2, _, _, _ = _tc_call(yq, r, R, p, k, True, False)
    assert torch.equal(idx, idx2) and torch.equal(val, val2)
    out = torch.empty_like(r)
    w2 = torch.empty_like(weights)
    _lib.call("clc_gather_blend_fwd", r.data_ptr(), idx.data_ptr(), val.data_ptr(), 15.0, out.data_ptr(), w2.data_ptr(),
              NQ * R, Cc, h, w, p, p, w - p + 1, k, 0, _stream())
    assert torch.equal(weights, w2)
    assert torch.allclose(aligned, out, atol=1e-6, rtol=0)


def test_stacked_and_general_gemm_kernels_agree():
    """Small latents take the stacked-shift tcgen05 kernel; experiment bit 16 of the bring-up build forces
    the general one."""
    from clc_b200 import _lib
    NQ, R, Cc, h, w, p, k = 3, 2, 320, 16, 16, 4, 4
    y, refs = _inputs(NQ, R, Cc, h, w, seed=123)
    d = _dev()
    yq = y.to(d)
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    v1, i1, a1, _, _ = _tc_call(yq, r, R, p, k, True, True)
    H = _lib.debug_lib()
    H.clc_debug_set_stage_mask(0xff | (16 << 8))
    try:
        v2, i2, a2, _, _ = _tc_call(yq, r, R, p, k, True, True, debug_build=True)
        torch.cuda.synchronize()
    finally:
        H.clc_debug_set_stage_mask(0xff)
    assert torch.equal(i1, i2), "stacked / general screening must select the same windows"
    assert torch.equal(v1, v2) and torch.equal(a1, a2)      # final values come from the same fp32 re-scoring


@pytest.mark.parametrize("geom", [(2, 3, 320, 16, 16, 4), (1, 2, 128, 8, 12, 3), (2, 2, 64, 16, 32, 4)])
def test_match_bwd_variants_agree(geom):
    """clc_match_bwd: thread-owns-items kernel (default) == register kernel == re-read kernel ==
    workspace-free kernel; overwrite vs accumulate; supplied vs internally made channels-last copy;
    pre-zeroed workspace."""
    import ctypes as C
    from clc_b200 import _lib
    from clc_b200.matching import _patch_view_from_image
    from clc_b200.ops import _stream
    NQ, R, Cc, h, w, k = geom
    p = 4
    y, refs = _inputs(NQ, R, Cc, h, w, seed=31 + sum(geom))
    d = _dev()
    yq = y.to(d)
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    val, idx, aligned, weights, ws_f = _tc_call(yq, r, R, p, k, False, True)
    NP, P = r.shape[0], idx.shape[1]
    g_out = torch.randn(r.shape, generator=torch.Generator().manual_seed(2)).to(d)
    view = _patch_view_from_image(yq, p, p, R)
    H = _lib.debug_lib()
    r_cl = H.clc_match_topk_tc_ref_cl(ws_f.data_ptr(), NP, R, Cc, h, w, p, p, k)
    nb = H.clc_match_bwd_workspace_bytes(NP, Cc, h, w)

    def run(flags, use_ws=True, use_cl=True, dbg=0, pre=None, zero_ws=False):
        ws = torch.empty(nb, dtype=torch.uint8, device=d)
        g_r = torch.zeros_like(r) if pre is None else pre.clone()
        g_q = torch.zeros_like(yq)
        g_val = torch.empty_like(val)
        if zero_ws:
            _dcall("clc_match_bwd_zero_workspace", ws.data_ptr(), ws.numel(), NP, Cc, h, w, _stream())
        H.clc_debug_set_stage_mask(0xff | (dbg << 8))
        try:
            _dcall("clc_match_bwd", C.byref(view), r.data_ptr(), r_cl if use_cl else None, None, idx.data_ptr(),
                      weights.data_ptr(), 15.0, g_out.data_ptr(), g_r.data_ptr(), g_q.data_ptr(), g_val.data_ptr(),
                      NP, P, Cc, p, p, h, w, k, flags, ws.data_ptr() if use_ws else None, ws.numel() if use_ws else 0,
                      _stream())
        finally:
            H.clc_debug_set_stage_mask(0xff)
        torch.cuda.synchronize()
        return g_r, g_q, g_val

    base = run(1)
    scale = base[0].abs().max().item()
    for name, other in (("register kernel", run(1, dbg=8)), ("re-read kernel", run(1, dbg=8 | 4)),
                        ("workspace-free kernel", run(1, use_ws=False)), ("own channels-last copy", run(1, use_cl=False)),
                        ("pre-zeroed workspace", run(3, zero_ws=True))):
        for a, b in zip(base, other):
            # different summation orders (block reductions, atomics): fp32 round-off relative to each tensor's scale
            assert (a - b).abs().max().item() <= 1e-4 * max(a.abs().max().item(), 1e-3), name
    pre = torch.randn(r.shape, generator=torch.Generator().manual_seed(3)).to(d)
    acc = run(0, pre=pre)
    assert (acc[0] - (base[0] + pre)).abs().max().item() <= 2e-5 * max(scale, 1.0)
    # overwrite mode ignores what g_r held before
    ow = run(1, pre=pre)
    assert (ow[0] - base[0]).abs().max().item() <= 2e-5 * max(scale, 1.0)


def test_degenerate_inputs_keep_indices_in_bounds():
    """Zero-variance patches / windows have NaN correlation (0/0) in the reference; torch.topk then
    returns NaN values at valid positions.  The tc path must do the same: NaN values, indices that are
    valid windows (the gather and the backward index memory with them), no fault -- e.g. the warm-up
    replay of a freshly allocated, zero-filled LatentPath."""
    from clc_b200.latent_path import LatentPath
    d = _dev()
    lp = LatentPath(2, 256, 256, n_refs=2, train=True, match_mode="tc", device="cuda:0")    # inputs all zero
    lp.step()
    torch.cuda.synchronize()
    L = (lp.h - 3) * (lp.w - 3)
    assert int(lp.idx.min()) >= 0 and int(lp.idx.max()) < L
    assert torch.isnan(lp.val).all()
    # one constant query patch among normal ones: only that patch degenerates
    NQ, R, Cc, h, w, p, k = 1, 2, 64, 16, 16, 4, 4
    y, refs = _inputs(NQ, R, Cc, h, w, seed=77)
    y[:, :, 4:8, 8:12] = 0.25                               # patch (1, 2) -> index 1 * 4 + 2 = 6
    r = refs.to(d).reshape(NQ * R, Cc, h, w).contiguous()
    val, idx, aligned, _, _ = _tc_call(y.to(d), r, R, p, k, True, True)
    torch.cuda.synchronize()
    Lw = (h - p + 1) * (w - p + 1)
    assert int(idx.min()) >= 0 and int(idx.max()) < Lw
    # the constant patch has (numerically) zero variance: its values are NaN / inf like the reference's 0/0 and
    # x/0, every other patch is untouched
    assert not torch.isfinite(val[:, 6]).any()
    assert torch.isfinite(val[:, :6]).all() and torch.isfinite(val[:, 7:]).all()


def test_public_api_falls_back_to_fp32_outside_the_tc_kernels_and_strict_mode():
    """The public functions route shapes the tensor-core kernels do not cover (C % 64 != 0) to the exact fp32
    path (with a warning) instead of raising; strict=True re-runs a tc call in fp32 mode when the certification
    counter is non-zero (exact ties cannot be certified)."""
    import clc_b200
    from clc_b200 import matching
    d = _dev()
    y, refs = _inputs(1, 2, 100, 8, 12, seed=3)
    with pytest.warns(UserWarning, match="outside the tensor-core"):
        v_t, i_t, _ = clc_b200.match_topk(y.to(d), refs.to(d), 4, 4, 2, mode="tc")
    v_f, i_f, _ = clc_b200.match_topk(y.to(d), refs.to(d), 4, 4, 2, mode="fp32")
    assert torch.equal(i_t, i_f) and torch.equal(v_t, v_f)
    a_t = clc_b200.match_and_gather(y.to(d), refs.to(d), 4, 4, 2, mode="tc")
    a_f = clc_b200.match_and_gather(y.to(d), refs.to(d), 4, 4, 2, mode="fp32")
    assert torch.equal(a_t, a_f)
    # periodic reference: exact ties -> uncertified -> strict mode answers with the fp32 path's result
    g = torch.Generator().manual_seed(5)
    yq = torch.randn(1, 64, 16, 32, generator=g)
    tile = torch.randn(1, 1, 64, 4, 4, generator=g)
    rp = tile.repeat(1, 1, 1, 4, 8)
    v_s, i_s, _ = clc_b200.match_topk(yq.to(d), rp.to(d), 4, 4, 4, gaussian_mask=False, mode="tc", strict=True)
    assert matching.last_uncertified(d) > 0
    v_e, i_e, _ = clc_b200.match_topk(yq.to(d), rp.to(d), 4, 4, 4, gaussian_mask=False, mode="fp32")
    assert torch.equal(i_s, i_e) and torch.equal(v_s, v_e)
