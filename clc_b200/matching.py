"""Reference matching: host-side mirror of models/Patch_Matching.py over the C ABI.

Same function names, argument meaning and return shapes as the reference
(`L2_or_pearson_corr` :854-910, `create_gaussian_masks` :779-807, `SI_Wraper` :218-240,
`SI_Finder_at_Decoder_Feature_Domain` :157-216) plus the fused fast path `match_topk` /
`match_and_gather` that never materialises the P x L correlation map.  Unlike the reference
nothing here hard-codes `.cuda()` on new tensors: outputs live on the inputs' device, which
must be CUDA (there is no CPU fallback).
"""
import ctypes as C

import torch

from ._lib import PatchView, call, lib, ptr
from .ops import _check, _stream


def _patch_view_from_patches(x):
    """x: [P, C, ph, pw] contiguous extracted patches (reference calling convention)."""
    P, Cc, ph, pw = x.shape
    v = PatchView()
    v.q = x.data_ptr()
    v.q_sn = 0
    v.q_spy = 0
    v.q_spx = Cc * ph * pw
    v.q_sc = ph * pw
    v.q_sy = pw
    v.npx = P
    v.q_repeat = 1  # q_sn = 0: every problem reads the same query set
    return v


def _patch_view_from_image(img, ph, pw, q_repeat=1):
    """img: [NQ, C, H, W] contiguous; patches addressed in place (no reshape/permute copy)."""
    NQ, Cc, H, W = img.shape
    v = PatchView()
    v.q = img.data_ptr()
    v.q_sn = Cc * H * W
    v.q_spy = ph * W
    v.q_spx = pw
    v.q_sc = H * W
    v.q_sy = W
    v.npx = W // pw
    v.q_repeat = int(q_repeat)
    return v


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ---- certification of the tensor-core path ---------------------------------------------------------------
# clc_match_topk_tc screens with a bf16 GEMM and re-scores the best candidates exactly; a patch whose margin
# is below the screening-error bound even after the second re-scoring round is counted as "uncertified" (its
# indices may differ from the exact fp32 ranking at a near-tie).  The counter of the LAST tc call on a device is
# kept here; `strict=True` on the public functions reads it (one host sync) and re-runs the call in fp32 mode.
_UNCERT = {}
_WARNED = set()


def _uncert_counter(device):
    idx = torch.device(device).index
    if idx not in _UNCERT:
        _UNCERT[idx] = torch.zeros(1, dtype=torch.int32, device=device)
    return _UNCERT[idx]


def last_uncertified(device=None):
    """Patches of the last tensor-core match call on `device` whose top-k could not be certified against the
    screening error (0 = indices provably equal the exact fp32 ranking).  Synchronises."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return int(_uncert_counter(dev).item())


def _tc_supported(NP, q_repeat, Cc, H, W, ph, pw, k):
    """Shapes the tensor-core path implements (C a multiple of 64, k <= 8, ...); others take the fp32 path."""
    ok = lib().clc_match_topk_tc_workspace_bytes(NP, q_repeat, Cc, H, W, ph, pw, k) > 0
    if not ok and (Cc, H, W, ph, pw, k) not in _WARNED:
        _WARNED.add((Cc, H, W, ph, pw, k))
        import warnings
        warnings.warn(f"clc_b200: match shape C={Cc} {H}x{W} patch {ph}x{pw} k={k} is outside the tensor-core "
                      "kernels; using the exact fp32 path")
    return ok


def _corr_raw(view, r, P, ph, pw, mask=None):
    NP, Cc, fh, fw = r.shape
    ch, cw = fh - ph + 1, fw - pw + 1
    corr = torch.empty((NP, P, ch, cw), dtype=torch.float32, device=r.device)
    nb = lib().clc_pearson_corr_workspace_bytes(NP, P, Cc, ph, pw, fh, fw)
    ws = _workspace(nb, r.device)
    call("clc_pearson_corr", C.byref(view), ptr(r), ptr(mask), ptr(corr), NP, P, Cc, ph, pw, fh, fw,
         ptr(ws), ws.numel(), _stream())
    return corr


def create_gaussian_masks(img_h, img_w, patch_h, patch_w, device=None):
    """[1, P, img_h-ph+1, img_w-pw+1] Gaussian masks (Patch_Matching.py:779-807), built on the device."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    P = (img_h * img_w) // (patch_h * patch_w)
    mask = torch.empty((1, P, img_h - patch_h + 1, img_w - patch_w + 1), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        call("clc_gaussian_mask", ptr(mask), img_h, img_w, patch_h, patch_w, _stream())
    mask._clc_gaussian = True  # lets the tc path rebuild it from separable tables
    return mask


def topk_rows(x, k):
    """(values, int32 indices) of the k largest entries of each row of x [R, L]; ties -> lowest index."""
    _check(x, "x")
    x = x.contiguous()
    R, L = x.shape
    val = torch.empty((R, k), dtype=torch.float32, device=x.device)
    idx = torch.empty((R, k), dtype=torch.int32, device=x.device)
    call("clc_topk_rows", ptr(x), R, L, k, ptr(val), ptr(idx), _stream())
    return val, idx


class _PearsonTopkFn(torch.autograd.Function):
    """(q_img, r) -> (val, idx): masked Pearson correlation + top-k, differentiable w.r.t. the
    selected values with the reference's autograd semantics (see clc_pearson_topk_bwd)."""

    @staticmethod
    def forward(ctx, q_img, r, mask, ph, pw, k, q_repeat, mode):
        NP, Cc, fh, fw = r.shape
        NQ, _, H, W = q_img.shape
        P = (H // ph) * (W // pw)
        view = _patch_view_from_image(q_img, ph, pw, q_repeat)
        if mode == "fp32":
            corr = _corr_raw(view, r, P, ph, pw, mask)
            L = corr.shape[2] * corr.shape[3]
            val, idx = topk_rows(corr.view(NP * P, L), k)
            val, idx = val.view(NP, P, k), idx.view(NP, P, k)
        elif mode == "tc" and not _tc_supported(NP, q_repeat, Cc, H, W, ph, pw, k) and (fh, fw) == (H, W):
            return _PearsonTopkFn.forward(ctx, q_img, r, mask, ph, pw, k, q_repeat, "fp32")
        elif mode == "tc":
            if (fh, fw) != (H, W):
                raise ValueError("tc mode needs query and reference latents of the same spatial size")
            val = torch.empty((NP, P, k), dtype=torch.float32, device=r.device)
            idx = torch.empty((NP, P, k), dtype=torch.int32, device=r.device)
            gauss = 0
            if mask is not None:
                if not getattr(mask, "_clc_gaussian", False):
                    raise ValueError("tc mode supports mask=None or the mask from create_gaussian_masks")
                gauss = 1
            nb = lib().clc_match_topk_tc_workspace_bytes(NP, q_repeat, Cc, H, W, ph, pw, k)
            ws = _workspace(nb, r.device)
            call("clc_match_topk_tc", ptr(q_img), ptr(r), NP, q_repeat, Cc, H, W, ph, pw, k, gauss,
                 ptr(val), ptr(idx), ptr(_uncert_counter(r.device)), 0.0, None, None, ptr(ws), ws.numel(), _stream())
        else:
            raise ValueError(f'Invalid match mode "{mode}" (expected "fp32" or "tc")')
        ctx.save_for_backward(q_img, r, mask, idx)
        ctx.geom = (ph, pw, k, q_repeat, P)
        ctx.mark_non_differentiable(idx)
        return val, idx

    @staticmethod
    def backward(ctx, g_val, _g_idx):
        q_img, r, mask, idx = ctx.saved_tensors
        ph, pw, k, q_repeat, P = ctx.geom
        NP, Cc, fh, fw = r.shape
        g_r = torch.zeros_like(r)
        g_q = torch.zeros_like(q_img) if ctx.needs_input_grad[0] else None
        view = _patch_view_from_image(q_img, ph, pw, q_repeat)
        call("clc_pearson_topk_bwd", C.byref(view), ptr(r), ptr(mask), ptr(idx), ptr(g_val.contiguous()),
             ptr(g_r), ptr(g_q), NP, P, Cc, ph, pw, fh, fw, k, _stream())
        return g_q, g_r, None, None, None, None, None, None


class _GatherBlendFn(torch.autograd.Function):
    """(feat, val, idx) -> blended / stacked reference (SI_Wraper :226-238)."""

    @staticmethod
    def forward(ctx, feat, val, idx, gh, gw, corr_w, temperature, is_stack):
        NP, Cc, fh, fw = feat.shape
        k = idx.shape[-1]
        out = torch.empty((NP, Cc * (k if is_stack else 1), fh, fw), dtype=torch.float32, device=feat.device)
        weights = None if is_stack else torch.empty(idx.shape, dtype=torch.float32, device=feat.device)
        call("clc_gather_blend_fwd", ptr(feat), ptr(idx), ptr(val), float(temperature), ptr(out),
             ptr(weights), NP, Cc, fh, fw, gh, gw, corr_w, k, int(is_stack), _stream())
        ctx.save_for_backward(feat, idx, weights)
        ctx.geom = (gh, gw, corr_w, float(temperature), bool(is_stack))
        return out

    @staticmethod
    def backward(ctx, g_out):
        feat, idx, weights = ctx.saved_tensors
        gh, gw, corr_w, temperature, is_stack = ctx.geom
        NP, Cc, fh, fw = feat.shape
        k = idx.shape[-1]
        g_feat = torch.zeros_like(feat)
        g_val = torch.empty(idx.shape, dtype=torch.float32, device=feat.device)
        call("clc_gather_blend_bwd", ptr(feat), ptr(idx), ptr(weights), temperature,
             ptr(g_out.contiguous()), ptr(g_feat), ptr(g_val), NP, Cc, fh, fw, gh, gw, corr_w, k,
             int(is_stack), _stream())
        return g_feat, g_val, None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
# Reference-named API
# ---------------------------------------------------------------------------------------------
def L2_or_pearson_corr(x, y, patch_h, patch_w, is_cpu=False):
    """Pearson correlation of each patch x[P,C,ph,pw] with every window of y[1,C,fh,fw]
    -> [1, P, fh-ph+1, fw-pw+1]   (Patch_Matching.py:854-910).  `is_cpu` (a memory-saving
    offload switch in the reference) is accepted and ignored: the map stays in HBM."""
    _check(x, "x")
    _check(y, "y")
    x = x.detach().contiguous()
    y = y.detach().contiguous()
    P, Cc, ph, pw = x.shape
    if (ph, pw) != (patch_h, patch_w) or y.shape[1] != Cc:
        raise ValueError("patch / feature shapes disagree")
    return _corr_raw(_patch_view_from_patches(x), y, P, ph, pw)


def SI_Wraper(cross_corr, patch_h, patch_w, patchs_num, y, k=1, temperature=15, is_stack=False):
    """top-k -> softmax(value*T) -> gather k patches -> weighted sum / stack -> reassembly
    (Patch_Matching.py:218-240).  cross_corr [1,P,ch,cw], y [1,C,fh,fw]."""
    _check(cross_corr, "cross_corr")
    _check(y, "y")
    _, P, ch, cw = cross_corr.shape
    if P != patchs_num:
        raise ValueError("patchs_num disagrees with cross_corr")
    corr2 = cross_corr.reshape(P, ch * cw)
    val, idx = _TopkValuesFn.apply(corr2, int(k))
    return _GatherBlendFn.apply(y.contiguous(), val.view(1, P, k), idx.view(1, P, k), patch_h, patch_w, cw,
                                temperature, is_stack)


class _TopkValuesFn(torch.autograd.Function):
    """torch.topk(dim=-1) replacement that keeps the value gradient (scatter to the map)."""

    @staticmethod
    def forward(ctx, x, k):
        val, idx = topk_rows(x, k)
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        ctx.mark_non_differentiable(idx)
        return val, idx

    @staticmethod
    def backward(ctx, g_val, _):
        (idx,) = ctx.saved_tensors
        g = torch.zeros(ctx.shape, dtype=g_val.dtype, device=g_val.device)
        g.scatter_(1, idx.long(), g_val)  # plumbing: routes k gradients per row back to the map
        return g, None


def SI_Finder_at_Decoder_Feature_Domain(x_decs, ys, patch_h, patch_w, y_decs, layer_names, args, mask=None,
                                        other_ys=None, is_img_patch_matching=False,
                                        is_pearson_corr_cpu=False, mode="fp32"):
    """Batched match + gather (Patch_Matching.py:157-216): for every image n, correlate the
    patches of x_decs[n] with y_decs[n], take the top-k positions (optionally masked) and gather
    from ys[n] (and from the coarser `other_ys` with the sub-sampled map, :198-208).

    Returns {layer_name: [N, C(*k if stack), fh_i, fw_i]}.  The whole batch runs in a handful of
    launches instead of the reference's per-image Python loop.  `args` needs num_k, temperature,
    is_stack, single_layer.  mode: "fp32" (exact, materialised map) or "tc" (tcgen05 screening +
    fp32 re-scoring; single scale only)."""
    if is_img_patch_matching:
        raise NotImplementedError("RGB-domain matching (KITTI statistics, :177-179) is outside the latent path")
    _check(x_decs, "x_decs")
    N, Cc, H, W = x_decs.shape
    _, _, fh, fw = ys.shape
    P = (H // patch_h) * (W // patch_w)
    k, T, is_stack = int(args.num_k), float(args.temperature), bool(args.is_stack)
    cw = fw - patch_w + 1
    m = None
    if mask is not None:
        m = mask.reshape(P, fh - patch_h + 1, cw).contiguous()
        if getattr(mask, "_clc_gaussian", False):
            m._clc_gaussian = True
    multi = (args.single_layer != 0 and args.single_layer != 4) or (args.single_layer == 0 and other_ys is not None)
    out = {}
    if not multi:
        val, idx = _PearsonTopkFn.apply(x_decs.contiguous(), y_decs.contiguous(), m, patch_h, patch_w, k, 1, mode)
        out[layer_names[0]] = _GatherBlendFn.apply(ys.contiguous(), val, idx, patch_h, patch_w, cw, T, is_stack)
        return out
    # Multi-scale: the reference re-runs top-k on the strided map cross_corr[:, :, ::s, ::s].
    # (The map itself is not differentiated on this path: gradients reach `ys` / `other_ys`
    # through the gather only.  Use the single-scale path for training through the values.)
    xq = x_decs.detach().contiguous()
    corr = _corr_raw(_patch_view_from_image(xq, patch_h, patch_w, 1), y_decs.detach().contiguous(), P,
                     patch_h, patch_w, m)
    L = corr.shape[2] * corr.shape[3]
    val, idx = _TopkValuesFn.apply(corr.view(N * P, L), k)
    out[layer_names[0]] = _GatherBlendFn.apply(ys.contiguous(), val.view(N, P, k), idx.view(N, P, k),
                                               patch_h, patch_w, cw, T, is_stack)
    if args.single_layer != 0:
        scales = [(3 - args.single_layer, str(4 - (3 - args.single_layer)))]
    else:
        scales = [(i, layer_names[i + 1]) for i in range(len(other_ys))]
    for i, name in scales:
        s = 2 ** (i + 1)
        oy = other_ys[i].contiguous()
        assert fh // oy.shape[2] == s and fw // oy.shape[3] == s
        sub = corr[:, :, ::s, ::s].contiguous()
        Ls = sub.shape[2] * sub.shape[3]
        v_i, i_i = _TopkValuesFn.apply(sub.view(N * P, Ls), k)
        out[name] = _GatherBlendFn.apply(oy, v_i.view(N, P, k), i_i.view(N, P, k), patch_h // s, patch_w // s,
                                         sub.shape[3], T, is_stack)
    return out


def match_topk(y, refs, patch_h=4, patch_w=4, k=4, gaussian_mask=True, mode="tc", strict=False):
    """Fused match for a batch of images against R references each.
    y [B,C,h,w]; refs: list of R tensors [B,C,h,w] or a stacked [B,R,C,h,w] tensor.
    Returns (val [B,R,P,k], idx int32 [B,R,P,k], refs_stacked [B*R,C,h,w]).
    mode="tc": indices equal the exact fp32 ranking whenever last_uncertified() == 0 (always on the shapes and
    data of the test-suite); strict=True checks that (one host sync) and re-runs in fp32 mode otherwise."""
    _check(y, "y")
    if isinstance(refs, (list, tuple)):
        refs = torch.stack(list(refs), dim=1)
    B, R, Cc, h, w = refs.shape
    r = refs.reshape(B * R, Cc, h, w).contiguous()
    mask = None
    if gaussian_mask:
        mask = _cached_mask(h, w, patch_h, patch_w, y.device)
    val, idx = _PearsonTopkFn.apply(y.contiguous(), r, mask, patch_h, patch_w, int(k), R, mode)
    if strict and mode == "tc" and last_uncertified(y.device) > 0:
        val, idx = _PearsonTopkFn.apply(y.contiguous(), r, mask, patch_h, patch_w, int(k), R, "fp32")
    P = val.shape[1]
    return val.view(B, R, P, k), idx.view(B, R, P, k), r


_MASKS = {}


def _cached_mask(h, w, ph, pw, device):
    key = (h, w, ph, pw, str(device))
    m = _MASKS.get(key)
    if m is None:
        m = create_gaussian_masks(h, w, ph, pw, device=device)[0].contiguous()
        m._clc_gaussian = True
        _MASKS[key] = m
    return m


class _MatchGatherFn(torch.autograd.Function):
    """(y, r) -> aligned references: masked Pearson top-k + softmax-weighted gather in the forward,
    ONE fused backward kernel (clc_match_bwd) for both -- the single-scale wiring where the
    gathered feature map is the matched reference itself."""

    @staticmethod
    def forward(ctx, q_img, r, mask, ph, pw, k, q_repeat, temperature, mode):
        NP, Cc, fh, fw = r.shape
        r_cl = ws = None
        if mode == "tc" and r.shape[2:] == q_img.shape[2:] and \
                not _tc_supported(NP, q_repeat, Cc, q_img.shape[2], q_img.shape[3], ph, pw, k):
            mode = "fp32"
        if mode == "tc":
            # fused: screening GEMM -> exact re-scoring + top-k + softmax + gather/blend in one C call
            NQ, _, H, W = q_img.shape
            if (fh, fw) != (H, W):
                raise ValueError("tc mode needs query and reference latents of the same spatial size")
            gauss = 0
            if mask is not None:
                if not getattr(mask, "_clc_gaussian", False):
                    raise ValueError("tc mode supports mask=None or the mask from create_gaussian_masks")
                gauss = 1
            P = (H // ph) * (W // pw)
            val = torch.empty((NP, P, k), dtype=torch.float32, device=r.device)
            idx = torch.empty((NP, P, k), dtype=torch.int32, device=r.device)
            weights = torch.empty_like(val)
            out = torch.empty_like(r)
            nb = lib().clc_match_topk_tc_workspace_bytes(NP, q_repeat, Cc, H, W, ph, pw, k)
            ws = _workspace(nb, r.device)
            call("clc_match_topk_tc", ptr(q_img), ptr(r), NP, q_repeat, Cc, H, W, ph, pw, k, gauss, ptr(val), ptr(idx),
                 ptr(_uncert_counter(r.device)), float(temperature), ptr(out), ptr(weights), ptr(ws), ws.numel(),
                 _stream())
            if any(ctx.needs_input_grad[:2]):
                # channels-last fp32 copy of r left in the workspace: the backward reuses it
                r_cl = lib().clc_match_topk_tc_ref_cl(ptr(ws), NP, q_repeat, Cc, H, W, ph, pw, k)
        else:
            with torch.no_grad():
                val, idx = _PearsonTopkFn.apply(q_img, r, mask, ph, pw, k, q_repeat, mode)
            out = torch.empty_like(r)
            weights = torch.empty(idx.shape, dtype=torch.float32, device=r.device)
            call("clc_gather_blend_fwd", ptr(r), ptr(idx), ptr(val), float(temperature), ptr(out), ptr(weights), NP,
                 Cc, fh, fw, ph, pw, fw - pw + 1, k, 0, _stream())
        ctx.save_for_backward(q_img, r, mask, idx, weights)
        ctx.geom = (ph, pw, k, q_repeat, float(temperature))
        ctx.r_cl, ctx.fwd_ws = r_cl, (ws if r_cl else None)   # keeps the workspace alive for the backward
        ctx.mark_non_differentiable(idx)
        return out, val, idx

    @staticmethod
    def backward(ctx, g_out, _g_val, _g_idx):
        q_img, r, mask, idx, weights = ctx.saved_tensors
        ph, pw, k, q_repeat, temperature = ctx.geom
        NP, Cc, fh, fw = r.shape
        P = idx.shape[1]
        g_r = torch.empty_like(r)                 # written, not accumulated (CLC_MATCH_BWD_OVERWRITE_G_R)
        g_q = torch.zeros_like(q_img) if ctx.needs_input_grad[0] else None
        view = _patch_view_from_image(q_img, ph, pw, q_repeat)
        ws = _workspace(lib().clc_match_bwd_workspace_bytes(NP, Cc, fh, fw), r.device)
        call("clc_match_bwd", C.byref(view), ptr(r), ctx.r_cl, ptr(mask), ptr(idx), ptr(weights), temperature,
             ptr(g_out.contiguous()), ptr(g_r), ptr(g_q), None, NP, P, Cc, ph, pw, fh, fw, k, 1, ptr(ws), ws.numel(),
             _stream())
        ctx.fwd_ws = None
        return g_q, g_r, None, None, None, None, None, None, None


def match_and_gather(y, refs, patch_h=4, patch_w=4, k=4, temperature=15.0, gaussian_mask=True,
                     is_stack=False, mode="tc", strict=False):
    """match_topk + gather/blend: aligned references [B, R, C(*k), h, w].  `strict`: see match_topk."""
    if is_stack:
        val, idx, r = match_topk(y, refs, patch_h, patch_w, k, gaussian_mask, mode)
        B, R, P, _ = val.shape
        h, w = r.shape[2], r.shape[3]
        out = _GatherBlendFn.apply(r, val.view(B * R, P, k), idx.view(B * R, P, k), patch_h, patch_w,
                                   w - patch_w + 1, temperature, is_stack)
        return out.view(B, R, -1, h, w)
    _check(y, "y")
    if isinstance(refs, (list, tuple)):
        refs = torch.stack(list(refs), dim=1)
    B, R, Cc, h, w = refs.shape
    r = refs.reshape(B * R, Cc, h, w).contiguous()
    mask = _cached_mask(h, w, patch_h, patch_w, y.device) if gaussian_mask else None
    out, _, _ = _MatchGatherFn.apply(y.contiguous(), r, mask, patch_h, patch_w, int(k), R, float(temperature), mode)
    if strict and mode == "tc" and last_uncertified(y.device) > 0:
        out, _, _ = _MatchGatherFn.apply(y.contiguous(), r, mask, patch_h, patch_w, int(k), R, float(temperature), "fp32")
    return out.view(B, R, Cc, h, w)
