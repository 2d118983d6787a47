#!/usr/bin/env python
"""Determinism stress: the tc match call must return bit-identical (val, idx, aligned) every time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import clc_b200  # noqa: E402
from clc_b200 import _lib  # noqa: E402
from clc_b200.ops import _stream  # noqa: E402

d = torch.device("cuda:0")
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 300
NQ, R, Cc, h, w, p, k = 2, 3, 320, 16, 16, 4, 4
bad = 0
for seed in range(3):
    g = torch.Generator().manual_seed(100 + seed)
    y = torch.randn(NQ, Cc, h, w, generator=g).to(d)
    refs = (0.5 * y.cpu().unsqueeze(1) + torch.randn(NQ, R, Cc, h, w, generator=g)).to(d)
    r = refs.reshape(NQ * R, Cc, h, w).contiguous()
    P = (h // p) * (w // p)
    ref_out = None
    for it in range(n_iter):
        val = torch.empty(NQ * R, P, k, device=d)
        idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
        al = torch.empty_like(r)
        wt = torch.empty_like(val)
        nb = _lib.lib().clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
        ws = torch.empty(nb, dtype=torch.uint8, device=d)
        if it % 3 == 1:
            ws.random_(0, 255)          # stale garbage in the workspace must not matter
        _lib.call("clc_match_topk_tc", y.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, 1, val.data_ptr(),
                  idx.data_ptr(), None, 15.0, al.data_ptr(), wt.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
        if it % 5 == 2:                  # interleave the fp32 path like the test does
            clc_b200.match_topk(y, refs, p, p, k, gaussian_mask=True, mode="fp32")
        out = (val.clone(), idx.clone(), al.clone())
        if ref_out is None:
            ref_out = out
        else:
            for name, a, b in zip(("val", "idx", "aligned"), ref_out, out):
                if not torch.equal(a, b):
                    bad += 1
                    diff = (a.float() - b.float()).abs()
                    print(f"seed {seed} iter {it}: {name} differs, max {diff.max().item():.3e}, n {int((diff > 0).sum())}")
torch.cuda.synchronize()
print("mismatches:", bad)
