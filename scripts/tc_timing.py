#!/usr/bin/env python
"""Bring-up: per-CTA pipeline stamps of the tcgen05 match kernel (clock64 deltas, cycles)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from clc_b200 import _lib  # noqa: E402

NQ, R, Cc, h, w, p, k = [int(x) for x in (sys.argv[1:8] if len(sys.argv) > 7 else (8, 3, 320, 16, 16, 4, 4))]
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(0)
y = torch.randn(NQ, Cc, h, w, device=d, generator=g)
r = (0.5 * y.unsqueeze(1) + torch.randn(NQ, R, Cc, h, w, device=d, generator=g)).reshape(NQ * R, Cc, h, w).contiguous()
P = (h // p) * (w // p)
val = torch.empty(NQ * R, P, k, device=d)
idx = torch.empty(NQ * R, P, k, dtype=torch.int32, device=d)
H = _lib.debug_lib()
fn = H.clc_debug_match_tc_timing
nb = H.clc_match_topk_tc_workspace_bytes(NQ * R, R, Cc, h, w, p, p, k)
ws = torch.empty(nb, dtype=torch.uint8, device=d)
for it in range(3):
    tm = torch.zeros(148, 16, dtype=torch.int64, device=d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(y.data_ptr(), r.data_ptr(), NQ * R, R, Cc, h, w, p, p, k, 1, val.data_ptr(), idx.data_ptr(), tm.data_ptr(),
            ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, rc
    print("call ms", e0.elapsed_time(e1))
t = tm.cpu()
live = t[:, 0] != 0
t = t[live]
names = ["start", "setup_done", "first_A_issued", "producer_done", "mma_first_B", "mma_first_A", "mma_all_issued",
         "epi_colstat", "epi_acc_full", "epi_done", "epi_sum_done", "mma_wait_acc_empty(sum)", "mma_wait_B_full(sum)",
         "mma_wait_A_full(sum)"]
rel = t - t[:, :1]
print("CTAs:", t.shape[0])
for i, nme in enumerate(names):
    col = rel[:, i].float()
    print(f"{nme:>26}: median {col.median().item():9.0f}  min {col.min().item():9.0f}  max {col.max().item():9.0f} cycles")
