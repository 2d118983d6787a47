#!/usr/bin/env python
"""Bitstream path: throughput of the host C coder (clc_rans_encode / clc_rans_decode)
on the stream of one Kodak-shaped image (491 520 y symbols, 64 scale-table CDFs), next to the oracle's
pure-Python restatement on a bounded prefix; with --model also CLC(N=64).compress/decompress wall time
on cuda:0.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--symbols", type=int, default=491520)
ap.add_argument("--model", action="store_true")
a = ap.parse_args()

import clc_b200  # noqa: E402
from clc_b200 import ans as A  # noqa: E402
from clc_b200.models import get_scale_table  # noqa: E402

gc = clc_b200.GaussianConditional(None)
gc.update_scale_table(get_scale_table())
tab = gc.coder_tables()
rng = np.random.default_rng(0)
idx = rng.integers(0, 64, a.symbols).astype(np.int32)
scale = get_scale_table().numpy()[idx]
sym = np.rint(rng.standard_normal(a.symbols) * scale).astype(np.int32)
best_e = best_d = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    s = A.RansEncoder().encode_with_indexes(sym, idx, tab, None, None)
    t1 = time.perf_counter()
    out = A.RansDecoder().decode_with_indexes(s, idx, tab, None, None, as_tensor=True)
    t2 = time.perf_counter()
    best_e, best_d = min(best_e, t1 - t0), min(best_d, t2 - t1)
assert (out.numpy() == sym).all()
res = {"symbols": a.symbols, "bytes": len(s), "bits_per_symbol": 8 * len(s) / a.symbols,
       "c_encode_msym_s": a.symbols / best_e / 1e6, "c_decode_msym_s": a.symbols / best_d / 1e6}
import oracle  # noqa: E402
oracle.enable_shim()
from compressai import ans as O  # noqa: E402
n = 20000
cd, sz, of = tab.cdfs.tolist(), tab.sizes.tolist(), tab.offsets.tolist()
t0 = time.perf_counter()
so = O.RansEncoder().encode_with_indexes(sym[:n].tolist(), idx[:n].tolist(), cd, sz, of)
t1 = time.perf_counter()
O.RansDecoder().decode_with_indexes(so, idx[:n].tolist(), cd, sz, of)
t2 = time.perf_counter()
assert so == A.RansEncoder().encode_with_indexes(sym[:n], idx[:n], tab, None, None)
res.update(oracle_encode_msym_s=n / (t1 - t0) / 1e6, oracle_decode_msym_s=n / (t2 - t1) / 1e6, oracle_sample=n)
if a.model:
    from clc_b200.models import CLC
    from oracle import detfill
    d = torch.device("cuda:0")
    m = detfill.fill_(CLC(N=64), seed=0).eval().to(d)
    m.update()
    x = detfill.det_image((1, 3, 512, 768), 1).to(d)
    refs = [detfill.det_image((1, 3, 512, 768), 2 + i).to(d) for i in range(3)]
    with torch.no_grad():
        for _ in range(2):
            o = m.compress(x, refs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        o = m.compress(x, refs)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        r = m.decompress(o["strings"], o["shape"], refs)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    res.update(model="CLC N=64, 1x3x512x768, 3 refs", compress_ms=1e3 * (t1 - t0), decompress_ms=1e3 * (t2 - t1),
               bpp=8 * (len(o["strings"][0][0]) + len(o["strings"][1][0])) / (512 * 768))
print(json.dumps(res))
