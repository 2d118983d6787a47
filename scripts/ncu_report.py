#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small text table + traffic entries.

    python scripts/ncu_report.py gpurun_out/r13_full_cfg2.ncu-rep cfg2 profiles/r1_full_cfg2.txt profiles/traffic.json
Writes per kernel: duration, dram bytes read/written (= roofline `traffic`), DRAM / tensor-pipe /
SM throughput percentages, registers, achieved occupancy, top stall reasons.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, workload, out_txt, traffic_json = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}

# kernel function name -> the trace label bench.py uses
LABELS = [("prepass", "clc_match_topk_tc(prepass)"), ("select_kernel", "clc_match_topk_tc(select)"),
          ("match_bwd_own_kernel<320, 1>", "clc_match_clm_bwd(main)"), ("match_bwd_own_kernel<320, 0>", "clc_match_bwd(main)"),
          ("match_bwd_own", "clc_match_bwd(main)"), ("bpp_finalize", "clc_bpp_finalize"),
          ("cl_to_nchw_kernel", "clc_match_bwd(cl_to_nchw)"),
          ("pack_ref", "clc_match_topk_tc(pack_ref)"), ("pack_query", "clc_match_topk_tc(pack_query)"),
          ("patch_stats", "patch_stats"), ("match_gemm", "clc_match_topk_tc(gemm)"),
          ("rescore", "clc_match_topk_tc(rescore)"), ("gather_blend_fwd", "clc_gather_blend_fwd"),
          ("clm_fuse_fwd", "clc_clm_fuse_fwd"), ("clm_fuse_bwd", "clc_clm_fuse_bwd"), ("eb_fwd", "clc_eb_fwd"),
          ("eb_bwd", "clc_eb_bwd"), ("gc_fwd", "clc_gc_fwd"), ("gc_bwd", "clc_gc_bwd"),
          ("lrp_kernel<1, 0>", "clc_lrp_add_fwd"), ("lrp_kernel<1, 1>", "clc_lrp_add_bwd"),
          ("lrp_kernel<0, 0>", "clc_lrp_add_fwd"), ("lrp_kernel<0, 1>", "clc_lrp_add_bwd"),
          ("nchw_to_cl", "clc_match_bwd(nchw_to_cl)"), ("match_bwd_cl", "clc_match_bwd(main)"),
          ("cl_to_nchw_add", "clc_match_bwd(cl_to_nchw)"), ("match_bwd_kernel", "clc_match_bwd")]


def f(r, name, scale=1.0):
    i = ix.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    if u.startswith("Mbyte"):
        v *= 1e6
    elif u.startswith("Kbyte"):
        v *= 1e3
    elif u.startswith("Gbyte"):
        v *= 1e9
    elif u == "ms" or u == "msecond":
        v *= 1e3
    elif u in ("ns", "nsecond"):
        v /= 1e3
    return v * scale


stall_cols = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
traffic = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {}
lines = [f"# ncu --set full --clock-control none, workload {workload}, report {os.path.basename(rep)}",
         "# (cold-cache, serialised replays: compare shares / traffic, not absolute times)",
         f"{'kernel':44s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'dram%':>6s} {'tensor%':>7s} {'sm%':>6s} {'regs':>5s} {'occ%':>5s}  top stalls"]
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    label = next((lab for key, lab in LABELS if key in name), None)
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("clc::", "")
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    dur = f(r, "gpu__time_duration.sum")
    sv = [(float(r[ix[c]].replace(",", "") or 0), c) for c in stall_cols]
    tot = sum(v for v, _ in sv) or 1.0
    st = ", ".join(f"{c.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%"
                   for v, c in sorted(sv, reverse=True)[:3])
    lines.append(f"{short[:44]:44s} {dur or 0:8.1f} {(rd or 0) / 1e6:10.2f} {(wr or 0) / 1e6:10.2f} "
                 f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0:6.1f} "
                 f"{f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0:7.1f} "
                 f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed') or 0:6.1f} "
                 f"{int(f(r, 'launch__registers_per_thread') or 0):5d} "
                 f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active') or 0:5.1f}  {st}")
    if label and rd is not None:
        traffic[f"{workload}:{label}"] = rd + (wr or 0)
open(out_txt, "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(traffic_json, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
