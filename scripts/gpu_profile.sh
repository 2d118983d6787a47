#!/bin/bash
# Reproduces the ncu evidence under profiles/ (run on the GPU box through gpurun):
#   scripts/gpu_profile.sh <tag> [workload ...]      e.g.  scripts/gpu_profile.sh r1b cfg2 cfg4
# For each workload: (1) the launch list of scripts/profile_step.py (gpu__time_duration.sum,
# --clock-control none), (2) one `--set full` capture of the last step's kernels.
# Outputs land in gpurun_out/ (reports) -- summarise them with scripts/ncu_summary.py / ncu_report.py.
set -u
tag=$1; shift
mkdir -p gpurun_out
for wl in "$@"; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/${tag}_launches_${wl}.csv \
      python scripts/profile_step.py --workload $wl --steps 3 > gpurun_out/${tag}_launches_${wl}.log 2>&1
  # full set: only this library's kernels (base-name filter skips the ATen randomize()/fill kernels)
  ncu --set full --clock-control none --import-source on -k regex:'^(prepass|rescore|select|match_|clm_|eb_|gc_|lrp_|cl_to|nchw_|gather|patch_stats|channel|topk|pearson|bpp)' \
      -c 60 -o gpurun_out/${tag}_full_${wl} -f \
      python scripts/profile_step.py --workload $wl --steps 1 > gpurun_out/${tag}_full_${wl}.log 2>&1
done
ls -la gpurun_out
