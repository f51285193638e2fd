#!/bin/bash
# round 2, GPU call 34: per-launch times of the BDS error kernel with the two-pixels-per-warp variant at C = 64
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"^bds_feature_error_kernel|^reconstruct_bds_kernel|^inv_" -c 80 --csv --log-file gpurun_out/c34_bds.csv python tools/one_pair.py 700 1 > gpurun_out/c34.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, io
txt=open("gpurun_out/c34_bds.csv").read()
rows=list(csv.DictReader(io.StringIO('\n'.join(l for l in txt.split('\n') if l.startswith('"')))))
for r in rows:
    if r["Metric Name"]=="gpu__time_duration.sum": print(r["Kernel Name"][:50], r["Grid Size"], r["Metric Value"])
PY
