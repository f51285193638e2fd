#!/bin/bash
# round 2, GPU call 2: full GPU suite with the graph-captured solvers + new goldens, smoke (plain and under ncu), WLS loop modes, bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c2_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/c2_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/c2_smoke_launches.csv python __graft_entry__.py smoke > gpurun_out/c2_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?"; tail -3 gpurun_out/c2_smoke_ncu.log
python - <<'PY'
import csv,collections
try:
    rows=[r for r in csv.reader(open('gpurun_out/c2_smoke_launches.csv')) if len(r)>5]
    c=collections.Counter(r[4].split('(')[0][-60:] for r in rows[1:])
    print(len(rows)-1,'launches;',len(c),'kernels'); print(sorted(c.items(),key=lambda t:-t[1])[:40])
except Exception as e: print('csv',e)
PY
for m in 2 1 0; do
  NCT_WLS_LOOP=$m timeout 600 python bench.py --pairs-in-flight 1 --no-cpu-baseline --steps 3 > gpurun_out/c2_bench_p1_loop$m.json 2> gpurun_out/c2_bench_p1_loop$m.err; echo "bench P1 loop$m rc=$?"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c2_bench_p1_loop$m.json
done
timeout 900 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c2_bench_p6.json 2> gpurun_out/c2_bench_p6.err; echo "bench P6 rc=$?"; cat gpurun_out/c2_bench_p6.json; tail -3 gpurun_out/c2_bench_p6.err
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 900 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c2_bench_p6_conn8.json 2>/dev/null; python -c "import json,sys; d=json.load(open(sys.argv[1])); print('conn8', d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c2_bench_p6_conn8.json
timeout 900 python bench.py --no-cpu-baseline --steps 8 --pairs-in-flight 10 > gpurun_out/c2_bench_p10.json 2>/dev/null; python -c "import json,sys; d=json.load(open(sys.argv[1])); print('P10', d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c2_bench_p10.json
