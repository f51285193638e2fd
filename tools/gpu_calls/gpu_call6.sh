#!/bin/bash
# round 2, GPU call 6: suite after pruning the PatchMatch variants, smoke, the ncu launch list with plain launches, value-vs-e2e experiment
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c6_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
NCT_BENCH_SYNC_EACH=1 timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c6_bench_sync_each.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('sync_each', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])" gpurun_out/c6_bench_sync_each.json
timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c6_bench_base.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('base', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline_vgg'], d['parity'])" gpurun_out/c6_bench_base.json
NCT_BENCH_PROFILE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 0 --pairs-in-flight 1 > gpurun_out/c6_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/r2_launches_bench.csv; tail -2 gpurun_out/c6_bench_under_ncu.log
gzip -f gpurun_out/r2_launches_bench.csv
