#!/bin/bash
# round 2, GPU call 7: the bench line after the warm-up reorder (with the CPU baseline leg), the ncu launch list of one pair
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c7_bench.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline'], d['clocks'])" gpurun_out/r2_bench_n1.json
NCT_BENCH_PROFILE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 0 --pairs-in-flight 1 > gpurun_out/c7_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/r2_launches_bench.csv; tail -2 gpurun_out/c7_bench_under_ncu.log
gzip -f gpurun_out/r2_launches_bench.csv
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?"; cut -c1-1500 gpurun_out/r2_bench_reference_arm.json
