#!/bin/bash
# round 2, GPU call 1: fixed-point conv engine bring-up + PM_SPEC measurement + stage timings
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python tools/q_diag.py > gpurun_out/q_diag.log 2>&1; echo "q_diag rc=$?"; tail -40 gpurun_out/q_diag.log
timeout 900 python -m pytest tests/test_gpu_vgg_q.py -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo "pytest_q rc=$?"; tail -15 gpurun_out/pytest_q.log
for e in 2 3; do
  timeout 600 python bench.py --pairs-in-flight 1 --no-cpu-baseline --steps 3 --vgg-engine $e > gpurun_out/c1_bench_e$e.json 2> gpurun_out/c1_bench_e$e.err; echo "bench e$e rc=$?"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c1_bench_e$e.json
done
NCT_PM_SPEC=1 timeout 600 python -m pytest tests/test_gpu_pm.py -m gpu -q -x 2>&1 | tail -2
NCT_PM_SPEC=1 timeout 600 python bench.py --pairs-in-flight 1 --no-cpu-baseline --steps 3 > gpurun_out/c1_bench_spec.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['ms_per_step'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c1_bench_spec.json
