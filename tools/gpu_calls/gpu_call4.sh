#!/bin/bash
# round 2, GPU call 4: cp.async-staged PatchMatch kernel (parity + timing per staging depth), stream stagger
set -u
mkdir -p gpurun_out
NCT_PM_ASYNC=6,3 timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/c4_pytest_async.log 2>&1; echo "pytest async 6,3 rc=$?"; tail -3 gpurun_out/c4_pytest_async.log
NCT_PM_ASYNC=3,2 timeout 900 python -m pytest tests/test_gpu_pm.py -m gpu -q -x > gpurun_out/c4_pytest_async32.log 2>&1; echo "pytest async 3,2 rc=$?"; tail -2 gpurun_out/c4_pytest_async32.log
for v in 0,0 3,2 4,2 6,3 4,3 6,0 0,3; do
  echo "== NCT_PM_ASYNC=$v"; NCT_PM_ASYNC=$v timeout 300 python tools/pm_levels.py 700 2>&1 | tail -3
done
for v in 6,3 4,2; do
  NCT_PM_ASYNC=$v timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c4_bench_p6_async_$v.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('async', sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c4_bench_p6_async_$v.json $v
done
timeout 600 python bench.py --no-cpu-baseline --steps 8 --stagger-ms 17 > gpurun_out/c4_bench_p6_stagger.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('stagger17', d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c4_bench_p6_stagger.json
timeout 600 python bench.py --no-cpu-baseline --steps 8 --pairs-in-flight 8 --stagger-ms 13 > gpurun_out/c4_bench_p8_stagger.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('P8 stagger13', d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c4_bench_p8_stagger.json
