#!/bin/bash
# round 2, GPU call 12: PatchMatch register budget (4 / 5 / 6 CTAs per SM), k-means member lists
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cluster.py -m gpu -q -x 2>&1 | tail -2
for m in 4 5 6; do
  echo "== NCT_PM_MINB=$m"; NCT_PM_MINB=$m timeout 300 python tools/pm_levels.py 700 2>&1 | tail -3
done
for m in 5 6; do
  NCT_PM_MINB=$m timeout 600 python -m pytest tests/test_gpu_pm.py -m gpu -q -x 2>&1 | tail -1
  NCT_PM_MINB=$m timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c12_bench_minb$m.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('minb', sys.argv[2], d['value'], d['e2e']['value'], d.get('value_fp16_feature_store',{}).get('value'), d['stage_ms_per_pair_single_stream'])" gpurun_out/c12_bench_minb$m.json $m
done
