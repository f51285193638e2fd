#!/bin/bash
# round 2, GPU call 30: PatchMatch level steps as a replayed graph -- parity, stage times
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_pipeline.py tests/test_oracle_ref.py -m gpu -q -x 2>&1 | tail -2
for gph in 1 0; do
NCT_PM_GRAPH=$gph timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 8 > gpurun_out/c30_bench_g$gph.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('pm graph', sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['ms_per_pair'], d['roofline']['avg_launch_ms'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'), d['gpu_launches'])" gpurun_out/c30_bench_g$gph.json $gph
done
