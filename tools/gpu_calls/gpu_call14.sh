#!/bin/bash
# round 2, GPU call 14: tile-fused V-cycle legs (parity + timing per level threshold)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_color.py -m gpu -q -x -s -k "wls" > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; grep -h "iterations unfused\|passed\|failed\|Error\|error" gpurun_out/c14_pytest.log | tail -8
NCT_MG_FUSED_MIN_N=1 timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_pipeline.py -m gpu -q -x -k "wls or golden or independent" 2>&1 | tail -2
for n in 0 400000 100000 20000 1; do
  NCT_MG_FUSED_MIN_N=$n timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 6 > gpurun_out/c14_bench_fused$n.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('fused_min_n', sys.argv[2], d['value'], d['e2e']['value'], d['gpu_launches'], d['parity'].get('bytes_differing_from_committed_700x700_golden'), d['stage_ms_per_pair_single_stream']['wls'])" gpurun_out/c14_bench_fused$n.json $n
done
