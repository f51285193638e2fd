#!/bin/bash
# round 2, GPU call 19: exact dense coarse solve in the WLS V-cycle -- parity, iteration counts, stage times
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_pipeline.py -m gpu -q -x -s -k "wls or golden or independent" > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?"; grep -h "dense coarse\|passed\|failed\|rror" gpurun_out/c19_pytest.log | tail -8
for d in 1 0; do
NCT_MG_DENSE=$d NCT_WLS_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 6 --pairs-in-flight 1 > gpurun_out/c19_bench_dense$d.json 2> gpurun_out/c19_dense$d.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('dense', sys.argv[2], d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c19_bench_dense$d.json $d
grep "MG-PCG" gpurun_out/c19_dense$d.err | tail -5
done
timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 8 > gpurun_out/c19_bench_p6.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('P6', d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c19_bench_p6.json
