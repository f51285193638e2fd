#!/bin/bash
# round 2, GPU call 18: restriction with up-front loads -- WLS parity, single-stream stage times
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_pipeline.py -m gpu -q -x -k "wls or golden or independent" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 6 > gpurun_out/c18_bench.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['e2e']['value'], d['roofline']['ms_per_pair'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c18_bench.json
