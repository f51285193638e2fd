#!/bin/bash
# round 2, GPU call 20: non-local CG with the index / weight loads hoisted -- parity, stage times
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_color.py tests/test_gpu_pipeline.py -m gpu -q -x -k "nonlocal or golden or independent or ls_cg" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 8 > gpurun_out/c20_bench.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c20_bench.json
