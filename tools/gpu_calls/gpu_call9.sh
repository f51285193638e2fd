#!/bin/bash
set -u
mkdir -p gpurun_out
NCT_BENCH_NO_PROFILE=1 NCT_BENCH_TRACE=1 timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-f16-line > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c9_bench.err | cut -c1-300
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])" gpurun_out/c9_bench.json
grep "trace. dev" gpurun_out/c9_bench.err | head -7
