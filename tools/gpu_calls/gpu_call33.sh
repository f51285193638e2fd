#!/bin/bash
# round 2, GPU call 33: parity suite with the two-pixels-per-warp BDS error kernel (C = 64) + A/B of the matvec's blocks-per-SM
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c33_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/c33_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c33_smoke.log 2>&1; echo "smoke rc=$?"
for mb in 2 3 4; do
  NCT_NL_MINB=$mb timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c33_bench_minb$mb.json 2> gpurun_out/c33_minb$mb.err; echo "bench minb=$mb rc=$?"
  grep '^{' gpurun_out/c33_bench_minb$mb.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'])"
done
NCT_NL_MINB=3 timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_color.py -x -q -m gpu > gpurun_out/c33_pytest_minb3.log 2>&1; echo "pytest minb3 rc=$?"; tail -1 gpurun_out/c33_pytest_minb3.log
