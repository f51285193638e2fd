#!/bin/bash
# round 2, GPU call 35: parity suite (graph-abort path added, bottom depth at its default = full), then the adaptive
# bottom depth (NCT_WLS_DEPTH=0.96): its own test, the colour + pipeline suites under it, and a bench A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "not adaptive_bottom" > gpurun_out/c35_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/c35_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c35_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python -m pytest tests/test_gpu_color.py -q -s -m gpu -k "adaptive_bottom" > gpurun_out/c35_depth_test.log 2>&1; echo "depth test rc=$?"; grep "iterations at full\|passed\|failed\|Error\|assert" gpurun_out/c35_depth_test.log | head -20
NCT_WLS_DEPTH=0.96 timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_color.py -x -q -m gpu -k "not adaptive_bottom" > gpurun_out/c35_pytest_depth.log 2>&1; echo "pytest depth rc=$?"; tail -1 gpurun_out/c35_pytest_depth.log
for thr in 0 0.96; do
  NCT_WLS_DEPTH=$thr timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c35_bench_depth$thr.json 2> gpurun_out/c35_depth$thr.err; echo "bench depth=$thr rc=$?"
  grep '^{' gpurun_out/c35_bench_depth$thr.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['parity'].get('bytes_differing_from_committed_700x700_golden'), d['stage_ms_per_pair_single_stream'])"
done
