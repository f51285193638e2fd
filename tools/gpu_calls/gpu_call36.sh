#!/bin/bash
# round 2, GPU call 36: final state with the mean + max depth criterion: parity suite, the depth test (incl. low-roughness holes), smoke, bench
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -k "not adaptive_bottom" > gpurun_out/c36_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/c36_pytest.log
timeout 100 python -m pytest tests/test_gpu_color.py -q -s -m gpu -k "adaptive_bottom" > gpurun_out/c36_depth_test.log 2>&1; echo "depth test rc=$?"; grep "iterations at full\|passed\|failed\|Error\|assert" gpurun_out/c36_depth_test.log | head -20
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c36_smoke.log 2>&1; echo "smoke rc=$?"
timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/c36_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(d['value'], d['e2e']['value'], d['parity'].get('bytes_differing_from_committed_700x700_golden'), d['stage_ms_per_pair_single_stream'])"
