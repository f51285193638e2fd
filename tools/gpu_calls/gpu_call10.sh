#!/bin/bash
# round 2, GPU call 10: full suite (FP16 store, nct_solve_direct), smoke, the bench line in its final structure
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c10_pytest.log; grep -h "nct_solve_direct\|FP16 feature" gpurun_out/c10_pytest.log | head
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c10_bench.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d.get('value_fp16_feature_store',{}).get('value'), d['cpu_baseline']['value'], d['parity'])" gpurun_out/r2_bench_n1.json
