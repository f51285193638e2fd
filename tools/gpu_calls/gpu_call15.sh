#!/bin/bash
# round 2, GPU call 15: final state -- full suite, smoke, the bench line (with the CPU baseline leg)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c15_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/c15_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c15_bench.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['single_stream'], d['roofline_vgg']['tensor_pipe_work'], d.get('value_fp16_feature_store',{}).get('value'), d['cpu_baseline']['value'], d['parity'], d['stage_ms_per_pair_single_stream'], d['gpu_launches'])" gpurun_out/r2_bench_n1.json
