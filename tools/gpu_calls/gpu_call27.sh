#!/bin/bash
# round 2, GPU call 27: final state -- full suite, smoke, bench line with the CPU baseline
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c27_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 12 --warmup 4 > gpurun_out/r2_bench_n1.json 2> gpurun_out/c27_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c27_bench.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['co_running']['frac'], d['roofline_vgg']['tensor_pipe_work']['frac'], d.get('value_fp16_feature_store',{}).get('value'), d['cpu_baseline']['value'], d['parity'].get('bytes_differing_from_committed_700x700_golden'), d['stage_ms_per_pair_single_stream'], d['roofline']['ms_per_pair'], d['gpu_launches'], d['clocks'])" gpurun_out/r2_bench_n1.json
