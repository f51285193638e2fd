#!/bin/bash
# round 2, GPU call 22: default mix of conv kernels (persistent for K <= 1152) -- parity, VGG stage time, throughput
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_vgg_q.py tests/test_gpu_vgg.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c22_bench.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['e2e']['value'], d.get('value_fp16_feature_store',{}).get('value'), d['stage_ms_per_pair_single_stream'], d['roofline_vgg']['tensor_pipe_work']['frac'], d['parity'])" gpurun_out/c22_bench.json
