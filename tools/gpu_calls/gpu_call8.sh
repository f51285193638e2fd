#!/bin/bash
# round 2, GPU call 8: FP16 feature store (parity + throughput line), value-vs-e2e trace
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_pipeline.py -m gpu -q -x -k "fp16 or golden or independent" -s > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?"; grep -h "FP16\|passed\|failed\|Error" gpurun_out/c8_pytest.log | tail -8
NCT_BENCH_TRACE=1 timeout 900 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "bench rc=$?"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('value_fp16_feature_store'))" gpurun_out/c8_bench.json
grep trace gpurun_out/c8_bench.err | head -14
