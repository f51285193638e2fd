#!/bin/bash
# round 2, GPU call 16: 2-D query tiles in the PatchMatch kernel (parity, per-level timing vs the 1 x 32 row tiles, bench), ncu of the conv and PM kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -2
echo "== 2-D tiles"; timeout 300 python tools/pm_levels.py 700 2>&1 | tail -6
echo "== 1-D tiles"; NCT_PM_TILE1D=1 timeout 300 python tools/pm_levels.py 700 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c16_bench_2d.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('2-D', d['value'], d['e2e']['value'], d.get('value_fp16_feature_store',{}).get('value'), d['roofline']['achieved'], d['roofline']['frac'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c16_bench_2d.json
NCT_PM_TILE1D=1 timeout 600 python bench.py --no-cpu-baseline --steps 8 > gpurun_out/c16_bench_1d.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('1-D', d['value'], d['e2e']['value'], d.get('value_fp16_feature_store',{}).get('value'), d['roofline']['achieved'], d['stage_ms_per_pair_single_stream']['patchmatch'])" gpurun_out/c16_bench_1d.json
# ncu --set full: the tensor-core conv kernels of one forward (tensor-pipe utilisation) and two PM launches with the 2-D tiles
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_i8_kernel -c 13 -o gpurun_out/r2_conv_i8_full python tools/one_pair.py 700 1 > gpurun_out/c16_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:pm_step_t_kernel -s 176 -c 4 -o gpurun_out/r2_pm_step_2d_full python tools/one_pair.py 700 1 > gpurun_out/c16_ncu_pm.log 2>&1; echo "ncu pm rc=$?"
