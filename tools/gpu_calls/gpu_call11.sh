#!/bin/bash
# round 2, GPU call 11: experiments -- PatchMatch-only stage events in the co-running pass; PatchMatch CTAs per SM capped by shared-memory padding
set -u
mkdir -p gpurun_out
for lvl in 2 1; do
NCT_BENCH_PROFILE_LEVEL=$lvl NCT_BENCH_TRACE=1 timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-f16-line > gpurun_out/c11_bench_lvl$lvl.json 2> gpurun_out/c11_lvl$lvl.err
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('profile level', sys.argv[2], d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['single_stream']['avg_launch_ms'])" gpurun_out/c11_bench_lvl$lvl.json $lvl
done
for pad in 60000 80000; do
NCT_PM_SMEM_PAD=$pad timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-f16-line > gpurun_out/c11_bench_pad$pad.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('smem pad', sys.argv[2], d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'])" gpurun_out/c11_bench_pad$pad.json $pad
done
