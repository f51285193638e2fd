#!/bin/bash
# Measurement queue left by round 1 (run on a B200 box, e.g. `gpurun -- bash tools/round2_queue.sh`); everything here is
# already compiled into libnct.so behind environment switches whose defaults are the measured round-1 configuration.
set -u
mkdir -p gpurun_out
stage() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[2], d['ms_per_step'], d['stage_ms_per_pair_single_stream'])" "$1" "$2"; }

# 1. speculative next-candidate prefetch in the PatchMatch random search (C = 64): parity first, then time
NCT_PM_SPEC=1 python -m pytest tests/test_gpu_pm.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -2
python bench.py --pairs-in-flight 1 --no-cpu-baseline --steps 4 > gpurun_out/q_base.json 2>/dev/null; stage gpurun_out/q_base.json base
NCT_PM_SPEC=1 python bench.py --pairs-in-flight 1 --no-cpu-baseline --steps 4 > gpurun_out/q_spec.json 2>/dev/null; stage gpurun_out/q_spec.json pm_spec

# 2. full ncu launch list of one pair (~6200 launches, ~0.1 s each under ncu: give it 15 minutes)
NCT_BENCH_PROFILE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --pairs-in-flight 1 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench.csv

# 3. BASELINE configs[4]: PatchMatch iteration sweep at relu3_1 geometry -- parity at every count is in
#    tests/test_gpu_pm.py::test_config5_iteration_sweep; per-level timings of the 700^2 shapes:
python tools/pm_levels.py 700 2>&1 | tail -6
