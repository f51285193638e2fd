#!/bin/bash
# round 2, GPU call 3: full suite (vis / resume / noise floor), config[4] sweep, config[3] bench, ncu launch list + full capture of the PM kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c3_pytest.log
grep -h "noise floor\|end to end vs\|end-to-end\|700x700\|k-NN vs\|WLS .* iterations in every" gpurun_out/c3_pytest.log | head -30
timeout 300 python tools/pm_sweep.py r2_pm_sweep_config4 > gpurun_out/c3_sweep.log 2>&1; echo "sweep rc=$?"; tail -11 gpurun_out/c3_sweep.log
timeout 600 python bench.py --side 1000 --steps 4 --no-cpu-baseline > gpurun_out/r2_bench_side1000.json 2> gpurun_out/c3_bench1000.err; echo "bench 1000 rc=$?"; cut -c1-600 gpurun_out/r2_bench_side1000.json; tail -3 gpurun_out/c3_bench1000.err
timeout 600 python tools/pm_levels.py 700 > gpurun_out/c3_pm_levels.log 2>&1; tail -6 gpurun_out/c3_pm_levels.log
# ncu: full metric set for the PM step kernel at the finest level (steps 176-183 = iteration 4 of level 4: jump 8,4,2,1 twice)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pm_step_t_kernel -s 176 -c 8 -o gpurun_out/r2_pm_step_full python tools/one_pair.py 700 1 > gpurun_out/c3_ncu_pm.log 2>&1; echo "ncu pm rc=$?"; tail -2 gpurun_out/c3_ncu_pm.log
# ncu: launch list of one pair through the bench's own command (warm-up 1 + 1 timed step, one pair in flight)
NCT_BENCH_PROFILE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --pairs-in-flight 1 > gpurun_out/c3_bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"; wc -l gpurun_out/r2_launches_bench.csv; tail -2 gpurun_out/c3_bench_under_ncu.log
gzip -f gpurun_out/r2_launches_bench.csv
