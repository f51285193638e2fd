#!/bin/bash
# round 2, GPU call 29: pairs in flight per GPU with the final kernels
set -u
mkdir -p gpurun_out
for P in 4 6 8 12; do
  timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 8 --pairs-in-flight $P > gpurun_out/c29_bench_p$P.json 2>/dev/null
  python -c "import json,sys; d=json.load(open(sys.argv[1])); print('P', sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c29_bench_p$P.json $P
done
