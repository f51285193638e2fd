#!/bin/bash
# round 2, GPU call 31 (4 GPUs): N = 4 with the final code
set -u
mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c31_bench_n4.json 2> gpurun_out/c31_n4.err; echo "N=4 rc=$?"; tail -2 gpurun_out/c31_n4.err | cut -c1-200
grep '^{' gpurun_out/c31_bench_n4.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('N4', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
