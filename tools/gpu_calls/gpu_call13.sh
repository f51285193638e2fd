#!/bin/bash
# round 2, GPU call 13 (8 GPUs): N = 1 and N = 8 on the same box
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 400 python bench.py --gpus 1 --steps 6 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c13_bench_n1.json 2> gpurun_out/c13_n1.err; echo "N=1 rc=$?"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('N1', d['value'], d['ms_per_step'], d['e2e']['value'])" gpurun_out/c13_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --no-f16-line > gpurun_out/c13_bench_n8.json 2> gpurun_out/c13_n8.err; echo "N=8 rc=$?"; tail -3 gpurun_out/c13_n8.err | cut -c1-300
grep '^{' gpurun_out/c13_bench_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('N8', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
