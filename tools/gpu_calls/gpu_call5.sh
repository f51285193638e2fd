#!/bin/bash
# round 2, GPU call 5 (2 GPUs): the streamed NCCL gather path -- N = 1 and N = 2 back to back on the same box
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nproc
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err; echo "N=1 rc=$?"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('N1', d['value'], d['ms_per_step'], d['e2e'])" gpurun_out/c5_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err; echo "N=2 rc=$?"; tail -5 gpurun_out/c5_bench_n2.err
grep '^{' gpurun_out/c5_bench_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('N2', d['value'], d['ms_per_step'], d['e2e'], d['config']['collective'][:40])"
