#!/bin/bash
# round 2, GPU call 25: conv1_1 with the conflict-free weight layout -- parity (all engines share the kernel), time
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_vgg_q.py tests/test_gpu_vgg.py -m gpu -q -x 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:conv_first_kernel -c 2 python tools/one_pair.py 700 1 2>&1 | grep -E "gpu__time|bank_conflicts" | head -4
timeout 600 python bench.py --no-cpu-baseline --no-f16-line --steps 6 > gpurun_out/c25_bench.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('bench', d['value'], d['e2e']['value'], d['stage_ms_per_pair_single_stream'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c25_bench.json
