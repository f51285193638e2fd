#!/bin/bash
# round 2, GPU call 21: persistent INT8 conv kernel -- parity (bit-exact suite), per-layer times (ncu), stage time
set -u
mkdir -p gpurun_out
NCT_I8_PERSIST=1 timeout 300 python -m pytest tests/test_gpu_vgg_q.py -m gpu -q -x > gpurun_out/c21_pytest.log 2>&1; echo "pytest persistent rc=$?"; tail -3 gpurun_out/c21_pytest.log
for p in 0 1; do
NCT_I8_PERSIST=$p timeout 300 python bench.py --no-cpu-baseline --no-f16-line --steps 4 --pairs-in-flight 1 > gpurun_out/c21_bench_persist$p.json 2>/dev/null
python -c "import json,sys; d=json.load(open(sys.argv[1])); print('persist', sys.argv[2], d['value'], d['stage_ms_per_pair_single_stream']['vgg'], d['roofline_vgg']['tensor_pipe_work']['frac'], d['parity'].get('bytes_differing_from_committed_700x700_golden'))" gpurun_out/c21_bench_persist$p.json $p
done
NCT_I8_PERSIST=1 timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:conv3x3_i8 -c 13 --csv --log-file gpurun_out/c21_conv_persist.csv python tools/one_pair.py 700 1 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c21_conv_persist.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{})[r[mi]]=r[vi]; d[r[ii]]['grid']=r[gi]; d[r[ii]]['k']=r[ki][:60]
for k,v in d.items(): print(k, v['k'][-40:], v['grid'], v.get('gpu__time_duration.sum'), v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))
PY
