#!/bin/bash
# bring-up loop on the GPU box: two parity cases, a short bench, the launch list of one step (tools; not part of the product)
mkdir -p gpurun_out
timeout 300 python tools/gpu_check.py ${CASES:-replica_color_mapper replica_color_tracker scannet_color_tracker_exposure tum_color_mapper_dynr} 2>&1 | grep -v Warn | tail -8
python bench.py --steps 50 --warmup 5 --no-cpu-baseline ${BENCH_ARGS:---no-extra} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -c 400 gpurun_out/bench_q.err
python -c "
import json;d=json.load(open('gpurun_out/bench_q.json'));print('value',d['value'],'ms',d['ms_per_step'],d['kernels'],'e2e',d['e2e']['value'],'launches',d['gpu_launches']);[print(' ',k,v) for k,v in d.get('extra',{}).items()]"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-450} -c ${COUNT:-200} --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_q.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
[print(r[ki][:70], r[vi]) for r in rows[1:] if "lsr::" in r[ki] and "sample_rays" not in r[ki]]
PY
