#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 5 --warmup 3 --ref-gpu 0 --legacy-cpu 0 --cpu-budget 1 --no-root-line ) > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -3 gpurun_out/d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])
r=d['roofline']; print({k:r[k] for k in r if k not in ('kernel','peak_kind','traffic_source')})
print(d['clocks'])
PY
( LZB_CONV_IMPL=1 timeout 900 python bench.py --steps 5 --warmup 3 --ref-gpu 0 --legacy-cpu 0 --cpu-budget 1 --no-root-line ) > gpurun_out/d_bench_old.json 2> gpurun_out/d_bench_old.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench_old.json').read().strip().splitlines()[-1])
print('OLD KERNEL value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'])
r=d['roofline']; print({k:r[k] for k in r if k not in ('kernel','peak_kind','traffic_source')})
PY
