#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "cudnn_path or no_library or fused_heads or fast_path_equals or fused_trunk" ) > gpurun_out/h_pytest.log 2>&1
tail -8 gpurun_out/h_pytest.log
( timeout 900 python bench.py --steps 5 --warmup 3 --ref-gpu 0 --legacy-cpu 0 --cpu-budget 1 --no-root-line ) > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -3 gpurun_out/h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'], d['e2e']['step'])
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','kernel_ms','forward_ms','wave_ms')})
print(d['clocks'])
PY
