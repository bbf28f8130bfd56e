#!/bin/bash
# One gpurun call: GPU test-suite, smoke, default bench line.  Outputs under gpurun_out/.
#   gpurun --timeout 1700 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 -p no:cacheprovider --durations=20 ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -5 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
head -c 3000 gpurun_out/${TAG}_bench.json
