#!/bin/bash
mkdir -p gpurun_out
echo "== cluster 2 (regression) =="; timeout 100 python tools/trunk_scaling.py 4096 2>&1 | tail -1
echo "== cluster 4 =="; LZB_TRUNK_CLUSTER=4 timeout 100 python tools/trunk_scaling.py 4096 2>&1 | grep -v "timed out" | tail -3
( LZB_TRUNK_CLUSTER=4 timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_chessnet_golden.py -m gpu -q -p no:cacheprovider -x -k "fused_trunk or golden" ) > gpurun_out/i_pytest.log 2>&1
grep -v "timed out" gpurun_out/i_pytest.log | tail -12
