#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_chessnet_golden.py -m gpu -q -p no:cacheprovider -x -s -k "fused_trunk or golden" ) > gpurun_out/f_pytest.log 2>&1
tail -30 gpurun_out/f_pytest.log
timeout 120 python tools/trunk_scaling.py 2048 4096 2>&1 | tail -6
