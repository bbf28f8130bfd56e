#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_chessnet_golden.py -m gpu -q -p no:cacheprovider -x -s ) > gpurun_out/c_pytest.log 2>&1
tail -25 gpurun_out/c_pytest.log
timeout 600 python tools/conv_decompose.py > gpurun_out/c_conv_decompose.txt 2>&1
grep -v "timed out" gpurun_out/c_conv_decompose.txt
