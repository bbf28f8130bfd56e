#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "no_library or test_gpu_eval or scalar_api or noise_and_sampling or fused_heads or policy_target" ) > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
timeout 600 python tools/conv_decompose.py > gpurun_out/b_conv_decompose.txt 2>&1
cat gpurun_out/b_conv_decompose.txt
LZB_CONV_DEBUG=8 timeout 120 python tools/trace_conv.py > gpurun_out/b_conv_trace.txt 2>&1
tail -25 gpurun_out/b_conv_trace.txt
