#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_trunk.py > gpurun_out/e_plain.log 2>&1 && \
ncu --set full --import-source on --cache-control none --clock-control none -k regex:conv_pad_kernel -s 24 -c 4 -f -o gpurun_out/r02_conv_pad_full python tools/ncu_trunk.py > gpurun_out/e_ncu.log 2>&1
tail -5 gpurun_out/e_ncu.log
ls -la gpurun_out/*.ncu-rep
