#!/bin/bash
mkdir -p gpurun_out
python tools/ncu_ops.py > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"encode_actions_kernel|apply_moves|model_input_kernel|legal_masks|apply_actions" -s 5 -c 10 -f -o gpurun_out/r02_rule_ops_full python tools/ncu_ops.py > gpurun_out/o_ncu.log 2>&1
tail -3 gpurun_out/o_plain.log gpurun_out/o_ncu.log
python tools/bench_ops.py > gpurun_out/o_bench_ops.json 2> gpurun_out/o_bench_ops.err; tail -2 gpurun_out/o_bench_ops.err; grep -E '"op"|ours_ms|frac_of|speedup' gpurun_out/o_bench_ops.json | paste - - - - | head -12
