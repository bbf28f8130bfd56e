#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rules.py tests/test_gpu_root_ops.py tests/test_gpu_scalar_api.py tests/test_rule_cases.py tests/test_legacy_golden.py tests/test_gpu_tree.py tests/test_gpu_reference_over_shim.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
python tools/ncu_ops.py > gpurun_out/o_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"encode_actions_kernel|apply_moves|model_input_kernel|legal_masks|apply_actions" -s 5 -c 5 -f -o gpurun_out/r02_rule_ops_full_v2 python tools/ncu_ops.py > gpurun_out/o_ncu.log 2>&1
tail -n 2 gpurun_out/o_ncu.log
python tools/bench_ops.py > gpurun_out/o_bench_ops.json 2> gpurun_out/o_bench_ops.err; tail -n 2 gpurun_out/o_bench_ops.err; grep -E "\"op\"|ours_ms|frac_of|speedup|kernel_ms|kernel_gbs" gpurun_out/o_bench_ops.json | paste - - - - - - - | head -12
python bench.py --workload playout --steps 5 --warmup 3 2>/dev/null | head -c 700
