#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_selfplay.py tests/test_gpu_eval.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
LZB_TREE_TRACE=1 python tools/trace_tree.py 2>&1 | tail -9
for v in 0 0; do echo "variant $v: $(LZB_TREE_VARIANT=$v python bench.py --steps 4 --warmup 3 --profile-only 2>/dev/null | tail -1 | cut -c1-140)"; done
