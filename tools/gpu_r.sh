#!/bin/bash
export LZB_HEADS_OVERLAP=1
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_chessnet_golden.py tests/test_gpu_selfplay.py tests/test_gpu_tree.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
for o in 0 1 0 1; do echo "overlap $o: $(LZB_HEADS_OVERLAP=$o timeout 300 python bench.py --steps 4 --warmup 3 --profile-only 2>&1 | tail -1 | cut -c1-140)"; done
