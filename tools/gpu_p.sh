#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rules.py tests/test_gpu_tree.py tests/test_gpu_selfplay.py tests/test_legacy_golden.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
python bench.py --workload playout --steps 5 --warmup 3 2>/dev/null | cut -c1-200
python bench.py --steps 4 --warmup 3 --profile-only 2>/dev/null | tail -1 | cut -c1-140
