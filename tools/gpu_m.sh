#!/bin/bash
for v in 0 1 2 0 1 2; do echo "variant $v: $(LZB_TREE_VARIANT=$v python bench.py --steps 4 --warmup 3 --profile-only 2>/dev/null | tail -1 | cut -c1-160)"; done
