#!/bin/bash
python tools/trunk_exp2.py 0 512 0 512 2>&1 | grep -v "timed out"
for d in 0 512 0 512; do echo "debug $d: $(LZB_TRUNK_DEBUG=$d python bench.py --steps 4 --warmup 3 --profile-only 2>/dev/null | tail -1 | cut -c1-130)"; done
