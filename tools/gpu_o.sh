#!/bin/bash
for u in 1 8 32 1 8; do echo "unroll $u: $(LZB_WAVE_UNROLL=$u python bench.py --steps 4 --warmup 3 --profile-only 2>/dev/null | tail -1 | cut -c1-140)"; done
