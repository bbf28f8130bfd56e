#!/bin/bash
python tools/trunk_exp2.py 0 128 32 16 4 20 2>&1 | grep -v "timed out"
for s in 2 3 4; do echo "w_stages=$s"; LZB_TRUNK_W_STAGES=$s python tools/trunk_exp2.py 0 2>&1 | grep -v "timed out"; done
