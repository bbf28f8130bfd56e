#!/bin/bash
for d in 8 28 24 12 40; do echo "=== LZB_TRUNK_DEBUG=$d ==="; LZB_TRUNK_DEBUG=$d timeout 120 python tools/trace_trunk.py 2>&1 | grep -v "timed out" | tail -32; done
