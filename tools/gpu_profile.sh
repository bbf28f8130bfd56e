#!/bin/bash
# ncu evidence for the round: (1) launch list of a short self-play bench, (2) --set full of the trunk kernel, (3) --set full
# of the tree kernel and the heads tail.  Each ncu run follows a plain run of the same command (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --sims 6 --profile-only"
$CMD > gpurun_out/p_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_selfplay_wave_launches.csv $CMD > gpurun_out/p_ncu1.log 2>&1
tail -2 gpurun_out/p_plain.log; tail -2 gpurun_out/p_ncu1.log
python tools/ncu_trunk.py > gpurun_out/p_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:trunk_kernel -s 1 -c 2 -f -o gpurun_out/r02_trunk_full python tools/ncu_trunk.py > gpurun_out/p_ncu2.log 2>&1
tail -2 gpurun_out/p_ncu2.log
CMD3="python bench.py --steps 1 --warmup 3 --sims 12 --profile-only"
$CMD3 > gpurun_out/p_plain3.log 2>&1 && \
ncu --set full --clock-control none -k regex:"tree_expand_select_kernel|heads_tail_kernel" -s 60 -c 4 -f -o gpurun_out/r02_tree_heads_full $CMD3 > gpurun_out/p_ncu3.log 2>&1
tail -2 gpurun_out/p_ncu3.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_selfplay_wave_launches.csv
