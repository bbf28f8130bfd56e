#!/bin/bash
# N-GPU runs (gpurun --gpus N): default bench line at N ranks and the config-4 line.  usage: gpu_multi.sh N TAG [GAMES] [SIMS]
N=${1:-2}; TAG=${2:-m}; GAMES=${3:-4096}; SIMS=${4:-800}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench_n$N.err; head -c 1500 gpurun_out/${TAG}_bench_n$N.json; echo
( time timeout 1500 $TR bench.py --gpus $N --workload config4 --games $GAMES --sims $SIMS ) > gpurun_out/${TAG}_config4_n$N.json 2> gpurun_out/${TAG}_config4_n$N.err
echo "config4 rc=$?"; tail -3 gpurun_out/${TAG}_config4_n$N.err; head -c 3000 gpurun_out/${TAG}_config4_n$N.json
