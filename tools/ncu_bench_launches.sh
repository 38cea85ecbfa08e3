#!/bin/bash
# Launch list of the bench command itself (first 400 kernel launches: the warm-up epochs at B = 256), durations only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --kill 1 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02_ncu_bench.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r02_launches_bench.csv; tail -3 gpurun_out/r02_ncu_bench.log
