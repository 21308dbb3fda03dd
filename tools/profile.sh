#!/bin/bash
# Round profile capture (run under gpurun, ONE GPU): launch list of the bench command + full capture of the top kernel.
# Usage: tools/profile.sh <tag>   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_rollout.ncu-rep
set -e
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_target.py 4096 4 3 > gpurun_out/${TAG}_target_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 2 -c 1 -o gpurun_out/${TAG}_rollout \
    python tools/ncu_target.py 4096 4 3 > gpurun_out/${TAG}_ncu_target.log 2>&1
echo done
