#!/bin/bash
# Round profile capture (run under gpurun, ONE GPU): launch list of the bench command + full capture of the top kernel
# + DRAM bytes of one bench-sized launch.
# Usage: tools/profile.sh <tag> [humanoid|kbot]  -> gpurun_out/<tag>_launches.csv, <tag>_rollout.ncu-rep, <tag>_traffic.csv
set -e
TAG=${1:-r01}
WL=${2:-humanoid}
N=4096; T=4
if [ "$WL" = kbot ]; then N=4096; T=2; fi
mkdir -p gpurun_out
python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/ncu_target.py $N $T 3 $WL > gpurun_out/${TAG}_target_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 2 -c 1 -o gpurun_out/${TAG}_rollout \
    python tools/ncu_target.py $N $T 3 $WL > gpurun_out/${TAG}_ncu_target.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_rollout -s 3 -c 1 --csv \
    --log-file gpurun_out/${TAG}_traffic.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo done
