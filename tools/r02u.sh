#!/bin/bash
# GRU backward timing experiments: B2S_GRU_BWD_OPT bits (b2s_gru.cu) -> PPO iteration time (CUDA-graphed, tools/ppo_iteration.py)
TAG=${1:-r02u}
mkdir -p gpurun_out
: > gpurun_out/${TAG}.txt
for opt in 0 2 4 6 8 16 24 0; do
  B2S_GRU_BWD_OPT=$opt python tools/ppo_iteration.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('opt=$opt', d['value'], d['breakdown_ms'])" >> gpurun_out/${TAG}.txt
done
cat gpurun_out/${TAG}.txt
