#!/bin/bash
# Residency experiment (run under gpurun): humanoid rollout throughput vs warps (= environments) per SM, with the grid held at
# exactly two waves (envs = 2 x 148 x wpb) so that only the residency changes.  Usage: tools/sweep_wpb.sh <tag> "<wpb list>"
TAG=${1:-wpb}
mkdir -p gpurun_out
: > gpurun_out/${TAG}.txt
for wpb in ${2:-7 10 12 14}; do
  envs=$((296 * wpb))
  B2S_WPB=$wpb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --envs $envs --rollout 16 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wpb=$wpb envs=$envs', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}.txt
done
cat gpurun_out/${TAG}.txt
