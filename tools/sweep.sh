#!/bin/bash
# Tuning sweep (run under gpurun): rollout-kernel time for combinations of B2S_WPB / B2S_SYNC_MASK.
# Usage: tools/sweep.sh <tag> "<wpb list>" "<mask list>"
TAG=${1:-sweep}
mkdir -p gpurun_out
: > gpurun_out/${TAG}.txt
for wpb in ${2:-14}; do
  for mask in ${3:-0x7f}; do
    B2S_WPB=$wpb B2S_SYNC_MASK=$mask python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wpb=$wpb mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.2f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}.txt
  done
done
cat gpurun_out/${TAG}.txt
