#!/bin/bash
TAG=${1:-r02t}
mkdir -p gpurun_out
: > gpurun_out/${TAG}_kbot_masks.txt
for mask in 0x70 0x50 0x30 0x60 0x10 0x40 0x00 0xf0; do
  B2S_SYNC_MASK=$mask python bench.py --workload kbot --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kbot(reference solver) mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_kbot_masks.txt
done
for mask in 0x70 0x50 0x30; do
  B2S_SYNC_MASK=$mask python bench.py --workload kbot --kbot-solver fast --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kbot(fast solver) mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_kbot_masks.txt
done
cat gpurun_out/${TAG}_kbot_masks.txt
