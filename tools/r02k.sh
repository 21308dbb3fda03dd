#!/bin/bash
TAG=${1:-r02k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
: > gpurun_out/${TAG}_masks.txt
for mask in 0x7f 0x71 0x70 0x61 0x51 0x31 0x75 0x72 0x74 0x78 0x71; do
  B2S_SYNC_MASK=$mask python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('humanoid mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_masks.txt
done
for mask in 0x7f 0x71 0x73 0x79 0x7b 0x7d 0x75; do
  B2S_SYNC_MASK=$mask python bench.py --workload kbot --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kbot mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_masks.txt
done
cat gpurun_out/${TAG}_masks.txt
python tools/ppo_breakdown.py --phase update > gpurun_out/${TAG}_ppo_update_kernels.txt 2>&1; head -12 gpurun_out/${TAG}_ppo_update_kernels.txt | cut -c1-150
