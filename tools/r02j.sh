#!/bin/bash
TAG=${1:-r02j}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
: > gpurun_out/${TAG}_masks.txt
for mask in 0x7f 0x73 0x71 0x79 0x7b 0x7f; do
  B2S_SYNC_MASK=$mask python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_masks.txt
done
cat gpurun_out/${TAG}_masks.txt
for sv in reference fast; do
python bench.py --workload kbot --kbot-solver $sv --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read());print('kbot $sv',d['value'],d['roofline']['kernel_ms'])"
done
