#!/bin/bash
TAG=${1:-r02v}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
for i in 1 2; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('humanoid', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'], 'e2e=%.0f' % d['e2e']['value'])"; done
python bench.py --workload kbot --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kbot', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])"
