#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
for sv in reference fast; do
python bench.py --workload kbot --kbot-solver $sv --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_kbot_$sv.json 2> gpurun_out/${TAG}_kbot_$sv.err
python -c "import json;d=json.loads(open('gpurun_out/${TAG}_kbot_$sv.json').read().strip().splitlines()[-1]);print('kbot $sv',d['value'],d['roofline']['kernel_ms'])"
done
B2S_SYNC_MASK=0x27f python bench.py --workload kbot --kbot-solver reference --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json;d=json.loads(sys.stdin.read());print('kbot solve_primal',d['value'],d['roofline']['kernel_ms'])"
