#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_policy.py -m gpu -q -x > gpurun_out/${TAG}_policy_tests.log 2>&1; tail -3 gpurun_out/${TAG}_policy_tests.log
timeout 300 python tools/ppo_iteration.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['breakdown_ms'], d['final_loss'], d['grad_norm'])"
timeout 300 python tools/ppo_breakdown.py --phase update 2>/dev/null | head -12 | cut -c1-120
