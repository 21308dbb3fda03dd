#!/bin/bash
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_policy.py -m gpu -q -x > gpurun_out/${TAG}_policy_tests.log 2>&1; tail -3 gpurun_out/${TAG}_policy_tests.log
for mode in pair single; do
  if [ $mode = single ]; then export B2S_GRU_BWD_SINGLE=1; else unset B2S_GRU_BWD_SINGLE; fi
  timeout 300 python tools/ppo_iteration.py 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['value'], d['breakdown_ms'], d['final_loss'])"
done
unset B2S_GRU_BWD_SINGLE
timeout 300 python tools/ppo_breakdown.py --phase update 2>/dev/null | head -4 | cut -c1-120
