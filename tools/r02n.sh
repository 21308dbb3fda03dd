#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log
python tools/ppo_iteration.py > gpurun_out/${TAG}_ppo_iter.json 2> gpurun_out/${TAG}_ppo_iter.err; cut -c1-400 gpurun_out/${TAG}_ppo_iter.json
python tools/ppo_breakdown.py --phase update 2>/dev/null | head -6 | cut -c1-140
