#!/bin/bash
# Session measurement (under gpurun): GPU suite + K-Bot line on the current build, barrier-mask and residency sweeps of the
# humanoid kernel, then the stage profile of both kernels on an instrumented build made on the box.
TAG=${1:-r02h}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -1 gpurun_out/${TAG}_gpu_tests.log
python bench.py --workload kbot --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_kbot.json 2> gpurun_out/${TAG}_kbot.err
python -c "import json;d=json.loads(open('gpurun_out/${TAG}_kbot.json').read().strip().splitlines()[-1]);print('kbot',d['value'],d['roofline']['kernel_ms'])"
: > gpurun_out/${TAG}_masks.txt
for mask in 0x7f 0x5f 0x3f 0x6f 0x77 0x7b 0x7d 0x7e 0x1f 0x0; do
  B2S_SYNC_MASK=$mask python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mask=$mask', 'value=%.0f' % d['value'], 'kernel_ms=%.3f' % d['roofline']['kernel_ms'])" >> gpurun_out/${TAG}_masks.txt
done
cat gpurun_out/${TAG}_masks.txt
bash tools/sweep_wpb.sh ${TAG}_wpb "4 7 10 12 14"
B2S_EXTRA_NVCC=-DB2S_PROFILE_STAGES python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > gpurun_out/${TAG}_build.log 2>&1
python tools/stage_profile.py 4096 16 humanoid > gpurun_out/${TAG}_stages_humanoid.txt 2>&1
python tools/stage_profile.py 4096 8 kbot > gpurun_out/${TAG}_stages_kbot.txt 2>&1
cat gpurun_out/${TAG}_stages_humanoid.txt gpurun_out/${TAG}_stages_kbot.txt
