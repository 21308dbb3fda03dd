#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
B2S_EXTRA_NVCC=-DB2S_PROFILE_STAGES python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > gpurun_out/${TAG}_build.log 2>&1
python tools/stage_profile.py 4096 8 kbot > gpurun_out/${TAG}_stages_kbot.txt 2>&1
python tools/stage_profile.py 4096 16 humanoid > gpurun_out/${TAG}_stages_humanoid.txt 2>&1
cat gpurun_out/${TAG}_stages_kbot.txt gpurun_out/${TAG}_stages_humanoid.txt
