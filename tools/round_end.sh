#!/bin/bash
# Round-end measurement set (run under gpurun, ONE GPU): smoke, the GPU suite, both bench arms for both workloads, the
# scoring-pass rooflines and the ncu captures that profiles/ summarises.  Usage: tools/round_end.sh <tag>
TAG=${1:-r02z}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -1 gpurun_out/${TAG}_gpu_tests.log
python bench.py --impl reference > gpurun_out/${TAG}_ref_hum.json 2> gpurun_out/${TAG}_ref_hum.err
python bench.py > gpurun_out/${TAG}_hum.json 2> gpurun_out/${TAG}_hum.err
python bench.py --workload kbot --impl reference > gpurun_out/${TAG}_ref_kbot.json 2> gpurun_out/${TAG}_ref_kbot.err
python bench.py --workload kbot --no-extras > gpurun_out/${TAG}_kbot.json 2> gpurun_out/${TAG}_kbot.err
python tools/bench_ppo_loss.py > gpurun_out/${TAG}_scoring.jsonl 2> gpurun_out/${TAG}_scoring.err
bash tools/profile.sh ${TAG} humanoid > gpurun_out/${TAG}_profile.log 2>&1
# top kernels of the other rows of the path, one full capture each (each target has exited 0 without ncu in the runs above)
ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 1 -o gpurun_out/${TAG}_score python tools/score_target.py > gpurun_out/${TAG}_ncu_score.log 2>&1
python tools/ppo_iteration.py --eager --iters 1 --warmup 1 > gpurun_out/${TAG}_ppo_eager.json 2> gpurun_out/${TAG}_ppo_eager.err
ncu --set full --clock-control none --import-source on -k regex:k_gru_fwd -s 8 -c 1 -o gpurun_out/${TAG}_gru_fwd python tools/ppo_iteration.py --eager --iters 1 --warmup 1 > gpurun_out/${TAG}_ncu_gru_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gru_bwd -s 4 -c 1 -o gpurun_out/${TAG}_gru_bwd python tools/ppo_iteration.py --eager --iters 1 --warmup 1 > gpurun_out/${TAG}_ncu_gru_bwd.log 2>&1
python - <<PY
import json
for f in ("${TAG}_ref_hum","${TAG}_hum","${TAG}_ref_kbot","${TAG}_kbot"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("kernel_ms"), d.get("clocks"))
        if "ppo_iteration" in d: print(json.dumps(d["ppo_iteration"])[:1500]); print(json.dumps(d.get("secondary_fast_solver")))
    except Exception as e: print(f, "ERR", e)
PY
