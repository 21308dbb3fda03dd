mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01n_smoke.log 2>&1; tail -1 gpurun_out/r01n_smoke.log
python bench.py --impl reference > gpurun_out/r01n_ref_hum.json 2> gpurun_out/r01n_ref_hum.err
python bench.py > gpurun_out/r01n_hum.json 2> gpurun_out/r01n_hum.err
python bench.py --workload kbot > gpurun_out/r01n_kbot.json 2> gpurun_out/r01n_kbot.err
bash tools/profile.sh r01n humanoid > gpurun_out/r01n_profile.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ppo_loss -s 12 -c 1 -o gpurun_out/r01n_ppo_loss python tools/bench_ppo_loss.py --steps 2 > gpurun_out/ncu_ppo.log 2>&1
python - <<'PY'
import json
for f in ("r01n_ref_hum","r01n_hum","r01n_kbot"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("kernel_ms"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
