"""Development probe (run under gpurun): CUDA vs oracle error growth, and a first rollout timing."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from ksim_b200 import codes as C
from ksim_b200.engine import B200Engine
from ksim_b200.examples.walking import RANDOMIZERS, make_spec
from oracle import OracleEngine

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    level = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    spec = make_spec()
    rz = {k: f(spec.model) for k, f in RANDOMIZERS.items()}
    eng = B200Engine(spec, n, randomizers=rz, seed=0, curriculum_level=level)
    print("variant", eng.info("VARIANT"), "smem/env", eng.info("SMEM_PER_ENV"), "warps/cta", eng.info("WARPS_PER_CTA"))
    init = eng.env_init
    res = {}
    for prec in ("f32", "f64"):
        o = OracleEngine(eng.program, prec)
        st, k, rec0, ak = o.init_envs(init.init_keys, init.levels, init.rand)
        rec = np.zeros((t + 1, n, eng.R), o.real); rec[0] = rec0
        res[prec] = (o, st, k, rec)
        print(prec, "init: state maxdiff", np.abs(eng.state.cpu().numpy() - st).max(), "keys eq",
              (eng.keys.cpu().numpy().view(np.uint32) == k).all(), "rec0 maxdiff", np.abs(eng.first_rec.cpu().numpy() - rec0).max())
    actions = (0.3 * np.random.RandomState(1).randn(t, n, eng.NA)).astype(np.float32)
    rec_g = eng.rollout(torch.from_numpy(actions).cuda())
    torch.cuda.synchronize()
    rec_g = rec_g.cpu().numpy()
    for prec in ("f32", "f64"):
        o, st, k, rec = res[prec]
        o.rollout(actions.astype(o.real), init.rand, st, k, rec)
    ti = eng.program.ti
    nq = spec.model.nq
    qo, do = int(ti[C.TI_R_QPOS]), int(ti[C.TI_R_DONE])
    r32, r64 = res["f32"][3], res["f64"][3]
    for s in range(t):
        print(s, "qpos gpu-f64 %.2e  f32-f64 %.2e  gpu-f32 %.2e" % (
            np.abs(rec_g[s, :, qo:qo + nq] - r64[s, :, qo:qo + nq]).max(), np.abs(r32[s, :, qo:qo + nq] - r64[s, :, qo:qo + nq]).max(),
            np.abs(rec_g[s, :, qo:qo + nq] - r32[s, :, qo:qo + nq]).max()),
            "done gpu/f32/f64", int(rec_g[s, :, do].sum()), int(r32[s, :, do].sum()), int(r64[s, :, do].sum()),
            "eq32", bool((rec_g[s, :, do] == r32[s, :, do]).all()))
    print("keys eq after rollout", (eng.keys.cpu().numpy().view(np.uint32) == res["f32"][2]).all())
    # timing
    for nn, tt in ((4096, 64), (16384, 64)):
        e2 = B200Engine(spec, nn, randomizers=rz, seed=0)
        acts = torch.from_numpy((0.3 * np.random.RandomState(2).randn(tt, nn, e2.NA)).astype(np.float32)).cuda()
        rec = e2.new_record_buffer(tt)
        for _ in range(2):
            e2.rollout(acts, rec)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            e2.rollout(acts, rec)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 3
        print(f"rollout N={nn} T={tt}: {ms:.2f} ms -> {nn*tt/ms*1e3:.3e} env-steps/s; nonfinite envs {int((e2.state[:, int(ti[C.TI_S_NONFINITE])] != 0).sum())}; done frac {float(e2.view(rec[:tt]).done.float().mean()):.4f}")
        tot, comps = e2.rewards(rec, tt)
        d, s = e2.flags(rec, tt)
        adv, tgt, ret = e2.gae(torch.zeros_like(tot), tot, d, s, 0.97, 0.95)
        torch.cuda.synchronize()
        print("  reward mean", float(tot.mean()), "adv mean", float(adv.mean()))

if __name__ == "__main__":
    main()
