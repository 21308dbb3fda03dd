"""Small fixed workload for ncu captures: rollout of N envs x T control steps, W warm-up + 1 profiled launch.

    python tools/ncu_target.py [N] [T] [W] [humanoid|kbot]
"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from ksim_b200.engine import B200Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
t = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
workload = sys.argv[4] if len(sys.argv) > 4 else "humanoid"
if workload == "kbot":
    from ksim_b200.examples.kbot import RANDOMIZERS, make_spec
else:
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
spec = make_spec()
eng = B200Engine(spec, n, randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=0,
                 curriculum_level=1.0 if workload == "kbot" else 0.0)
acts = torch.from_numpy((0.3 * np.random.RandomState(2).randn(t, n, eng.NA)).astype(np.float32)).cuda()
rec = eng.new_record_buffer(t)
for _ in range(reps):
    eng.rollout(acts, rec)
tot, comps = eng.rewards(rec, t)
d, s = eng.flags(rec, t)
eng.gae(torch.zeros_like(tot), tot, d, s, 0.97, 0.95, normalize_advantages=True)
torch.cuda.synchronize()
print("ok", float(tot.mean()))
