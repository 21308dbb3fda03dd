"""One humanoid rollout (4096 x 64) followed by a few b200sim_score launches: the ncu target of the scoring kernel.

    ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 1 -o gpurun_out/<tag>_score python tools/score_target.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from ksim_b200.engine import B200Engine

workload = sys.argv[1] if len(sys.argv) > 1 else "humanoid"
if workload == "kbot":
    from ksim_b200.examples.kbot import RANDOMIZERS, make_spec
    n, t = 8192, 100
else:
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
    n, t = 4096, 64
spec = make_spec()
eng = B200Engine(spec, n, randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=0)
acts = torch.from_numpy((0.3 * np.random.RandomState(2).randn(t, n, eng.NA)).astype(np.float32)).cuda()
rec = eng.rollout(acts)
values = torch.zeros((n, t), device="cuda")
for _ in range(4):
    eng.score(rec, t, values=values, gamma=0.97, lam=0.95)
torch.cuda.synchronize()
print("ok")
