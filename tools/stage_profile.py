"""Per-stage cycle shares of the rollout kernel (needs a library built with -DB2S_PROFILE_STAGES, see __graft_entry__.build_cuda).

    B2S_EXTRA_NVCC=-DB2S_PROFILE_STAGES python -c "import __graft_entry__ as g; g.build_cuda(force=True)"
    gpurun -- python tools/stage_profile.py [N] [T] [humanoid|kbot]
"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from ksim_b200 import binding
from ksim_b200.engine import B200Engine

NAMES = ["kinematics", "com_pos+crb", "collision+com_vel+constraint", "passive+rne+actuation", "acceleration (M solve, Z)", "solve",
         "post-solve barrier", "sensors_acc", "implicit solve+integrate", "ctrl/actuators/events/rng", "action load", "terminations/record/reset/commands",
         "observe", " solve path: primal (dof space)", " solve path: dual (shared memory)", "(engine_step total)", " solve: setup (dual_setup | newton_dof start point)", " solve: start-point costs | newton_dof gradient", " solve: dual_direction | newton_dof refactorisation",
         " solve: convergence check + vote | newton_dof substitution-only solve", " solve: line search (newton_dof: + direction, jv)", " solve: step update + row costs",
         " solve path: dual8 (registers) cycles | newton_dof ITERATION COUNT", " newton_dof REFACTORISATION COUNT",
         " kin: A0 sin/cos", " kin: A local transforms", " kin: B level compose", " kin: C matrices/anchors/sites", "", "", "", ""]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t = int(sys.argv[2]) if len(sys.argv) > 2 else 16
workload = sys.argv[3] if len(sys.argv) > 3 else "humanoid"
if workload == "kbot":
    from ksim_b200.examples.kbot import RANDOMIZERS, make_spec
else:
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
spec = make_spec()
eng = B200Engine(spec, n, randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=0,
                 curriculum_level=1.0 if workload == "kbot" else 0.0)
acts = torch.from_numpy((0.3 * np.random.RandomState(2).randn(t, n, eng.NA)).astype(np.float32)).cuda()
rec = eng.new_record_buffer(t)
lib = binding.lib()
for _ in range(3):
    eng.rollout(acts, rec)
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
lib.b200sim_debug_stage_cycles(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.rollout(acts, rec)
e1.record()
torch.cuda.synchronize()
lib.b200sim_debug_stage_cycles(out, 0)
v = np.array(list(out), dtype=np.float64)
tot = v[:13].sum()
print(f"rollout {n} x {t}: {e0.elapsed_time(e1):.2f} ms (instrumented build); warp-cycles per env-step: {tot / (n * t):.0f}")
for name, c in zip(NAMES[:28], v[:28]):
    if c:
        print(f"{100 * c / tot:6.2f}%  {c / (n * t):10.0f} cyc/env-step  {name}")
nsolve = v[28:32].sum()
for name, c in zip(["no rows", "dual8 (registers)", "dual (shared memory)", "primal (dof space)"], v[28:32]):
    print(f"{100 * c / max(nsolve, 1):6.2f}%  of sub-step solves take the path: {name}")
