"""One humanoid control step from the same state through two kernel instantiations (B2S_MIN_VARIANT), field by field."""
import os, subprocess, sys
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ksim_b200.engine import B200Engine
from ksim_b200.examples.walking import RANDOMIZERS, make_spec
spec = make_spec()
n = 40
eng = B200Engine(spec, n, device="cuda:0", randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=3, curriculum_level=1.0)
a = torch.from_numpy((0.3 * np.random.RandomState(4).randn(n, eng.NA)).astype(np.float32)).cuda()
cur, nxt = (torch.zeros((n, eng.R), device="cuda") for _ in range(2))
eng.step(a, cur, nxt)
torch.cuda.synchronize()
np.save(sys.argv[1], np.concatenate([cur.cpu().numpy(), eng.state.cpu().numpy()], axis=1))
print(eng.info("VARIANT"), {k: eng.program[k] for k in ("TI_R_QPOS", "TI_R_QVEL", "TI_R_XPOS", "TI_R_XQUAT", "TI_R_CTRL", "TI_R_ACTION", "TI_R_TERM")}, eng.R)
'''
runs = [("v1", dict(B2S_MIN_VARIANT="1")), ("v2", dict(B2S_MIN_VARIANT="2")), ("v2_ref_solver", dict(B2S_MIN_VARIANT="2", B2S_SYNC_MASK="0x17f")), ("v1_ref_solver", dict(B2S_MIN_VARIANT="1", B2S_SYNC_MASK="0x17f")),
        ("v2_votes", dict(B2S_MIN_VARIANT="2", B2S_SYNC_MASK="0xff")), ("v0", dict(B2S_MIN_VARIANT="0"))]
import numpy as np
out = {}
for name, env in runs:
    f = f"/tmp/dv_{name}.npy"
    r = subprocess.run([sys.executable, "-c", code, f], env=dict(os.environ, **env), capture_output=True, text=True)
    print(name, r.returncode, r.stdout.strip()[-300:], r.stderr.strip().splitlines()[-1][-100:] if r.returncode else "")
    if r.returncode == 0:
        out[name] = np.load(f)
ref = out["v1"]
for name, x in out.items():
    d = np.abs(x - ref)
    cols = np.argsort(-d.max(axis=0))[:12]
    print(name, "max diff", d.max(), "cols", [(int(c), float(d[:, c].max())) for c in cols])
d = np.abs(out["v2"] - out["v1"])
print("rows (envs) that differ:", [(int(e), float(d[e].max()), int(np.argmax(d[e]))) for e in np.nonzero(d.max(axis=1) > 1e-3)[0]])
