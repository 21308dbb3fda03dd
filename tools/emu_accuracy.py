"""One-control-step accuracy of the warp code (host emulation, tests/emu) and of the f32 oracle against the f64 oracle.

    B2S_EMU_SYNC_MASK=0x7f python tools/emu_accuracy.py [humanoid|kbot] [N] [T]

All three sides step from the f64 oracle's state (cast to f32 for the f32 sides), so every comparison is one control
step (5 physics sub-steps) deep.  Test infrastructure; prints |dqvel| statistics.
"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

from ksim_b200.blob import pack_blob
from ksim_b200.envinit import make_env_init
from ksim_b200.program import build_program
from oracle import OracleEngine

workload = sys.argv[1] if len(sys.argv) > 1 else "humanoid"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
t = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if workload == "kbot":
    from ksim_b200.examples.kbot import RANDOMIZERS, make_spec
else:
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
spec = make_spec()
prog = build_program(spec)
rz = {k: f(spec.model) for k, f in RANDOMIZERS.items()}
emu = ctypes.CDLL(str(ROOT / "tests" / "emu" / "libb2s_emu.so"))
p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
init = make_env_init(prog, rz, n, seed=0, level=1.0)
o64, o32 = OracleEngine(prog, "f64"), OracleEngine(prog, "f32")
st64, k64, rec0, _ = o64.init_envs(init.init_keys, init.levels, init.rand)
R = prog.rec_size
actions = (0.3 * np.random.RandomState(1).randn(t, n, prog.action_size)).astype(np.float32)
qv, nv = prog["TI_R_QVEL"], spec.model.nv
err_e, err_o = [], []
rc64, rn64 = np.zeros((n, R)), np.zeros((n, R))
for s in range(t):
    st32, k32 = st64.astype(np.float32), k64.copy()
    ste, ke = st32.copy(), k64.copy()
    rc32, rn32 = np.zeros((n, R), np.float32), np.zeros((n, R), np.float32)
    rce, rne = np.zeros((n, R), np.float32), np.zeros((n, R), np.float32)
    o32.step_ctrl(actions[s], init.rand, st32, k32, rc32, rn32)
    emu.emu_step_ctrl(p(blob), n, p(actions[s]), p(init.rand), p(ste), p(ke), p(rce), p(rne), None)
    o64.step_ctrl(actions[s].astype(np.float64), init.rand.astype(np.float64), st64, k64, rc64, rn64)
    err_e.append(np.abs(rce[:, qv : qv + nv] - rc64[:, qv : qv + nv]))
    err_o.append(np.abs(rc32[:, qv : qv + nv] - rc64[:, qv : qv + nv]))
for name, e in (("emu (warp code)", np.concatenate(err_e).ravel()), ("f32 oracle", np.concatenate(err_o).ravel())):
    print(f"{name:16s} |dqvel| vs f64: mean {e.mean():.3e}  p99 {np.quantile(e, 0.99):.3e}  p99.9 {np.quantile(e, 0.999):.3e}  max {e.max():.3e}")
