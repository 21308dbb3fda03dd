"""Finds a workspace field that is read before it is written (-DB2S_POISON_SMEM build): one humanoid control step through a
chosen instantiation with bytes [lo, hi) of every workspace pre-filled with 0xFF, bisected on the outputs.

    B2S_EXTRA_NVCC=-DB2S_POISON_SMEM python -c "import __graft_entry__ as g; g.build_cuda(force=True)"
    python tools/debug_poison.py [variant] [level] [humanoid|kbot]
"""
import os, subprocess, sys
import numpy as np
variant = sys.argv[1] if len(sys.argv) > 1 else "2"
level = sys.argv[2] if len(sys.argv) > 2 else "1.0"
workload = sys.argv[3] if len(sys.argv) > 3 else "humanoid"
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ksim_b200.engine import B200Engine
if sys.argv[3] == "kbot":
    from ksim_b200.examples.kbot import RANDOMIZERS, make_spec
else:
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec
spec = make_spec()
n = 40
eng = B200Engine(spec, n, device="cuda:0", randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=3, curriculum_level=float(sys.argv[2]))
a = torch.from_numpy((0.3 * np.random.RandomState(4).randn(3, n, eng.NA)).astype(np.float32)).cuda()
cur, nxt = (torch.zeros((n, eng.R), device="cuda") for _ in range(2))
outs = []
for s in range(3):
    eng.step(a[s], cur, nxt)
    outs += [cur.cpu().numpy().copy(), nxt.cpu().numpy().copy()]
torch.cuda.synchronize()
np.save(sys.argv[1], np.concatenate(outs + [eng.state.cpu().numpy()], axis=1))
print(eng.info("SMEM_PER_ENV"))
'''
def run(lo, hi, tag):
    f = f"/tmp/poison_{tag}.npy"
    env = dict(os.environ, B2S_MIN_VARIANT=variant, B2S_POISON_LO=str(lo), B2S_POISON_HI=str(hi))
    r = subprocess.run([sys.executable, "-c", code, f, level, workload], env=env, capture_output=True, text=True)
    if r.returncode:
        print("run failed", r.stderr.strip().splitlines()[-1][-200:])
        return None, 0
    return np.load(f), int(r.stdout.strip().splitlines()[-1])
base, size = run(0, 0, "base")
def differs(lo, hi):
    x, _ = run(lo, hi, "x")
    return x is None or not np.array_equal(x, base, equal_nan=True)
whole = differs(0, size)
print("workspace bytes", size, "whole-workspace poison changes the outputs:", whole)
if not whole:
    print("clean: no workspace word is read before it is written on this path")
    sys.exit(0)
lo, hi = 0, size
while hi - lo > 4:
    mid = (lo + hi) // 2 // 4 * 4
    if differs(lo, mid): hi = mid
    elif differs(mid, hi): lo = mid
    else:
        print("only the union of", (lo, mid), (mid, hi), "matters; stopping"); break
    print("range", lo, hi, flush=True)
if hi - lo <= 4:
    print("first read-before-write word at byte offset", lo, "of Work<D>")
