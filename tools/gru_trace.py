"""Per-step timeline of the forward GRU kernel's roles in CTA 0 (debug build: B2S_EXTRA_NVCC=-DB2S_GRU_TRACE).

    B2S_EXTRA_NVCC=-DB2S_GRU_TRACE python -c "import __graft_entry__ as g; g.build_cuda()"; python tools/gru_trace.py
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ksim_b200 import binding, replay  # noqa: E402

t_, b_ = 24, 512
dev = torch.device("cuda:0")
stacks = [replay.GRUStack(1).to(dev) for _ in range(2)]
xs = [torch.randn(t_, b_, 512, device=dev).requires_grad_(True) for _ in stacks]
import contextlib
ctx = torch.no_grad() if "--nograd" in sys.argv else contextlib.nullcontext()
with ctx:
    for _ in range(3):
        outs = replay.gru_stacks_forward(stacks, xs, [None, None], None)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 256))()
lib = binding.lib()
lib.b200sim_debug_gru_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.b200sim_debug_gru_trace(buf, 8 * 256) == 0
tr = np.array(buf[: 8 * t_], dtype=np.int64).reshape(t_, 8)
names = ["E:acc_ready", "E:math done", "E:published", "-", "M:chunk0 landed", "M:chunk15 landed", "M:committed", "-"]
t0 = tr[2, 0]
print("cycles relative to step 2's acc_ready; per step: " + ", ".join(names))
for t in range(2, min(t_, 10)):
    print(t, " ".join(f"{int(v - t0):7d}" for v in tr[t]))
d = np.diff(tr[:-1, 0])
print("step period (cycles):", d[2:].mean(), "=", d[2:].mean() / 1.9e3, "us at 1.9 GHz")
seg = tr[3:-2]
print("mean segment lengths (cycles):")
print("  acc_ready(t) -> math done(t)           :", (seg[:, 1] - seg[:, 0]).mean())
print("  -> chunk stored + epilogue barrier      :", (seg[:, 3] - seg[:, 1]).mean())
print("  -> peers_done seen                      :", (seg[:, 7] - seg[:, 3]).mean())
print("  -> proxy fence + multicast issued       :", (seg[:, 2] - seg[:, 7]).mean())
print("  issued(t) -> chunk0 landed(t+1)         :", (tr[4:-1, 4] - seg[:, 2]).mean())
print("  chunk0 -> chunk15 landed (t+1)          :", (tr[4:-1, 5] - tr[4:-1, 4]).mean())
print("  chunk15 landed -> committed (t+1)       :", (tr[4:-1, 6] - tr[4:-1, 5]).mean())
print("  committed(t+1) -> acc_ready(t+1)        :", (tr[4:-1, 0] - tr[4:-1, 6]).mean())

buf2 = (ctypes.c_longlong * (16 * 256))()
lib.b200sim_debug_gru_trace_ch.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.b200sim_debug_gru_trace_ch(buf2, 16 * 256) == 0
tc = np.array(buf2[: 16 * t_], dtype=np.int64).reshape(t_, 16)
print("per-chunk wait-passed times of the MMA issuer, relative to that step's chunk 0 (cycles):")
for t in range(3, 8):
    print(t, " ".join(f"{int(v - tc[t, 0]):6d}" for v in tc[t]))

buf3 = (ctypes.c_longlong * (16 * 256))()
lib.b200sim_debug_gru_trace_ch0.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.b200sim_debug_gru_trace_ch0(buf3, 16 * 256) == 0
te = np.array(buf3[: 16 * t_], dtype=np.int64).reshape(t_, 16)
print("time spent inside each chunk's wait (cycles):")
for t in range(3, 8):
    print(t, " ".join(f"{int(a - b):6d}" for a, b in zip(tc[t], te[t])))

print("MMA issuer: time after each chunk's MMA pair was issued, relative to the half-0 wait passing (cycles):")
for t in range(3, 8):
    print(t, " ".join(f"{int(v - tc[t, 0]):6d}" for v in te[t]))
