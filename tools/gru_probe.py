"""Low-level GPU probe of b200sim_gru_forward: one layer through the raw C-ABI, every saved intermediate (r, z, n, hn, h - n)
decoded from the kernel's private layout and compared with the fp32 restatement, with the error broken down by gate and by
hidden-unit chunk (= owning CTA) so that a layout / descriptor mistake shows its pattern in one run.

    python tools/gru_probe.py [--t 1] [--b 128] [--stacks 1]
"""

from __future__ import annotations

import argparse
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200 import replay  # noqa: E402
from ksim_b200.replay import _FwdProblem, _lib, _p  # noqa: E402


def decode_saved(sv: torch.Tensor, t_: int, b_: int) -> torch.Tensor:
    n_mt = (b_ + 127) // 128
    v = sv.view(t_, n_mt, 16, 5, 8, 128, 4).permute(3, 0, 1, 5, 2, 4, 6).reshape(5, t_, n_mt * 128, 512)
    return v[:, :, :b_]


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=1)
    ap.add_argument("--b", type=int, default=128)
    ap.add_argument("--stacks", type=int, default=1)
    args = ap.parse_args()
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    lib = _lib()
    t_, b_, hd, n = args.t, args.b, 512, args.stacks
    w_hh = [torch.empty(3 * hd, hd, device=dev).uniform_(-0.044, 0.044) for _ in range(n)]
    gi = [torch.randn(t_, b_, 3 * hd, device=dev) * 0.5 for _ in range(n)]
    b_i = [torch.randn(3 * hd, device=dev) * 0.1 for _ in range(n)]
    b_hn = [torch.randn(hd, device=dev) * 0.1 for _ in range(n)]
    h0 = [torch.randn(b_, hd, device=dev) for _ in range(n)]
    keep = (torch.rand(t_, b_, device=dev) > 0.2).float()
    ws = torch.zeros(int(lib.b200sim_gru_workspace_bytes(n, t_, b_, 0)), dtype=torch.uint8, device=dev)
    probs = (_FwdProblem * n)()
    out = []
    for s in range(n):
        w16 = w_hh[s].to(torch.bfloat16).contiguous()
        y = torch.zeros(t_, b_, hd, dtype=torch.bfloat16, device=dev)
        hst = torch.zeros(t_, b_, hd, dtype=torch.bfloat16, device=dev)
        hl = torch.zeros(b_, hd, device=dev)
        sv = torch.zeros(int(lib.b200sim_gru_saved_floats(t_, b_)), device=dev)
        p = probs[s]
        p.w_hh, p.gi, p.b_i, p.b_hn, p.h0, p.y, p.hst, p.h_last, p.saved = _p(w16), _p(gi[s]), _p(b_i[s]), _p(b_hn[s]), _p(h0[s]), _p(y), _p(hst), _p(hl), _p(sv)
        out.append((w16, y, hst, hl, sv))
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = lib.b200sim_gru_forward(stream, n, ctypes.byref(probs), t_, b_, _p(keep), _p(ws), ws.numel())
    print("rc", rc, lib.b200sim_gru_last_error())
    torch.cuda.synchronize()
    print("status word", int(ws[:4].view(torch.int32).item()))
    bad = False
    for s in range(n):
        w16, y, hst, hl, sv = out[s]
        w = w16.float()  # the kernel's operand values; isolates layout errors from rounding
        h = h0[s].clone()
        dec = decode_saved(sv, t_, b_)
        for t in range(t_):
            hb = h.to(torch.bfloat16).float()
            gh = hb @ w.T
            g = gi[s][t] + b_i[s]
            r = torch.sigmoid(g[:, :hd] + gh[:, :hd]); z = torch.sigmoid(g[:, hd:2 * hd] + gh[:, hd:2 * hd])
            hn = gh[:, 2 * hd:] + b_hn[s]; nn_ = torch.tanh(g[:, 2 * hd:] + r * hn)
            hp = nn_ + z * (h - nn_)
            for name, got, want in (("r", dec[0, t], r), ("z", dec[1, t], z), ("n", dec[2, t], nn_), ("hn", dec[3, t], hn), ("h-n", dec[4, t], h - nn_),
                                    ("y", y[t].float(), hp), ("hst", hst[t].float(), h)):
                err = (got - want).abs()
                per_chunk = err.view(b_, 16, 32).amax(dim=(0, 2))
                print(f"stack {s} t {t} {name:4s} max err {err.max().item():.3e}  per chunk: " + " ".join(f"{e:.0e}" for e in per_chunk.tolist()))
                bad |= err.max().item() > 2e-2
            if t == 0 and err.max().item() > 2e-2:
                e = (dec[3, t] - hn).abs()
                print("  hn err by row block of 8:", " ".join(f"{x:.0e}" for x in e.view(-1, 8, hd).amax(dim=(1, 2))[:16].tolist()))
            h = hp * keep[t][:, None]
        err = (hl - h).abs().max().item()
        print(f"stack {s} carry max err {err:.3e}")
        bad |= err > 2e-2
    print("FAIL" if bad else "OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
