"""GPU check of the tcgen05 GRU replay kernels against the fp32 torch restatement (ksim_b200.replay.gru_stack_reference).

    python tools/gru_check.py [--t 24] [--b 512] [--depth 2] [--stacks 2] [--time]

Prints, per output, the max abs error and the error relative to the fp32 result's scale, the device-side status words, and
(with --time) CUDA-event timings of the forward and backward launches.  Exit code 1 when an error exceeds the bf16 budget.
"""

from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200 import replay  # noqa: E402


def rel(a: torch.Tensor, b: torch.Tensor) -> tuple[float, float]:
    d = (a.float() - b.float()).abs().max().item()
    return d, d / max(b.float().abs().max().item(), 1e-12)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=24)
    ap.add_argument("--b", type=int, default=512)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--stacks", type=int, default=2)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tol", type=float, default=3e-2)
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    dev = torch.device("cuda:0")
    t_, b_, hd = args.t, args.b, replay.HIDDEN
    stacks = [replay.GRUStack(args.depth).to(dev) for _ in range(args.stacks)]
    xs = [torch.randn(t_, b_, hd, device=dev).mul_(0.7).requires_grad_(True) for _ in stacks]
    carries = [torch.randn(args.depth, b_, hd, device=dev).mul_(0.5) for _ in stacks]
    keep = (torch.rand(t_, b_, device=dev) > 0.1).float()
    gy = [torch.randn(t_, b_, hd, device=dev) for _ in stacks]

    outs = replay.gru_stacks_forward(stacks, xs, carries, keep)
    torch.cuda.synchronize()
    st_f = replay.last_status(args.stacks, t_, b_, False, dev)
    loss = sum((y.float() * g).sum() for (y, _), g in zip(outs, gy))
    loss.backward()
    torch.cuda.synchronize()
    st_b = replay.last_status(args.stacks, t_, b_, True, dev)
    print(f"status words: forward {st_f}, backward {st_b}")
    got = {"x.grad": [x.grad.clone() for x in xs]}
    for i, s in enumerate(stacks):
        for name, prm in s.named_parameters():
            got[f"s{i}.{name}"] = prm.grad.clone()
            prm.grad = None
    for x in xs:
        x.grad = None

    # fp32 reference: same bf16-rounded weights / inputs would hide the rounding; compare against full fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    refs = [replay.gru_stack_reference(s, x, c, keep) for s, x, c in zip(stacks, xs, carries)]
    rloss = sum((y * g).sum() for (y, _), g in zip(refs, gy))
    rloss.backward()
    bad = False
    for i in range(args.stacks):
        for name, a, b in (("y", outs[i][0], refs[i][0]), ("carry", outs[i][1], refs[i][1]), ("x.grad", got["x.grad"][i], xs[i].grad)):
            d, r = rel(a, b)
            print(f"stack {i} {name:28s} max|d| {d:.3e}  rel-to-scale {r:.3e}")
            bad |= not (r < args.tol)
        for name, prm in stacks[i].named_parameters():
            d, r = rel(got[f"s{i}.{name}"], prm.grad)
            print(f"stack {i} grad {name:23s} max|d| {d:.3e}  rel-to-scale {r:.3e}")
            bad |= not (r < args.tol)
    if args.time:
        for label, fn in (("forward+backward", None),):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for it in range(5):
                if it == 2:
                    ev[0].record()
                o = replay.gru_stacks_forward(stacks, [x.detach().requires_grad_(True) for x in xs], carries, keep)
                if it >= 2 and it == 4:
                    pass
                l2 = sum((y.float() * g).sum() for (y, _), g in zip(o, gy))
                l2.backward()
            ev[1].record()
            torch.cuda.synchronize()
            print(f"{label}: {ev[0].elapsed_time(ev[1]) / 3:.3f} ms per call (T={t_}, B={b_}, depth={args.depth}, stacks={args.stacks}; includes the library GEMMs)")
    print("FAIL" if bad or st_f or st_b else "OK")
    return 1 if bad or st_f or st_b else 0


if __name__ == "__main__":
    sys.exit(main())
