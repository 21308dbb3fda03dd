"""Where a PPO iteration's time goes: per-kernel device time of ONE eager iteration (torch.profiler / CUPTI), grouped by kernel
name.  Not a benchmark (profiler overhead, eager launches) -- the shares direct the optimisation work.

    python tools/ppo_breakdown.py [--envs 4096] [--rollout 24] [--top 40] [--phase update|rollout|all]
"""

from __future__ import annotations

import argparse
import sys
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

from ppo_iteration import build_trainer  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--rollout", type=int, default=24)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--phase", default="all")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    tr = build_trainer(args.envs, args.rollout, 512, 4, dev, 0, 1, use_graphs=False)
    tr.iteration()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    if args.phase != "all":
        tr.rollout(); tr.score()
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if args.phase == "all":
            tr.iteration()
        elif args.phase == "update":
            tr.update_model()
        else:
            tr.rollout()
        torch.cuda.synchronize()
    tot: dict[str, list[float]] = defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            k = e.name
            tot[k][0] += e.device_time
            tot[k][1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"phase {args.phase}: {total / 1e3:.2f} ms of kernel time in {sum(v[1] for v in tot.values())} launches")
    for name, (us, cnt) in sorted(tot.items(), key=lambda kv: -kv[1][0])[: args.top]:
        print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  x{cnt:<6d} {us / cnt:8.1f} us  {name[:150]}")


if __name__ == "__main__":
    main()
