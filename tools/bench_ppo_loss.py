"""Times b200sim_ppo_loss (k_ppo_loss + k_ppo_reduce) and the scoring passes that follow a rollout (b200sim_rewards,
_extract_flags, _gae, _metrics; humanoid, 4096 envs x 64) against the measured copy bandwidth: one JSON line per shape with the algorithmic bytes, the CUDA-event time on the launch stream and
the roofline fraction.  Inputs smaller than 2x L2 get an untimed 256 MiB flush between iterations.

    python tools/bench_ppo_loss.py [--steps 20] [--warmup 3]
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200 import ppo  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("hbm_gbs", 6546.2))
    dev = torch.device("cuda:0")
    flush_buf = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)
    for b, t, a, with_entropy in [(4096, 64, 17, True), (8192, 100, 20, True), (32768, 64, 17, True), (32768, 64, 17, False)]:
        g = torch.Generator(device=dev).manual_seed(0)
        lp_on = -1.0 + 0.5 * torch.randn(b, t, a, device=dev, generator=g)
        lp_off = lp_on + 0.15 * torch.randn(b, t, a, device=dev, generator=g)
        ent = 1.0 + 0.2 * torch.randn(b, t, a, device=dev, generator=g) if with_entropy else None
        v_on, adv = torch.randn(b, t, device=dev, generator=g), torch.randn(b, t, device=dev, generator=g)
        v_off, tgt = v_on + 0.3 * torch.randn(b, t, device=dev, generator=g), v_on + 0.5 * torch.randn(b, t, device=dev, generator=g)
        inputs, on, off = ppo.PPOInputs(adv, tgt), ppo.PPOVariables(lp_on, v_on), ppo.PPOVariables(lp_off, v_off, ent)
        n = b * t
        # algorithmic bytes: 2A (+A) + 4 floats read, 4 loss terms + 2 gradients written, per (b, t)
        nbytes = 4 * n * ((3 if with_entropy else 2) * a + 4 + 6)
        flush = nbytes < 2 * 126 * 2**20
        for _ in range(args.warmup):
            ppo.ppo_loss_and_grads(inputs, on, off)
        torch.cuda.synchronize()
        total_ms = 0.0
        for i in range(args.steps):
            if flush:
                flush_buf.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ppo.ppo_loss_and_grads(inputs, on, off)
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        ms = total_ms / args.steps
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "k_ppo_loss+k_ppo_reduce", "shape": [b, t, a], "entropy": with_entropy, "ms": round(ms, 4),
                          "algorithmic_bytes": nbytes, "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 4),
                          "l2": "flushed between iterations" if flush else "inputs exceed 2x L2",
                          "elements_per_s": round(n / (ms * 1e-3))}), flush=True)
    scoring(args, peak)


def scoring(args: argparse.Namespace, peak: float) -> None:
    """The HBM-bound passes after the rollout, each alone, on a 4096 x 64 humanoid record buffer (634 MiB > 2x L2)."""
    import numpy as np

    from ksim_b200.engine import B200Engine
    from ksim_b200.examples.walking import RANDOMIZERS, make_spec

    spec = make_spec()
    n, t = 4096, 64
    eng = B200Engine(spec, n, randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=0)
    acts = torch.from_numpy((0.3 * np.random.RandomState(2).randn(t, n, eng.NA)).astype(np.float32)).cuda()
    rec = eng.rollout(acts)
    total, comps = eng.rewards(rec, t)
    dones, succ = eng.flags(rec, t)
    values = torch.zeros_like(total)
    k = int(comps.shape[0])
    cases = {
        "k_score (rewards + flags + GAE, one launch)": (lambda: eng.score(rec, t, values=values, gamma=0.97, lam=0.95),
                                                        n * t * (4 * (eng.R + 1 + k) + 2 + 16)),
        "k_score (rewards only)": (lambda: eng.rewards(rec, t), 4 * n * t * (eng.R + 1 + k)),
        "k_extract_flags": (lambda: eng.flags(rec, t), n * t * (8 + 2)),
        "k_gae": (lambda: eng.gae(values, total, dones, succ, 0.97, 0.95), n * t * 22),
        "k_gae+k_gae_normalize": (lambda: eng.gae(values, total, dones, succ, 0.97, 0.95, normalize_advantages=True), n * t * (22 + 12)),
        "k_metrics_env+k_metrics_reduce": (lambda: eng.metrics(rec, total, comps, t), 4 * n * t * (10 + 1 + k)),
    }
    flush_buf = torch.empty(256 * 2**20, dtype=torch.uint8, device="cuda")
    for name, (fn, nbytes) in cases.items():
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        total_ms = 0.0
        for i in range(args.steps):
            flush_buf.fill_(i & 0xff)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        ms = total_ms / args.steps
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "shape": [n, t], "ms": round(ms, 4), "algorithmic_bytes": nbytes, "achieved_gbs": round(gbs, 1),
                          "peak_gbs": peak, "frac": round(gbs / peak, 4), "l2": "flushed between iterations"}), flush=True)


if __name__ == "__main__":
    main()
