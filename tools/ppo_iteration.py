"""PPO iteration time (BASELINE.json: "... ; PPO iteration time") for the humanoid walking task, closed loop.

One iteration = ``RLTask._rl_train_loop_step`` (ksim/task/rl.py:1777) as ``ksim_b200.trainer.PPOTrainer`` runs it: a rollout
of T control steps with the random-init policy in the loop, rewards, the on-policy replay of all trajectories,
``num_passes`` x ``num_envs / batch_size`` minibatch steps (replay with gradient, GAE, PPO loss, backward, gradient mean over
ranks, optimiser), and the closing replay for the next model carry (ksim/task/ppo.py:572-694).  Shapes follow
examples/walking.py: actor 43 -> 512 -> 2 x GRU(512) -> MLP -> 17 x 10-component mixture; critic 340 -> 512 -> 2 x GRU(512) ->
MLP -> 1; 4096 envs per GPU, batch 512, 4 passes, rollout 24.

Under torchrun the environments shard across ranks and the flat gradient buffer is all-reduced with NCCL inside the captured
minibatch step.

    python tools/ppo_iteration.py [--envs 4096] [--rollout 24] [--iters 3] [--warmup 2] [--eager]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/ppo_iteration.py
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200.engine import B200Engine  # noqa: E402
from ksim_b200.examples import walking  # noqa: E402
from ksim_b200.trainer import PPOConfig, PPOTrainer  # noqa: E402


def build_trainer(envs: int, rollout: int, batch: int, passes: int, device: torch.device, rank: int, world: int, use_graphs: bool,
                  ) -> PPOTrainer:
    spec = walking.make_spec()
    eng = B200Engine(spec, envs, device=device, randomizers={k: f(spec.model) for k, f in walking.RANDOMIZERS.items()}, seed=0,
                     env_offset=rank * envs, total_envs=world * envs)
    torch.manual_seed(0)  # same initial parameters on every rank
    model = walking.make_model(spec).to(device)
    return PPOTrainer(eng, model, walking.network_columns(eng.program, walking.ACTOR_INPUTS),
                      walking.network_columns(eng.program, walking.CRITIC_INPUTS),
                      PPOConfig(batch_size=batch, num_passes=passes, rollout_length_frames=rollout), use_graphs=use_graphs, seed=rank)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--rollout", type=int, default=24)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--passes", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    tr = build_trainer(args.envs, args.rollout, args.batch, args.passes, dev, rank, world, not args.eager)
    for _ in range(args.warmup):
        tr.iteration()
    if world > 1:
        torch.distributed.barrier()
    runs = [tr.iteration() for _ in range(args.iters)]
    mean = {k: sum(r[k] for r in runs) / len(runs) for k in runs[0]}
    if world > 1:
        tmax = torch.tensor([mean["iteration_ms"]], device=dev)
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        mean["iteration_ms_max_over_ranks"] = float(tmax)
    if rank == 0:
        it_ms = mean.get("iteration_ms_max_over_ranks", mean["iteration_ms"])
        n, t = args.envs, args.rollout
        print(json.dumps({
            "metric": "PPO iteration time (humanoid walking, closed loop)", "value": round(it_ms, 2), "unit": "ms", "higher_is_better": False,
            "n_gpus": world, "env_steps_per_s": round(world * n * t / (it_ms * 1e-3)),
            "breakdown_ms": {k: round(v, 2) for k, v in mean.items() if k.endswith("_ms")}, "final_loss": float(tr.loss),
            "grad_norm": float(tr.flat.grad_norm), "optimizer_steps": int(tr.flat.step_count),
            "config": {"envs_per_gpu": n, "rollout": t, "batch_size": args.batch, "num_passes": args.passes,
                       "updates_per_iteration": args.passes * (n // args.batch),
                       "replays_per_iteration": "2 forward over all envs + one forward/backward per update (ppo.py:597-601, 668-692)",
                       "ours": "engine step, rewards, flags, GAE, PPO loss fwd+bwd, GRU sequence kernels (tcgen05), mixture head, optimiser step",
                       "library": "batched affine maps (cuBLAS bf16 through torch), gathers, NCCL all-reduce of the flat gradient",
                       "launching": "eager" if args.eager else "CUDA graphs: one for the T-step closed-loop rollout, one per minibatch step"},
            "iters": args.iters, "warmup": args.warmup}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
