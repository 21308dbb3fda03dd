"""PPO iteration time (BASELINE.json: "... ; PPO iteration time") for the humanoid walking task, closed loop.

One iteration = what ``RLTask._rl_train_loop_step`` (ksim/task/rl.py:1777) does: a rollout of T control steps with the
random-init policy in the loop (observation -> actor -> sampled action -> engine step), rewards, then ``num_passes`` x
``num_envs / batch_size`` PPO updates, each replaying the recurrent networks over the minibatch's T steps, computing GAE from
the new values (``compute_ppo_inputs``), the PPO loss and its gradient, a gradient all-reduce across ranks and an Adam step
(ksim/task/ppo.py:417-694).  Shapes follow examples/walking.py: actor 43 -> 512 -> 2 x GRU(512) -> MLP -> 17 x 10-component
mixture of Gaussians; critic 340 -> 512 -> 2 x GRU(512) -> MLP -> 1; 4096 envs per GPU, batch 512, 4 passes.

What runs where (and what this number is NOT): the engine step, rewards, flags, GAE and the PPO loss + its backward are this
repo's CUDA kernels; the policy / critic networks, the mixture log-probabilities and Adam are plain PyTorch (cuBLAS, eager,
TF32) -- library code standing in for the host framework (JAX / equinox / optax in the reference), not a product of this repo.
Under torchrun the environments shard across ranks and the gradients are averaged with NCCL (ksim_b200.distributed).

    python tools/ppo_iteration.py [--envs 4096] [--rollout 64] [--iters 3] [--warmup 1]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/ppo_iteration.py
"""

from __future__ import annotations

import argparse
import json
import math
import os
import sys
from pathlib import Path

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ksim_b200 import distributed as dist_util, ppo  # noqa: E402
from ksim_b200.engine import B200Engine  # noqa: E402
from ksim_b200.examples.walking import RANDOMIZERS, make_spec  # noqa: E402

NUM_MIXTURES, HIDDEN, DEPTH = 10, 512, 2


class Recurrent(nn.Module):
    """input_proj -> DEPTH GRU cells -> MLP with two hidden layers (examples/walking.py:40-205)."""

    def __init__(self, num_inputs: int, num_outputs: int, act: type[nn.Module]) -> None:
        super().__init__()
        self.input_proj = nn.Linear(num_inputs, HIDDEN)
        self.rnns = nn.ModuleList([nn.GRUCell(HIDDEN, HIDDEN) for _ in range(DEPTH)])
        self.output_proj = nn.Sequential(nn.Linear(HIDDEN, HIDDEN), act(), nn.Linear(HIDDEN, HIDDEN), act(), nn.Linear(HIDDEN, num_outputs, bias=False))

    def forward(self, x: torch.Tensor, carry: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        x = self.input_proj(x)
        out = []
        for i, rnn in enumerate(self.rnns):
            x = rnn(x, carry[i])
            out.append(x)
        return self.output_proj(x), torch.stack(out, dim=0)


def mixture(pred: torch.Tensor, nj: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Actor head -> (mean, std, log-weights), each [..., nj, M] (examples/walking.py:130-141)."""
    m = nj * NUM_MIXTURES
    shape = pred.shape[:-1] + (nj, NUM_MIXTURES)
    mean, std, logits = pred[..., :m].reshape(shape), pred[..., m : 2 * m].reshape(shape), pred[..., 2 * m :].reshape(shape)
    std = ((F.softplus(std) + 0.01) * 0.5).clamp(max=1.0)
    return mean, std, torch.log_softmax(logits, dim=-1)


def mixture_log_prob(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, logw: torch.Tensor) -> torch.Tensor:
    comp = -0.5 * ((x[..., None] - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
    return torch.logsumexp(logw + comp, dim=-1)


def mixture_sample(mean: torch.Tensor, std: torch.Tensor, logw: torch.Tensor) -> torch.Tensor:
    k = torch.distributions.Categorical(logits=logw).sample()[..., None]
    return torch.gather(mean, -1, k)[..., 0] + torch.gather(std, -1, k)[..., 0] * torch.randn_like(k[..., 0], dtype=mean.dtype)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--rollout", type=int, default=64)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--passes", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)  # same initial parameters on every rank

    spec = make_spec()
    n, t = args.envs, args.rollout
    eng = B200Engine(spec, n, device=dev, randomizers={k: f(spec.model) for k, f in RANDOMIZERS.items()}, seed=0,
                     env_offset=rank * n, total_envs=world * n)
    p = eng.program
    nj = eng.NA
    ob, cm = p["TI_R_OBS"], p["TI_R_CMD"]

    def cols(names_scales: list[tuple[str, float]], cmd_fields: list[int]) -> tuple[torch.Tensor, torch.Tensor]:
        idx, scale = [], []
        for name, s in names_scales:
            off, dim, _ = p.obs_slices[name]
            idx += list(range(ob + off, ob + off + dim)); scale += [s] * dim
        idx += [cm + c for c in cmd_fields]; scale += [1.0] * len(cmd_fields)
        return torch.tensor(idx, device=dev), torch.tensor(scale, device=dev)

    # examples/walking.py:585-615 (actor) and :617-670 (critic); command columns: linvel (vel, yaw, xvel, yvel), angvel vel
    a_idx, a_scale = cols([("noisy_joint_position", 1.0), ("noisy_joint_velocity", 0.1), ("noisy_imu_projected_gravity", 1.0),
                           ("noisy_imu_gyro", 1.0)], [0, 1, 6])
    c_idx, c_scale = cols([("joint_position", 1.0), ("joint_velocity", 0.1), ("center_of_mass_inertia", 1.0), ("center_of_mass_velocity", 1.0),
                           ("imu_acc", 1.0), ("imu_gyro", 1.0), ("projected_gravity", 1.0), ("actuator_force", 0.01), ("base_position", 1.0),
                           ("base_orientation", 1.0), ("base_linear_velocity", 1.0), ("base_angular_velocity", 1.0), ("feet_force", 1.0)],
                          [0, 1, 2, 3, 6])
    assert a_idx.numel() == 43 and c_idx.numel() == 340, (a_idx.numel(), c_idx.numel())
    actor = Recurrent(43, nj * 3 * NUM_MIXTURES, nn.GELU).to(dev)
    critic = Recurrent(340, 1, nn.ELU).to(dev)
    params = list(actor.parameters()) + list(critic.parameters())
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=1e-5)
    cfg = ppo.PPOLossConfig()
    a_carry = torch.zeros((DEPTH, n, HIDDEN), device=dev)
    c_carry = torch.zeros((DEPTH, n, HIDDEN), device=dev)

    def iteration() -> dict[str, float]:
        nonlocal a_carry, c_carry
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        rec = eng.new_record_buffer(t)
        actions = torch.empty((t, n, nj), device=dev)
        logp_on = torch.empty((n, t, nj), device=dev)
        values_on = torch.empty((n, t), device=dev)
        a0, c0 = a_carry, c_carry
        with torch.no_grad():
            for s in range(t):
                row = rec[s]
                pred, a_carry = actor(row[:, a_idx] * a_scale, a_carry)
                mean, std, logw = mixture(pred, nj)
                act = mixture_sample(mean, std, logw)
                logp_on[:, s] = mixture_log_prob(act, mean, std, logw)
                v, c_carry = critic(row[:, c_idx] * c_scale, c_carry)
                values_on[:, s] = v[:, 0]
                actions[s] = act
                eng.step(act, rec[s], rec[s + 1])
                done = rec[s, :, p["TI_R_DONE"]] != 0      # recurrent carries restart with the episode (rl.py:1100-1123)
                a_carry = torch.where(done[None, :, None], torch.zeros_like(a_carry), a_carry)
                c_carry = torch.where(done[None, :, None], torch.zeros_like(c_carry), c_carry)
        eng.first_rec = rec[t]
        ev[1].record()
        total, _ = eng.rewards(rec, t)
        dones, succ = eng.flags(rec, t)
        ev[2].record()
        obs_a = (rec[:t, :, a_idx] * a_scale).transpose(0, 1).contiguous()   # [N, T, 43]
        obs_c = (rec[:t, :, c_idx] * c_scale).transpose(0, 1).contiguous()
        act_nt = actions.transpose(0, 1).contiguous()
        done_nt = dones.bool()
        loss_val = 0.0
        for _ in range(args.passes):
            perm = torch.randperm(n, device=dev)
            for b0 in range(0, n, args.batch):
                ix = perm[b0 : b0 + args.batch]
                ha, hc = a0[:, ix], c0[:, ix]
                # only the GRU cells are sequential: the input projections, the output MLPs and the mixture log-probabilities
                # run once over all (B, T) rows of the minibatch
                xa, xc = actor.input_proj(obs_a[ix]), critic.input_proj(obs_c[ix])
                tops_a, tops_c = [], []
                for s in range(t):
                    h, nh = xa[:, s], []
                    for i, rnn in enumerate(actor.rnns):
                        h = rnn(h, ha[i]); nh.append(h)
                    tops_a.append(h)
                    g, ng = xc[:, s], []
                    for i, rnn in enumerate(critic.rnns):
                        g = rnn(g, hc[i]); ng.append(g)
                    tops_c.append(g)
                    keep = (~done_nt[ix, s])[None, :, None].to(h.dtype)
                    ha, hc = torch.stack(nh, dim=0) * keep, torch.stack(ng, dim=0) * keep
                pred = actor.output_proj(torch.stack(tops_a, dim=1))
                logp_off = mixture_log_prob(act_nt[ix], *mixture(pred, nj))
                values_off = critic.output_proj(torch.stack(tops_c, dim=1))[..., 0]
                adv, tgt, _ = eng.gae(values_off.detach().contiguous(), total[ix].contiguous(), dones[ix].contiguous(), succ[ix].contiguous(),
                                      0.99, 0.95, normalize_advantages=True)
                loss, _ = ppo.ppo_loss(ppo.PPOInputs(adv, tgt), ppo.PPOVariables(logp_on[ix].contiguous(), values_on[ix].contiguous()),
                                       ppo.PPOVariables(logp_off, values_off), cfg)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                if world > 1:
                    dist_util.allreduce_mean_([q.grad for q in params if q.grad is not None])
                torch.nn.utils.clip_grad_norm_(params, 2.0)
                opt.step()
                loss_val = loss
        ev[3].record()
        torch.cuda.synchronize()
        return {"rollout_ms": ev[0].elapsed_time(ev[1]), "scoring_ms": ev[1].elapsed_time(ev[2]), "update_ms": ev[2].elapsed_time(ev[3]),
                "iteration_ms": ev[0].elapsed_time(ev[3]), "loss": float(loss_val.detach()) if torch.is_tensor(loss_val) else float(loss_val)}

    for _ in range(args.warmup):
        iteration()
    runs = [iteration() for _ in range(args.iters)]
    mean = {k: sum(r[k] for r in runs) / len(runs) for k in runs[0]}
    if world > 1:
        tmax = torch.tensor([mean["iteration_ms"]], device=dev)
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        mean["iteration_ms_max_over_ranks"] = float(tmax)
    if rank == 0:
        it_ms = mean.get("iteration_ms_max_over_ranks", mean["iteration_ms"])
        print(json.dumps({
            "metric": "PPO iteration time (humanoid walking, closed loop)", "value": round(it_ms, 2), "unit": "ms", "higher_is_better": False,
            "n_gpus": world, "env_steps_per_s": round(world * n * t / (it_ms * 1e-3)),
            "breakdown_ms": {k: round(v, 2) for k, v in mean.items() if k.endswith("_ms")}, "final_loss": mean["loss"],
            "config": {"envs_per_gpu": n, "rollout": t, "batch_size": args.batch, "num_passes": args.passes,
                       "updates_per_iteration": args.passes * math.ceil(n / args.batch),
                       "networks": "PyTorch eager TF32 (library code): GRU actor 43->512x2->510 (10-mixture), GRU critic 340->512x2->1",
                       "ours": "engine step (k_rollout, one launch per control step), rewards, flags, GAE, PPO loss forward+backward"},
            "iters": args.iters, "warmup": args.warmup}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
