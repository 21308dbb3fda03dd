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
    # Gumbel-max instead of torch.distributions.Categorical: no host-side argument validation, so the step can be captured
    u = torch.rand_like(logw).clamp_(1e-20, 1.0)
    k = torch.argmax(logw - torch.log(-torch.log(u)), dim=-1, keepdim=True)
    return torch.gather(mean, -1, k)[..., 0] + torch.gather(std, -1, k)[..., 0] * torch.randn_like(k[..., 0], dtype=mean.dtype)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--rollout", type=int, default=64)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--passes", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--graph", action="store_true", help="capture one PPO update (replay, GAE, loss, backward, AdamW) in a CUDA graph "
                    "and replay it per minibatch (single-process runs)")
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
    use_graph = args.graph and world == 1
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=1e-5, capturable=use_graph)
    cfg = ppo.PPOLossConfig()
    a_carry = torch.zeros((DEPTH, n, HIDDEN), device=dev)
    c_carry = torch.zeros((DEPTH, n, HIDDEN), device=dev)

    bsz = args.batch
    assert n % bsz == 0, "the captured update needs equal minibatches"
    static = {"obs_a": torch.zeros((n, t, 43), device=dev), "obs_c": torch.zeros((n, t, 340), device=dev), "act": torch.zeros((n, t, nj), device=dev),
              "keep": torch.zeros((n, t), device=dev), "logp_on": torch.zeros((n, t, nj), device=dev), "values_on": torch.zeros((n, t), device=dev),
              "total": torch.zeros((n, t), device=dev), "dones": torch.zeros((n, t), dtype=torch.uint8, device=dev),
              "succ": torch.zeros((n, t), dtype=torch.uint8, device=dev), "a0": torch.zeros((DEPTH, n, HIDDEN), device=dev),
              "c0": torch.zeros((DEPTH, n, HIDDEN), device=dev), "ix": torch.zeros((bsz,), dtype=torch.long, device=dev),
              "loss": torch.zeros((), device=dev)}

    def update_step() -> None:
        """One PPO update on the minibatch static["ix"] (ppo.py:495-565): every input lives at a fixed address, so the same
        code runs eagerly or as a captured CUDA graph."""
        ix = static["ix"]
        ha, hc = static["a0"][:, ix], static["c0"][:, ix]
        # only the GRU cells are sequential: the input projections, the output MLPs and the mixture log-probabilities run
        # once over all (B, T) rows of the minibatch
        xa, xc = actor.input_proj(static["obs_a"][ix]), critic.input_proj(static["obs_c"][ix])
        keep = static["keep"][ix]
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
            k = keep[:, s][None, :, None]                # recurrent carries restart with the episode
            ha, hc = torch.stack(nh, dim=0) * k, torch.stack(ng, dim=0) * k
        pred = actor.output_proj(torch.stack(tops_a, dim=1))
        logp_off = mixture_log_prob(static["act"][ix], *mixture(pred, nj))
        values_off = critic.output_proj(torch.stack(tops_c, dim=1))[..., 0]
        adv, tgt, _ = eng.gae(values_off.detach().contiguous(), static["total"][ix], static["dones"][ix], static["succ"][ix],
                              0.99, 0.95, normalize_advantages=True)
        loss, _ = ppo.ppo_loss(ppo.PPOInputs(adv, tgt), ppo.PPOVariables(static["logp_on"][ix], static["values_on"][ix]),
                               ppo.PPOVariables(logp_off, values_off), cfg)
        opt.zero_grad(set_to_none=not use_graph)
        loss.backward()
        if world > 1:
            dist_util.allreduce_mean_([q.grad for q in params if q.grad is not None])
        torch.nn.utils.clip_grad_norm_(params, 2.0)
        opt.step()
        static["loss"].copy_(loss.detach())

    roll = {"cur": torch.zeros((n, eng.R), device=dev), "nxt": torch.zeros((n, eng.R), device=dev), "act": torch.zeros((n, nj), device=dev),
            "logp": torch.zeros((n, nj), device=dev), "val": torch.zeros((n,), device=dev)}

    def rollout_step() -> None:
        """One closed-loop control step on fixed buffers (rl.py:1164-1211): observation -> actor / critic -> sampled action ->
        b200sim_step_ctrl; the recurrent carries are updated in place and restart with the episode (rl.py:1100-1123)."""
        row = roll["cur"]
        pred, ac = actor(row[:, a_idx] * a_scale, a_carry)
        mean, std, logw = mixture(pred, nj)
        act = mixture_sample(mean, std, logw)
        roll["logp"].copy_(mixture_log_prob(act, mean, std, logw))
        v, cc = critic(row[:, c_idx] * c_scale, c_carry)
        roll["val"].copy_(v[:, 0])
        roll["act"].copy_(act)
        eng.step(roll["act"], roll["cur"], roll["nxt"])
        keep = (roll["cur"][:, p["TI_R_DONE"]] == 0).float()[None, :, None]
        a_carry.copy_(ac * keep); c_carry.copy_(cc * keep)

    def capture_rollout_step() -> None:
        # the warm-up steps are real steps of real environments; state and carries are put back afterwards
        saved = [x.clone() for x in (eng.state, eng.keys, eng.act_keys, a_carry, c_carry, roll["cur"])]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                rollout_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            rollout_step()
        for dst, src in zip((eng.state, eng.keys, eng.act_keys, a_carry, c_carry, roll["cur"]), saved):
            dst.copy_(src)
        roll["graph"] = g

    def capture() -> None:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                           # warm-up on a side stream: optimiser state, cuBLAS workspaces
                update_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            update_step()
        static["graph"] = g

    def iteration() -> dict[str, float]:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        rec = eng.new_record_buffer(t)
        actions = torch.empty((t, n, nj), device=dev)
        logp_on = torch.empty((n, t, nj), device=dev)
        values_on = torch.empty((n, t), device=dev)
        a0, c0 = a_carry.clone(), c_carry.clone()
        with torch.no_grad():
            roll["cur"].copy_(rec[0])
            for s in range(t):
                if use_graph:
                    if "graph" not in roll:
                        capture_rollout_step()
                    roll["graph"].replay()
                else:
                    rollout_step()
                rec[s].copy_(roll["cur"])               # pre-step block as it was + the post-step block the engine wrote
                actions[s].copy_(roll["act"]); logp_on[:, s].copy_(roll["logp"]); values_on[:, s].copy_(roll["val"])
                roll["cur"].copy_(roll["nxt"])
            rec[t].copy_(roll["nxt"])
        eng.first_rec = rec[t]
        ev[1].record()
        total, _ = eng.rewards(rec, t)
        dones, succ = eng.flags(rec, t)
        ev[2].record()
        data = {"obs_a": (rec[:t, :, a_idx] * a_scale).transpose(0, 1), "obs_c": (rec[:t, :, c_idx] * c_scale).transpose(0, 1),
                "act": actions.transpose(0, 1), "keep": (dones == 0).float(), "logp_on": logp_on, "values_on": values_on, "total": total,
                "dones": dones, "succ": succ, "a0": a0, "c0": c0}
        for k, v in data.items():
            static[k].copy_(v)
        loss_val = 0.0
        for _ in range(args.passes):
            perm = torch.randperm(n, device=dev)
            for b0 in range(0, n, args.batch):
                static["ix"].copy_(perm[b0 : b0 + args.batch])
                if use_graph:
                    if "graph" not in static:
                        capture()
                    static["graph"].replay()
                else:
                    update_step()
                loss_val = static["loss"]
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
                       "ours": "engine step (k_rollout, one launch per control step), rewards, flags, GAE, PPO loss forward+backward",
                       "launching": "CUDA graphs: one per closed-loop control step, one per PPO update" if use_graph else "eager"},
            "iters": args.iters, "warmup": args.warmup}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
