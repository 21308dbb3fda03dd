"""One PPO iteration on one rank: closed-loop rollout, scoring, and the model update (``BASELINE.json``: "PPO iteration time").

Mirrors ``RLTask._rl_train_loop_step`` (ksim/task/rl.py:1777-1817): ``_single_unroll`` (rl.py:1604-1665, the policy in the
loop: observation -> ``sample_action`` -> engine step, T times), ``get_rewards``, then ``PPOTask.update_model``
(ksim/task/ppo.py:572-694):

1. the on-policy variables are a replay of the *current* model over all N trajectories (ppo.py:597-601);
2. ``num_passes`` x ``N / batch_size`` minibatch steps (``_single_step``, ppo.py:495-565): replay with gradient
   (``get_ppo_variables``), GAE from the new values (``compute_ppo_inputs``), the PPO loss, its gradient, [mean over ranks],
   the optax chain;
3. a last replay with the updated model provides the next rollout's model carry (ppo.py:668-692).

Everything sits at fixed addresses so that the T-step rollout and one minibatch step are each captured once in a CUDA graph
(``use_graphs``); at world size > 1 the gradient mean is one in-place NCCL all-reduce of the flat gradient buffer INSIDE the
captured step (``ksim_b200.optim.FlatParams``).  What is this repo's CUDA: the engine step, rewards, flags, GAE, PPO loss
forward + backward, the GRU sequence kernels (tcgen05), the mixture head (log-prob, backward, sampling) and the optimiser
step.  Library code: the batched affine maps (cuBLAS bf16 GEMMs through torch), index gathers, NCCL.
"""

from __future__ import annotations

__all__ = ["PPOConfig", "PPOTrainer"]

from dataclasses import dataclass, field
from typing import Sequence

import torch

from . import ppo
from .engine import B200Engine
from .optim import FlatParams, OptimizerConfig
from .policy import MixtureOfGaussians, Model, get_ppo_variables
from .replay import gru_stacks_forward


@dataclass(frozen=True)
class PPOConfig:
    """The fields of ``PPOConfig`` / ``RLConfig`` an iteration depends on (ppo.py:251-306; walking.py:772-788 overrides)."""

    batch_size: int = 512
    num_passes: int = 4
    rollout_length_frames: int = 24
    gamma: float = 0.99
    lam: float = 0.95
    normalize_advantages: bool = True
    monte_carlo_returns: bool = False
    loss: ppo.PPOLossConfig = field(default_factory=ppo.PPOLossConfig)
    optimizer: OptimizerConfig = field(default_factory=OptimizerConfig)


class PPOTrainer:
    def __init__(self, engine: B200Engine, model: Model, actor_columns: tuple[Sequence[int], Sequence[float]],
                 critic_columns: tuple[Sequence[int], Sequence[float]], config: PPOConfig = PPOConfig(), use_graphs: bool = True,
                 seed: int = 0) -> None:
        self.eng, self.model, self.config, self.use_graphs = engine, model, config, use_graphs
        dev = engine.device
        self.device = dev
        n, t = engine.num_envs, config.rollout_length_frames
        if n % config.batch_size:
            raise ValueError("num_envs must be a multiple of batch_size (ppo.py:640 reshapes the permutation)")
        self.n, self.t = n, t
        self.flat = FlatParams(model, config.optimizer)
        mk = lambda v, dt: torch.tensor(list(v), dtype=dt, device=dev)  # noqa: E731
        self.a_idx, self.a_scale = mk(actor_columns[0], torch.long), mk(actor_columns[1], torch.float32)
        self.c_idx, self.c_scale = mk(critic_columns[0], torch.long), mk(critic_columns[1], torch.float32)
        nj, depth, hd = model.actor.spec.num_joints, model.actor.depth, model.actor.rnns.hidden
        if nj != engine.NA:
            raise ValueError(f"the actor emits {nj} joints, the engine takes {engine.NA} actions")
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)  # noqa: E731
        self.rec = z(t + 1, n, engine.R)
        self.actions = z(t, n, nj)
        self.carry = {"actor": z(depth, n, hd), "critic": z(depth, n, hd)}       # env_states.model_carry
        self.carry0 = {"actor": z(depth, n, hd), "critic": z(depth, n, hd)}      # ... as it was when the rollout started
        self.alive = z(n)
        self.sample_key = torch.tensor([seed & 0xFFFFFFFF, 0x6B73696D], dtype=torch.int64, device=dev).to(torch.int32)
        self.sample_counter = z(1, dt=torch.int32)
        b = config.batch_size
        self.ix = z(b, dt=torch.long)
        self.data: dict[str, torch.Tensor] = {
            "obs_a": z(t, n, len(actor_columns[0])), "obs_c": z(t, n, len(critic_columns[0])), "done": z(t, n),
            "logp_on": z(t, n, nj), "values_on": z(t, n), "total": z(n, t), "dones": z(n, t, dt=torch.uint8), "succ": z(n, t, dt=torch.uint8)}
        self.loss = z()
        self.loss_terms = z(4)
        self.graphs: dict[str, torch.cuda.CUDAGraph] = {}
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.kernel_launches = 0   # launches of this repo's kernels in the last iteration (counted while running eagerly / capturing)

    # ------------------------------------------------------------------------------------------------- rollout
    def _rollout_steps(self) -> None:
        """``_single_unroll``'s scan (rl.py:1620): per control step ``sample_action`` (walking.py:730-763: the actor only, the
        critic carry passes through) then ``step_engine``; the model carry restarts where the episode ended (rl.py:1100-1123)."""
        eng, actor, done_col = self.eng, self.model.actor, self.eng.program["TI_R_DONE"]
        a_carry = self.carry["actor"]
        for s in range(self.t):
            row = self.rec[s]
            obs = (row.index_select(1, self.a_idx) * self.a_scale).unsqueeze(0)
            (y, h), = gru_stacks_forward([actor.rnns], [actor.input_proj(obs)], [a_carry], None)
            dist = MixtureOfGaussians(actor.output_proj(y), actor.spec)
            self.actions[s].copy_(dist.sample(self.sample_key, self.sample_counter)[0])
            self.sample_counter += 1
            eng.step(self.actions[s], row, self.rec[s + 1])
            keep = (row[:, done_col] == 0).to(torch.float32)
            a_carry.copy_(h * keep[None, :, None])
            self.alive *= keep

    def rollout(self) -> torch.Tensor:
        eng = self.eng
        for k in ("actor", "critic"):
            self.carry0[k].copy_(self.carry[k])
        self.rec[0].copy_(eng.first_rec)
        self.alive.fill_(1.0)
        with torch.no_grad():
            if self.use_graphs:
                if "rollout" not in self.graphs:
                    self._capture_rollout()
                self.graphs["rollout"].replay()
            else:
                self._rollout_steps()
        self.carry["critic"] *= self.alive[None, :, None]
        eng.first_rec.copy_(self.rec[self.t])
        return self.rec

    def _capture_rollout(self) -> None:
        # the warm-up steps are real steps of real environments: state, keys and carries are put back afterwards
        eng = self.eng
        live = (eng.state, eng.keys, eng.act_keys, self.carry["actor"], self.alive, self.sample_counter, self.rec)
        saved = [x.clone() for x in live]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            self._rollout_steps()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for dst, src in zip(live, saved):
            dst.copy_(src)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            self._rollout_steps()
        self.graphs["rollout"] = g

    # ------------------------------------------------------------------------------------------------- scoring
    def score(self) -> None:
        """``get_rewards`` (rl.py:189-245) and the done / success flags of the trajectory."""
        total, _ = self.eng.rewards(self.rec, self.t)
        dones, succ = self.eng.flags(self.rec, self.t)
        d = self.data
        d["total"].copy_(total); d["dones"].copy_(dones); d["succ"].copy_(succ)

    # -------------------------------------------------------------------------------------------------- update
    def _replay_all(self, carry: dict[str, torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor, dict[str, torch.Tensor]]:
        d = self.data
        with torch.no_grad():
            return get_ppo_variables(self.model, d["obs_a"], d["obs_c"], self.actions, d["done"], carry)

    def _single_step(self) -> None:
        """``PPOTask._single_step`` (ppo.py:495-565) on the minibatch ``self.ix``."""
        d, ix, c, eng = self.data, self.ix, self.config, self.eng
        carry = {k: v.index_select(1, ix) for k, v in self.carry0.items()}
        logp, values, _ = get_ppo_variables(self.model, d["obs_a"].index_select(1, ix), d["obs_c"].index_select(1, ix),
                                            self.actions.index_select(1, ix), d["done"].index_select(1, ix), carry)
        adv, tgt, _ = eng.gae(values.detach().t().contiguous(), d["total"].index_select(0, ix), d["dones"].index_select(0, ix),
                              d["succ"].index_select(0, ix), c.gamma, c.lam, c.normalize_advantages, c.monte_carlo_returns)
        # the loss is a mean over (b, t) of per-element terms: it is evaluated in the replay's time-major order
        loss, terms = ppo.ppo_loss(ppo.PPOInputs(adv.t().contiguous(), tgt.t().contiguous()),
                                   ppo.PPOVariables(d["logp_on"].index_select(1, ix), d["values_on"].index_select(1, ix)),
                                   ppo.PPOVariables(logp, values), c.loss)
        self.flat.zero_grad()
        loss.backward()
        self.flat.step(self.flat.allreduce_grads())
        self.loss.copy_(loss.detach())
        self.loss_terms.copy_(torch.stack([terms[k] for k in ppo.TERMS]))

    def _capture_step(self) -> None:
        f = self.flat
        live = (f.flat, f.mu, f.nu, f.step_count, f.bf16)
        saved = [x.clone() for x in live]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):                       # warm-up: cuBLAS workspaces, NCCL communicator, autograd buffers
                self._single_step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for dst, src in zip(live, saved):
            dst.copy_(src)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._single_step()
        self.graphs["step"] = g

    def update_model(self) -> None:
        """``PPOTask.update_model`` (ppo.py:572-694)."""
        d, c, t = self.data, self.config, self.t
        pre = self.rec[:t]
        torch.mul(pre.index_select(2, self.a_idx), self.a_scale, out=d["obs_a"])
        torch.mul(pre.index_select(2, self.c_idx), self.c_scale, out=d["obs_c"])
        d["done"].copy_(d["dones"].t())
        logp_on, values_on, _ = self._replay_all(self.carry0)
        d["logp_on"].copy_(logp_on); d["values_on"].copy_(values_on)
        for _ in range(c.num_passes):
            perm = torch.randperm(self.n, device=self.device, generator=self.gen)
            for b0 in range(0, self.n, c.batch_size):
                self.ix.copy_(perm[b0 : b0 + c.batch_size])
                if self.use_graphs:
                    if "step" not in self.graphs:
                        self._capture_step()
                    self.graphs["step"].replay()
                else:
                    self._single_step()
        _, _, carry = self._replay_all(self.carry0)
        for k in ("actor", "critic"):
            self.carry[k].copy_(carry[k])

    # ----------------------------------------------------------------------------------------------- iteration
    def iteration(self) -> dict[str, float]:
        """One ``_rl_train_loop_step``; returns device-timed milliseconds per phase (synchronises at the end)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stream = torch.cuda.current_stream(self.device)
        ev[0].record(stream)
        self.rollout()
        ev[1].record(stream)
        self.score()
        ev[2].record(stream)
        self.update_model()
        ev[3].record(stream)
        torch.cuda.synchronize(self.device)
        return {"rollout_ms": ev[0].elapsed_time(ev[1]), "scoring_ms": ev[1].elapsed_time(ev[2]), "update_ms": ev[2].elapsed_time(ev[3]),
                "iteration_ms": ev[0].elapsed_time(ev[3])}
