"""Flat parameter / gradient buffers and the optimiser step of the PPO update (``SURVEY.md 8e / 8f.1``).

The reference applies an optax chain to the gradient pytree once per minibatch (``ksim/task/ppo.py:536-565`` with the chain
of ``examples/walking.py:373-387``); with env-sharded ranks the gradient is averaged over ranks first (one collective,
``SURVEY.md 8e``).  Here every parameter of the model is a view into ONE flat fp32 buffer, every ``.grad`` a view into a second
one, so that

* the cross-rank mean is a single in-place NCCL all-reduce of the flat gradient buffer (no pack / unpack copies), issued on
  the update's stream and therefore capturable inside the update's CUDA graph; the ``1 / world_size`` is folded into the
  optimiser kernel;
* the whole optax chain is one call of ``b200sim_optimizer_step`` (two HBM-bound passes), which also refreshes the bf16
  shadow copy of the parameters the tensor-core GEMMs read.

``optimizer_step_reference`` is the numpy (float64) restatement of the optax chain the tests compare against.
"""

from __future__ import annotations

__all__ = ["OptimizerConfig", "FlatParams", "optimizer_step_reference"]

import ctypes
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import binding


@dataclass(frozen=True)
class OptimizerConfig:
    """``learning_rate`` / ``warmup_steps`` / ``adam_weight_decay`` / ``grad_clip`` of ``WalkingConfig`` (walking.py:322-343)
    and optax's ``scale_by_adam`` defaults."""

    learning_rate: float = 3e-4
    warmup_steps: int = 100
    adam_weight_decay: float = 1e-5
    grad_clip: float = 2.0
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    init_scale: float = 0.01   # warmup_constant_schedule(init_value = learning_rate * 0.01)


def _align(n: int, a: int = 4) -> int:
    return (n + a - 1) // a * a


class FlatParams:
    """Moves a module's parameters into one flat fp32 buffer (each parameter becomes a view, 16-byte aligned), gives every
    parameter a persistent ``.grad`` view into a flat gradient buffer and a ``.shadow`` bf16 view, and owns the Adam moments."""

    def __init__(self, module: nn.Module, config: OptimizerConfig = OptimizerConfig()) -> None:
        self.config = config
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.device = dev
        offs, total = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatParams needs fp32 parameters on one device")
            offs.append(total)
            total += _align(p.numel())
        self.numel = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.mu = torch.zeros(total, dtype=torch.float32, device=dev)
        self.nu = torch.zeros(total, dtype=torch.float32, device=dev)
        self.bf16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.scratch = None
        for p, off in zip(self.params, offs):
            n = p.numel()
            self.flat[off : off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off : off + n].view(p.shape)
            p.grad = self.grad[off : off + n].view(p.shape)
            p.shadow = self.bf16[off : off + n].view(p.shape)
        self.refresh_shadow()

    def refresh_shadow(self) -> None:
        self.bf16.copy_(self.flat)

    def zero_grad(self) -> None:
        """Keeps the ``.grad`` views (``set_to_none`` would detach them from the flat buffer)."""
        self.grad.zero_()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr():
                raise RuntimeError("a parameter's .grad no longer lives in the flat gradient buffer (zero_grad(set_to_none=True)?)")

    def allreduce_grads(self) -> float:
        """Sum of the flat gradient over ranks, in place, one collective on the current stream.  Returns the factor the
        optimiser step folds in to make it the mean (``1 / world_size``).  Every rank reduces the same ``numel`` whatever
        subset of its parameters received a gradient (missing ones are zeros in the flat buffer)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 1.0
        dist.all_reduce(self.grad)
        return 1.0 / dist.get_world_size()

    def step(self, grad_scale: float = 1.0) -> None:
        """``optimizer.update`` + ``apply_updates`` (ppo.py:536-565): one ``b200sim_optimizer_step``."""
        lib, c = binding.lib(), self.config
        if self.device.type != "cuda":
            raise binding.B200SimError("b200sim_optimizer_step needs CUDA buffers; there is no CPU path")
        if self.scratch is None:
            self.scratch = torch.zeros(int(lib.b200sim_optimizer_scratch_bytes()), dtype=torch.uint8, device=self.device)
        cf, vp = ctypes.c_float, ctypes.c_void_p
        with torch.cuda.device(self.device):
            stream = vp(torch.cuda.current_stream(self.device).cuda_stream)
            rc = lib.b200sim_optimizer_step(stream, self.numel, vp(self.flat.data_ptr()), vp(self.grad.data_ptr()), vp(self.mu.data_ptr()),
                                            vp(self.nu.data_ptr()), vp(self.bf16.data_ptr()), vp(self.step_count.data_ptr()), cf(grad_scale),
                                            cf(c.learning_rate * c.init_scale), cf(c.learning_rate), int(c.warmup_steps), cf(c.b1), cf(c.b2),
                                            cf(c.eps), cf(c.adam_weight_decay), cf(c.grad_clip), vp(self.scratch.data_ptr()),
                                            vp(self.stats.data_ptr()))
        if rc != 0:
            msg = lib.b200sim_optimizer_last_error()
            raise binding.B200SimError(f"b200sim_optimizer_step failed with code {rc}: {msg.decode() if msg else ''}")

    @property
    def grad_norm(self) -> torch.Tensor:
        return self.stats[0]


def optimizer_step_reference(p: np.ndarray, g: np.ndarray, mu: np.ndarray, nu: np.ndarray, count: int, c: OptimizerConfig,
                             grad_scale: float = 1.0) -> tuple[np.ndarray, np.ndarray, np.ndarray, float, float]:
    """The optax chain of walking.py:380-387 in float64, transformation by transformation.  ``count`` = steps taken before this
    one.  Returns ``(p, mu, nu, grad_norm, lr)``."""
    g = np.asarray(g, dtype=np.float64) * grad_scale
    p, mu, nu = (np.asarray(x, dtype=np.float64) for x in (p, mu, nu))
    g = np.where(np.isnan(g), 0.0, g)                                        # zero_nans
    norm = float(np.sqrt(np.sum(g * g)))
    if not norm < c.grad_clip:                                               # clip_by_global_norm
        g = g / norm * c.grad_clip
    g = g + c.adam_weight_decay * p                                          # add_decayed_weights
    mu = c.b1 * mu + (1 - c.b1) * g                                          # scale_by_adam
    nu = c.b2 * nu + (1 - c.b2) * g * g
    t = count + 1
    u = (mu / (1 - c.b1**t)) / (np.sqrt(nu / (1 - c.b2**t)) + c.eps)
    init = c.learning_rate * c.init_scale                                    # warmup_constant_schedule at `count`
    frac = min(count / c.warmup_steps, 1.0) if c.warmup_steps > 0 else 1.0
    lr = init + (c.learning_rate - init) * frac
    return p - lr * u, mu, nu, norm, lr                                      # scale_by_schedule, scale(-1), apply_updates
