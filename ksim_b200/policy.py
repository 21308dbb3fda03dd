"""The walking / K-Bot policy networks and their action distribution on this repo's kernels (``SURVEY.md 8f.1``).

Mirrors ``examples/walking.py:40-253`` -- ``Actor`` / ``Critic`` / ``Model``: ``input_proj`` (Linear) -> ``depth`` x
``eqx.nn.GRUCell(hidden)`` -> ``eqx.nn.MLP`` (``num_hidden_layers`` hidden layers, gelu / elu, no final bias) -- and the
``xax.MixtureOfGaussians`` the actor returns (means + ``ZEROS`` bias clipped to the joint ranges, ``std = min((softplus +
min_std) var_scale, max_std)``, logits; walking.py:129-146).  ``get_ppo_variables`` is the batched form of
``WalkingTask.get_ppo_variables`` / ``_scan_fn`` (walking.py:674-728): only the GRU cells are sequential, so the input
projections, the output MLPs and the mixture log-probabilities run once over all ``(T, B)`` rows.

What runs where: the sequential GRU part (``b200sim_gru_forward`` / ``_backward``, tcgen05), the mixture log-prob / its
backward / sampling (``b200sim_mixture_*``) are this repo's kernels; the batched affine maps are plain library GEMMs (bf16
operands, fp32 accumulation) issued through torch.  Parameters are fp32; GEMMs read a bf16 shadow copy when the parameter
carries one (``ksim_b200.optim.FlatParams`` keeps it up to date inside the optimiser step).  All sequence tensors are
time-major ``[T, B, ...]``.  There is no CPU path for the kernels; the ``*_reference`` functions are the fp32 restatements the
tests compare against.
"""

from __future__ import annotations

__all__ = ["MixtureSpec", "MixtureOfGaussians", "mixture_params_reference", "mixture_log_prob_reference", "Linear", "MLP",
           "Actor", "Critic", "Model", "get_ppo_variables", "get_ppo_variables_reference", "shadow"]

import ctypes
import math
from dataclasses import dataclass
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import binding
from .replay import HIDDEN, GRUStack, gru_stack_reference, gru_stacks_forward, shadow


def _p(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(None if t is None else t.data_ptr())


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = binding.lib().b200sim_policy_last_error()
        raise binding.B200SimError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


# ------------------------------------------------------------------------------------------------ action distribution
@dataclass(frozen=True)
class MixtureSpec:
    """Static part of the actor head (walking.py:48-55, 129-143)."""

    num_joints: int
    num_mixtures: int
    min_std: float = 0.01
    max_std: float = 1.0
    var_scale: float = 0.5
    mean_bias: tuple[float, ...] | None = None   # ZEROS / JOINT_BIASES, per joint
    clip_lo: tuple[float, ...] | None = None      # ksim.ClipPositions: the joint ranges
    clip_hi: tuple[float, ...] | None = None


def mixture_params_reference(pred: torch.Tensor, spec: MixtureSpec) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """walking.py:129-143 in plain torch: ``(mean, std, log-weights)``, each ``[..., J, M]``."""
    jm = spec.num_joints * spec.num_mixtures
    shape = pred.shape[:-1] + (spec.num_joints, spec.num_mixtures)
    mean, std, logits = pred[..., :jm].reshape(shape), pred[..., jm : 2 * jm].reshape(shape), pred[..., 2 * jm :].reshape(shape)
    std = ((F.softplus(std) + spec.min_std) * spec.var_scale).clamp(max=spec.max_std)
    if spec.mean_bias is not None:
        mean = mean + torch.tensor(spec.mean_bias, dtype=pred.dtype, device=pred.device)[:, None]
    if spec.clip_lo is not None:
        lo = torch.tensor(spec.clip_lo, dtype=pred.dtype, device=pred.device)[:, None]
        hi = torch.tensor(spec.clip_hi, dtype=pred.dtype, device=pred.device)[:, None]
        mean = torch.minimum(torch.maximum(mean, lo), hi)
    return mean, std, torch.log_softmax(logits, dim=-1)


def mixture_log_prob_reference(pred: torch.Tensor, action: torch.Tensor, spec: MixtureSpec) -> torch.Tensor:
    """``xax.MixtureOfGaussians.log_prob``: ``logsumexp_k(log w_k + log N(x; mean_k, std_k))``, per joint."""
    mean, std, logw = mixture_params_reference(pred, spec)
    comp = -0.5 * ((action[..., None] - mean) / std) ** 2 - torch.log(std) - 0.5 * math.log(2 * math.pi)
    return torch.logsumexp(logw + comp, dim=-1)


class _SpecBuffers:
    """Per-device copies of a spec's per-joint arrays."""

    _cache: dict[tuple, tuple[torch.Tensor | None, torch.Tensor | None, torch.Tensor | None]] = {}

    @classmethod
    def get(cls, spec: MixtureSpec, device: torch.device) -> tuple[torch.Tensor | None, torch.Tensor | None, torch.Tensor | None]:
        key = (spec, str(device))
        if key not in cls._cache:
            mk = lambda v: None if v is None else torch.tensor(v, dtype=torch.float32, device=device)  # noqa: E731
            cls._cache[key] = (mk(spec.mean_bias), mk(spec.clip_lo), mk(spec.clip_hi))
        return cls._cache[key]


def _rows(pred: torch.Tensor, spec: MixtureSpec) -> int:
    if pred.shape[-1] != 3 * spec.num_joints * spec.num_mixtures:
        raise ValueError(f"prediction width {pred.shape[-1]} != 3 x {spec.num_joints} x {spec.num_mixtures}")
    if not pred.is_cuda:
        raise binding.B200SimError("the mixture kernels need CUDA tensors; there is no CPU path")
    return pred.numel() // pred.shape[-1]


class _MixtureLogProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: torch.Tensor, action: torch.Tensor, spec: MixtureSpec) -> torch.Tensor:  # noqa: ANN001
        rows = _rows(pred, spec)
        pred = pred.detach().to(torch.float32).contiguous()
        action = action.detach().to(torch.float32).contiguous()
        bias, lo, hi = _SpecBuffers.get(spec, pred.device)
        logp = torch.empty(pred.shape[:-1] + (spec.num_joints,), dtype=torch.float32, device=pred.device)
        cf = ctypes.c_float
        with torch.cuda.device(pred.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
            _check(binding.lib().b200sim_mixture_logprob(stream, rows, spec.num_joints, spec.num_mixtures, _p(pred), _p(action), _p(bias),
                                                         _p(lo), _p(hi), cf(spec.min_std), cf(spec.var_scale), cf(spec.max_std), _p(logp)),
                   "b200sim_mixture_logprob")
        ctx.save_for_backward(pred, action)
        ctx.spec = spec
        return logp

    @staticmethod
    def backward(ctx, g: torch.Tensor):  # noqa: ANN001, ANN205
        pred, action = ctx.saved_tensors
        spec: MixtureSpec = ctx.spec
        rows = _rows(pred, spec)
        bias, lo, hi = _SpecBuffers.get(spec, pred.device)
        # the PPO loss hands back one gradient per (t, b) row broadcast over the joints (a stride-0 view): pass it as such
        per_row = g.stride(-1) == 0 and g.shape[-1] > 1
        gg = (g[..., 0] if per_row else g).to(torch.float32).contiguous()
        d_pred = torch.empty_like(pred)
        cf = ctypes.c_float
        with torch.cuda.device(pred.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
            _check(binding.lib().b200sim_mixture_logprob_backward(stream, rows, spec.num_joints, spec.num_mixtures, _p(pred), _p(action),
                                                                  _p(bias), _p(lo), _p(hi), cf(spec.min_std), cf(spec.var_scale),
                                                                  cf(spec.max_std), _p(gg), int(per_row), _p(d_pred), 0),
                   "b200sim_mixture_logprob_backward")
        return d_pred, None, None


class MixtureOfGaussians:
    """The distribution ``Actor.forward`` returns (``xax.MixtureOfGaussians(mean_nm, std_nm, logits_nm)``), kept as the raw
    prediction ``pred [..., 3 J M]`` plus the static head; ``log_prob`` is differentiable in ``pred``."""

    def __init__(self, pred: torch.Tensor, spec: MixtureSpec) -> None:
        self.pred, self.spec = pred, spec

    def log_prob(self, action: torch.Tensor) -> torch.Tensor:
        return _MixtureLogProbFn.apply(self.pred, action, self.spec)

    def _draw(self, key: torch.Tensor, counter: torch.Tensor | None, argmax: bool, want_logp: bool) -> tuple[torch.Tensor, torch.Tensor | None]:
        spec, pred = self.spec, self.pred.detach().to(torch.float32).contiguous()
        rows = _rows(pred, spec)
        bias, lo, hi = _SpecBuffers.get(spec, pred.device)
        action = torch.empty(pred.shape[:-1] + (spec.num_joints,), dtype=torch.float32, device=pred.device)
        logp = torch.empty_like(action) if want_logp else None
        cf = ctypes.c_float
        with torch.cuda.device(pred.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
            _check(binding.lib().b200sim_mixture_sample(stream, rows, spec.num_joints, spec.num_mixtures, _p(pred), _p(bias), _p(lo), _p(hi),
                                                        cf(spec.min_std), cf(spec.var_scale), cf(spec.max_std), _p(key), _p(counter),
                                                        int(argmax), _p(action), _p(logp)), "b200sim_mixture_sample")
        return action, logp

    def sample(self, key: torch.Tensor, counter: torch.Tensor | None = None, with_log_prob: bool = False,
               ) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor]:
        """``key``: int32/uint32 ``[2]`` device tensor; ``counter``: int32 ``[1]`` device tensor (the call number: a captured
        CUDA graph draws fresh numbers by advancing it on the device)."""
        a, lp = self._draw(key, counter, False, with_log_prob)
        return (a, lp) if with_log_prob else a

    def mode(self) -> torch.Tensor:
        dummy = torch.zeros(2, dtype=torch.int32, device=self.pred.device)
        return self._draw(dummy, None, True, False)[0]


# ---------------------------------------------------------------------------------------------------------- networks
_ONES: dict[tuple, torch.Tensor] = {}


def _ones_row(n: int, device: torch.device) -> torch.Tensor:
    """``[1, n]`` bf16 ones (cached): column sums as a tensor-core GEMM, ``ones @ dy``, instead of a strided reduction."""
    key = (n, str(device))
    if key not in _ONES:
        _ONES[key] = torch.ones((1, n), dtype=torch.bfloat16, device=device)
    return _ONES[key]


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class _AffineFn(torch.autograd.Function):
    """``y = x W^T + b`` with bf16 tensor-core operands and fp32 accumulation; ``W`` read from its bf16 shadow.  Gradients:
    ``dW = dy^T x`` and ``db`` in fp32 (they land in the fp32 flat gradient buffer), ``dx`` in bf16.  Operand widths that are
    not a multiple of 8 elements (the 43 actor inputs, the 510 mixture outputs) are zero-padded so that the library picks its
    16-byte-aligned tensor-core kernels; a single output column (the critic's value head) is a matrix-vector product."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, out_fp32: bool) -> torch.Tensor:  # noqa: ANN001
        k, n_out = int(x.shape[-1]), int(weight.shape[0])
        x16 = x.detach().to(torch.bfloat16).reshape(-1, k)
        w16 = shadow(weight)
        kp, np_ = _pad8(k), (n_out if n_out == 1 else _pad8(n_out))
        if kp != k:
            x16 = F.pad(x16, (0, kp - k))
        if kp != k or np_ != n_out:
            w16 = F.pad(w16, (0, kp - k, 0, np_ - n_out))
        if n_out == 1:
            y = torch.mv(x16, w16[0]).unsqueeze(-1)
            y = y.float() if out_fp32 else y
            if bias is not None:
                y = y + (bias.detach() if out_fp32 else shadow(bias))
        elif out_fp32:
            y = torch.mm(x16, w16.t(), out_dtype=torch.float32)
            if bias is not None:
                y[:, :n_out] += bias.detach()
        elif bias is not None:
            b16 = shadow(bias)
            y = torch.addmm(b16 if np_ == n_out else F.pad(b16, (0, np_ - n_out)), x16, w16.t())
        else:
            y = torch.mm(x16, w16.t())
        if np_ != n_out:
            y = y[:, :n_out].contiguous()
        ctx.save_for_backward(x16, w16)
        ctx.has_bias, ctx.x_shape, ctx.x_dtype, ctx.dims = bias is not None, x.shape, x.dtype, (k, kp, n_out, np_)
        return y.view(*x.shape[:-1], n_out)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):  # noqa: ANN001, ANN205
        x16, w16 = ctx.saved_tensors
        k, kp, n_out, np_ = ctx.dims
        dy16 = dy.to(torch.bfloat16).reshape(-1, n_out)
        rows = dy16.shape[0]
        db = None
        if n_out == 1:
            dx = (dy16 * w16) if ctx.needs_input_grad[0] else None        # outer product [rows, 1] x [1, kp]
            dw = torch.mv(x16.t(), dy16[:, 0]).float().unsqueeze(0)
            if ctx.has_bias:
                db = dy16.sum(0, dtype=torch.float32)
        else:
            if np_ != n_out:
                dy16 = F.pad(dy16, (0, np_ - n_out))
            dx = torch.mm(dy16, w16) if ctx.needs_input_grad[0] else None
            dw = torch.mm(dy16.t(), x16, out_dtype=torch.float32)[:n_out]
            if ctx.has_bias:
                db = torch.mm(_ones_row(rows, dy16.device), dy16, out_dtype=torch.float32)[0, :n_out]
        if dx is not None:
            dx = (dx[:, :k] if kp != k else dx).reshape(ctx.x_shape).to(ctx.x_dtype)
        return dx, (dw[:, :k] if kp != k else dw), db, None


class Linear(nn.Module):
    """``eqx.nn.Linear`` (weight ``[out, in]``, uniform ``+-1/sqrt(in)`` init)."""

    def __init__(self, in_features: int, out_features: int, use_bias: bool = True) -> None:
        super().__init__()
        lim = 1.0 / math.sqrt(in_features)
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-lim, lim))
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-lim, lim)) if use_bias else None

    def forward(self, x: torch.Tensor, out_fp32: bool = False) -> torch.Tensor:
        return _AffineFn.apply(x, self.weight, self.bias, out_fp32)

    def reference(self, x: torch.Tensor) -> torch.Tensor:
        return F.linear(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype))


class MLP(nn.Module):
    """``eqx.nn.MLP(in, out, width, depth, activation, use_final_bias=False)``: ``depth`` hidden layers, ``depth + 1`` affine
    maps, the activation after each hidden layer, identity at the end.  ``"gelu"`` is ``jax.nn.gelu``'s default (tanh form)."""

    def __init__(self, in_size: int, out_size: int, width_size: int, depth: int, activation: str, use_final_bias: bool = False) -> None:
        super().__init__()
        sizes = [in_size] + [width_size] * depth + [out_size]
        self.layers = nn.ModuleList([Linear(sizes[i], sizes[i + 1], use_bias=(i < depth) or use_final_bias) for i in range(depth + 1)])
        if activation not in ("gelu", "elu"):
            raise ValueError(f"activation {activation!r}")
        self.activation = activation

    def _act(self, x: torch.Tensor) -> torch.Tensor:
        return F.gelu(x, approximate="tanh") if self.activation == "gelu" else F.elu(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers[:-1]:
            x = self._act(layer(x))
        return self.layers[-1](x, out_fp32=True)

    def reference(self, x: torch.Tensor) -> torch.Tensor:
        for layer in self.layers[:-1]:
            x = self._act(layer.reference(x))
        return self.layers[-1].reference(x)


class _Recurrent(nn.Module):
    def __init__(self, num_inputs: int, num_outputs: int, hidden_size: int, depth: int, num_hidden_layers: int, activation: str) -> None:
        super().__init__()
        if hidden_size != HIDDEN:
            raise ValueError(f"the tensor-core GRU kernels are built for hidden_size = {HIDDEN}")
        self.num_inputs, self.num_outputs, self.depth = num_inputs, num_outputs, depth
        self.input_proj = Linear(num_inputs, hidden_size)
        self.rnns = GRUStack(depth, hidden_size)
        self.output_proj = MLP(hidden_size, num_outputs, hidden_size, num_hidden_layers, activation)

    def reference(self, obs: torch.Tensor, carry: torch.Tensor | None, keep: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor]:
        """fp32 (or fp64) step-by-step restatement over ``obs [T, B, I]``: returns the head output and the final carry."""
        y, carry_out = gru_stack_reference(self.rnns, self.input_proj.reference(obs), carry, keep)
        return self.output_proj.reference(y), carry_out


class Actor(_Recurrent):
    """``examples/walking.py:40-146``; ``forward`` is the batched-over-time form of ``Actor.forward`` (``keep [T, B]`` carries
    the ``jnp.where(transition.done, initial_carry, next_carry)`` of ``_scan_fn``, walking.py:715-719)."""

    def __init__(self, *, num_inputs: int, num_joints: int, hidden_size: int = HIDDEN, depth: int = 2, num_hidden_layers: int = 2,
                 num_mixtures: int = 10, min_std: float = 0.01, max_std: float = 1.0, var_scale: float = 0.5,
                 mean_bias: Sequence[float] | None = None, clip_lo: Sequence[float] | None = None, clip_hi: Sequence[float] | None = None) -> None:
        super().__init__(num_inputs, num_joints * 3 * num_mixtures, hidden_size, depth, num_hidden_layers, "gelu")
        tup = lambda v: None if v is None else tuple(float(x) for x in v)  # noqa: E731
        self.spec = MixtureSpec(num_joints, num_mixtures, min_std, max_std, var_scale, tup(mean_bias), tup(clip_lo), tup(clip_hi))


class Critic(_Recurrent):
    """``examples/walking.py:149-205``."""

    def __init__(self, *, num_inputs: int, hidden_size: int = HIDDEN, depth: int = 2, num_hidden_layers: int = 2) -> None:
        super().__init__(num_inputs, 1, hidden_size, depth, num_hidden_layers, "elu")


class Model(nn.Module):
    """``examples/walking.py:208-253``."""

    def __init__(self, *, num_actor_inputs: int, num_critic_inputs: int, num_joints: int, hidden_size: int = HIDDEN, depth: int = 2,
                 num_hidden_layers: int = 2, num_mixtures: int = 10, min_std: float = 0.01, max_std: float = 1.0, var_scale: float = 0.5,
                 mean_bias: Sequence[float] | None = None, clip_lo: Sequence[float] | None = None, clip_hi: Sequence[float] | None = None) -> None:
        super().__init__()
        self.actor = Actor(num_inputs=num_actor_inputs, num_joints=num_joints, hidden_size=hidden_size, depth=depth,
                           num_hidden_layers=num_hidden_layers, num_mixtures=num_mixtures, min_std=min_std, max_std=max_std,
                           var_scale=var_scale, mean_bias=mean_bias, clip_lo=clip_lo, clip_hi=clip_hi)
        self.critic = Critic(num_inputs=num_critic_inputs, hidden_size=hidden_size, depth=depth, num_hidden_layers=num_hidden_layers)

    def forward(self, actor_obs: torch.Tensor, critic_obs: torch.Tensor, carry: dict[str, torch.Tensor | None], keep: torch.Tensor | None,
                ) -> tuple[MixtureOfGaussians, torch.Tensor, dict[str, torch.Tensor]]:
        """Both networks over ``[T, B, ...]`` observations from ``carry = {"actor": [depth, B, H], "critic": ...}``; the two GRU
        stacks run side by side in the same persistent kernel launches.  Returns the action distribution for every ``(t, b)``,
        the values ``[T, B]`` and the carry after the last step."""
        a, c = self.actor, self.critic
        xa, xc = a.input_proj(actor_obs), c.input_proj(critic_obs)
        (ya, ha), (yc, hc) = gru_stacks_forward([a.rnns, c.rnns], [xa, xc], [carry.get("actor"), carry.get("critic")], keep)
        dist = MixtureOfGaussians(a.output_proj(ya), a.spec)
        values = c.output_proj(yc)[..., 0]
        return dist, values, {"actor": ha, "critic": hc}


def get_ppo_variables(model: Model, actor_obs: torch.Tensor, critic_obs: torch.Tensor, action: torch.Tensor, done: torch.Tensor,
                      model_carry: dict[str, torch.Tensor | None]) -> tuple[torch.Tensor, torch.Tensor, dict[str, torch.Tensor]]:
    """``WalkingTask.get_ppo_variables`` (walking.py:674-728) for a time-major minibatch: ``(log_probs [T, B, J], values [T, B],
    next_model_carry)``.  ``done [T, B]`` (any dtype): the carry restarts from the initial (zero) carry after a terminal step."""
    keep = (done == 0).to(torch.float32)
    dist, values, carry = model(actor_obs, critic_obs, model_carry, keep)
    return dist.log_prob(action), values, carry


def get_ppo_variables_reference(model: Model, actor_obs: torch.Tensor, critic_obs: torch.Tensor, action: torch.Tensor, done: torch.Tensor,
                                model_carry: dict[str, torch.Tensor | None]) -> tuple[torch.Tensor, torch.Tensor, dict[str, torch.Tensor]]:
    """The same, step by step in the dtype of the observations (fp32 / fp64), on any device: the test restatement."""
    keep = (done == 0).to(actor_obs.dtype)
    pred, ha = model.actor.reference(actor_obs, model_carry.get("actor"), keep)
    val, hc = model.critic.reference(critic_obs, model_carry.get("critic"), keep)
    return mixture_log_prob_reference(pred, action.to(pred.dtype), model.actor.spec), val[..., 0], {"actor": ha, "critic": hc}
