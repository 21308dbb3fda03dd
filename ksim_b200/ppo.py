"""Host mirror of the PPO loss of ``ksim/task/ppo.py`` (``PPOInputs`` :24-30, ``compute_ppo_loss`` :149-246 and the scalar
``_get_ppo_loss_and_metrics`` reduces it to, :471-486) on top of ``b200sim_ppo_loss``: one CUDA pass computes the per-(b, t)
loss terms, their mean, and the gradient of that mean with respect to the new policy's log-probs / values / entropy -- the
backward pass ``jax.grad`` derives in ``_single_step`` (:521-534).  This is the loss half of ``SURVEY.md 8f.1``; the policy
replay (the user's networks, ``get_ppo_variables``) stays in the host framework and meets the kernel through
``PPOLossFunction``, a ``torch.autograd.Function`` whose backward hands the kernel's gradients to autograd.

There is no fallback: without ``libb200sim.so`` or off a CUDA device every entry point raises.
"""

from __future__ import annotations

__all__ = ["PPOInputs", "PPOVariables", "PPOLossConfig", "compute_ppo_loss", "ppo_loss_and_grads", "PPOLossFunction", "ppo_loss"]

import ctypes
from dataclasses import dataclass

import torch

from . import binding
from . import codes as C

TERMS = ("policy", "value", "kl", "entropy")


@dataclass
class PPOInputs:
    """``ksim.task.ppo.PPOInputs`` (ppo.py:24-30); produced by ``B200Engine.gae`` (``compute_ppo_inputs``)."""

    advantages_bt: torch.Tensor
    value_targets_bt: torch.Tensor
    gae_bt: torch.Tensor | None = None
    returns_bt: torch.Tensor | None = None


@dataclass
class PPOVariables:
    """``ksim.task.ppo.PPOVariables`` (ppo.py:33-39; ``aux_losses`` stay in the host framework): ``log_probs [B, T, A]``, ``values [B, T]``, ``entropy [B, T, A]``."""

    log_probs: torch.Tensor
    values: torch.Tensor
    entropy: torch.Tensor | None = None


@dataclass(frozen=True)
class PPOLossConfig:
    """The loss fields of ``PPOConfig`` (ppo.py:251-306) with the reference's defaults."""

    clip_param: float = 0.2
    value_loss_coef: float = 0.5
    entropy_coef: float = 0.008
    kl_coef: float = 1e-3
    log_clip_value: float = 5.0
    use_clipped_value_loss: bool = True


def _ptr(t: torch.Tensor | None) -> ctypes.c_void_p:
    return ctypes.c_void_p(None if t is None else t.data_ptr())


def _chk(t: torch.Tensor, shape: tuple[int, ...], device: torch.device, name: str) -> torch.Tensor:
    if t.device != device or t.dtype != torch.float32 or tuple(t.shape) != shape:
        raise ValueError(f"{name}: expected float32 {shape} on {device}, got {t.dtype} {tuple(t.shape)} on {t.device}")
    return t.contiguous()


def ppo_loss_and_grads(ppo_inputs: PPOInputs, on_policy_variables: PPOVariables, off_policy_variables: PPOVariables,
                       config: PPOLossConfig = PPOLossConfig(), kl_scale: float = 1.0, want_losses: bool = True, want_grads: bool = True,
                       ) -> tuple[torch.Tensor | None, torch.Tensor, torch.Tensor | None, torch.Tensor | None]:
    """One launch pair of ``b200sim_ppo_loss``.  Returns ``(losses[4, B, T] | None, out[5], d_logp[B, T] | None,
    d_values[B, T] | None)``: ``out`` = mean over (B, T) of the summed terms, then of each term (policy, value, kl, entropy);
    ``d_logp[b, t]`` is the gradient of ``out[0]`` for every ``log_probs[b, t, a]`` of the new policy."""
    lp_on = on_policy_variables.log_probs
    if lp_on.dim() != 3:
        raise ValueError("log_probs must be [B, T, A] (ppo.py:196-198)")
    if not lp_on.is_cuda:
        raise binding.B200SimError("b200sim_ppo_loss needs CUDA tensors; there is no CPU path")
    b, t, a = (int(x) for x in lp_on.shape)
    dev = lp_on.device
    lp_on = _chk(lp_on, (b, t, a), dev, "on_policy log_probs")
    lp_off = _chk(off_policy_variables.log_probs, (b, t, a), dev, "off_policy log_probs")
    v_on = _chk(on_policy_variables.values, (b, t), dev, "on_policy values")
    v_off = _chk(off_policy_variables.values, (b, t), dev, "off_policy values")
    adv = _chk(ppo_inputs.advantages_bt, (b, t), dev, "advantages_bt")
    tgt = _chk(ppo_inputs.value_targets_bt, (b, t), dev, "value_targets_bt")
    ent = None if off_policy_variables.entropy is None else _chk(off_policy_variables.entropy, (b, t, a), dev, "entropy")
    losses = torch.empty((C.PPO_N_TERMS, b, t), dtype=torch.float32, device=dev) if want_losses else None  # every element is written
    out = torch.zeros(1 + C.PPO_N_TERMS, dtype=torch.float32, device=dev)
    d_logp = torch.empty((b, t), dtype=torch.float32, device=dev) if want_grads else None
    d_values = torch.empty((b, t), dtype=torch.float32, device=dev) if want_grads else None
    scratch = torch.empty(C.PPO_SCRATCH_BYTES, dtype=torch.uint8, device=dev)
    flags = C.PPO_CLIPPED_VALUE_LOSS if config.use_clipped_value_loss else 0
    cf = ctypes.c_float
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        binding.check(binding.lib().b200sim_ppo_loss(
            stream, b, t, a, _ptr(lp_on), _ptr(lp_off), _ptr(v_on), _ptr(v_off), _ptr(adv), _ptr(tgt), _ptr(ent), cf(config.clip_param),
            cf(config.value_loss_coef), cf(config.entropy_coef), cf(config.kl_coef), cf(config.log_clip_value), cf(kl_scale), flags,
            _ptr(losses), _ptr(out), _ptr(d_logp), _ptr(d_values), _ptr(scratch)), "b200sim_ppo_loss")
    return losses, out, d_logp, d_values


def compute_ppo_loss(ppo_inputs: PPOInputs, on_policy_variables: PPOVariables, off_policy_variables: PPOVariables, *,
                     clip_param: float, value_loss_coef: float, entropy_coef: float, kl_coef: float, log_clip_value: float,
                     use_clipped_value_loss: bool) -> dict[str, torch.Tensor]:
    """``ksim.task.ppo.compute_ppo_loss`` (ppo.py:149-246), same arguments: a dict of loss terms, each ``[B, T]``
    (``"entropy"`` only when the new policy reports one)."""
    cfg = PPOLossConfig(clip_param, value_loss_coef, entropy_coef, kl_coef, log_clip_value, use_clipped_value_loss)
    losses, _, _, _ = ppo_loss_and_grads(ppo_inputs, on_policy_variables, off_policy_variables, cfg, want_grads=False)
    assert losses is not None
    res = {"policy": losses[0], "value": losses[1], "kl": losses[2]}
    if off_policy_variables.entropy is not None:
        res["entropy"] = losses[3]
    return res


class PPOLossFunction(torch.autograd.Function):
    """``loss = mean_bt(sum of terms)`` (ppo.py:484-486) with the kernel's analytic backward.  Differentiable inputs: the new
    policy's ``log_probs``, ``values`` and (optional) ``entropy``; everything else is a constant, as in the reference
    (``stop_gradient`` on the PPO inputs, ppo.py:121; the on-policy variables are data)."""

    @staticmethod
    def forward(ctx, log_probs, values, entropy, on_log_probs, on_values, advantages, value_targets, config, kl_scale):  # noqa: ANN001, ANN205
        _, out, d_logp, d_values = ppo_loss_and_grads(
            PPOInputs(advantages, value_targets), PPOVariables(on_log_probs, on_values), PPOVariables(log_probs, values, entropy),
            config, kl_scale, want_losses=False)
        ctx.save_for_backward(d_logp, d_values)
        ctx.n_act = int(log_probs.shape[-1])
        ctx.has_entropy = entropy is not None
        ctx.d_entropy = -config.entropy_coef / float(values.numel())
        ctx.entropy_shape = None if entropy is None else tuple(entropy.shape)
        terms = out[1:]
        ctx.mark_non_differentiable(terms)
        return out[0], terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):  # noqa: ANN001, ANN205
        d_logp, d_values = ctx.saved_tensors
        g_lp = (g_loss * d_logp).unsqueeze(-1).expand(-1, -1, ctx.n_act)
        g_v = g_loss * d_values
        g_e = None
        if ctx.has_entropy:
            g_e = (g_loss * ctx.d_entropy).expand(ctx.entropy_shape)
        return g_lp, g_v, g_e, None, None, None, None, None, None


def ppo_loss(ppo_inputs: PPOInputs, on_policy_variables: PPOVariables, off_policy_variables: PPOVariables,
             config: PPOLossConfig = PPOLossConfig(), kl_scale: float = 1.0) -> tuple[torch.Tensor, dict[str, torch.Tensor]]:
    """The scalar the optimiser minimises (``_get_ppo_loss_and_metrics``, ppo.py:458-488) and the mean of each term (for
    logging), differentiable with respect to ``off_policy_variables`` through ``PPOLossFunction``."""
    loss, terms = PPOLossFunction.apply(off_policy_variables.log_probs, off_policy_variables.values, off_policy_variables.entropy,
                                        on_policy_variables.log_probs, on_policy_variables.values, ppo_inputs.advantages_bt,
                                        ppo_inputs.value_targets_bt, config, kl_scale)
    return loss, {name: terms[i] for i, name in enumerate(TERMS)}
