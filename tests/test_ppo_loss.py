"""PPO loss oracle (``compute_ppo_loss`` + the mean of the summed terms, ksim/task/ppo.py:124-135, 149-246, 471-486) and its
analytic gradient against an independent restatement: the reference's jnp expressions written line by line in torch
(float64) with the gradient taken by autograd instead of derived by hand.  The reference's tests hold no PPO vectors
(parity unpinned by the reference; pinned by this restatement)."""

import numpy as np
import pytest
import torch

from ksim_b200 import codes as C


def torch_ppo_loss(lp_on, lp_off, v_on, v_off, adv, tgt, ent, clip_param, value_loss_coef, entropy_coef, kl_coef, log_clip_value,  # noqa: ANN001, ANN201
                   kl_scale, clipped):
    """ppo.py:204-244 and :471-486, expression by expression."""
    log_ratio = torch.sum(lp_off - lp_on, dim=-1)
    ratio = torch.exp(torch.clip(log_ratio, -log_clip_value, log_clip_value))
    clipped_ratio = torch.clip(ratio, 1 - clip_param, 1 + clip_param)
    policy_loss = -torch.minimum(ratio * adv, clipped_ratio * adv)
    if clipped:
        value_clipped = v_on + (v_off - v_on).clip(-clip_param, clip_param)
        value_objective = 0.5 * torch.maximum((v_off - tgt) ** 2, (value_clipped - tgt) ** 2)
    else:
        value_objective = 0.5 * (tgt - v_off) ** 2
    value_loss = value_objective * value_loss_coef
    kl_loss = (lp_on - lp_off).sum(dim=-1) * kl_coef
    losses = {"policy": policy_loss, "value": value_loss, "kl": kl_loss * kl_scale}
    if ent is not None:
        losses["entropy"] = -ent.sum(dim=-1) * entropy_coef
    loss = torch.stack(list(losses.values()), dim=-1).sum(dim=-1).mean()
    return losses, loss


def make_case(b, t, a, seed, spread=1.0):  # noqa: ANN001, ANN201
    """Random PPO variables that exercise every branch: ratios inside and outside the clip range on both signs of the
    advantage, log-ratios beyond +-log_clip_value, value moves inside and outside +-clip_param."""
    r = np.random.RandomState(seed)
    lp_on = -1.0 + 0.5 * r.randn(b, t, a)
    lp_off = lp_on + spread * 0.15 * r.randn(b, t, a)
    if b * t > 4:
        lp_off.reshape(-1, a)[0] += 3.0   # log ratio far above the log clip
        lp_off.reshape(-1, a)[1] -= 3.0
    v_on = r.randn(b, t)
    v_off = v_on + 0.3 * r.randn(b, t)
    adv = r.randn(b, t)
    tgt = v_on + 0.5 * r.randn(b, t)
    ent = 1.0 + 0.2 * r.randn(b, t, a)
    return lp_on, lp_off, v_on, v_off, adv, tgt, ent


HYPER = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.008, kl_coef=1e-3, log_clip_value=5.0)


@pytest.fixture(scope="module")
def ora(oracle_built, walking_program):  # noqa: ANN001, ANN201
    from oracle import OracleEngine

    return OracleEngine(walking_program, "f64")


@pytest.mark.parametrize("b,t,a", [(1, 1, 1), (3, 5, 17), (16, 24, 20), (7, 33, 4)])
@pytest.mark.parametrize("clipped,with_entropy,kl_scale", [(True, True, 1.0), (False, True, 2.5), (True, False, 0.5)])
def test_oracle_matches_torch_autograd(ora, b, t, a, clipped, with_entropy, kl_scale) -> None:  # noqa: ANN001
    lp_on, lp_off, v_on, v_off, adv, tgt, ent = make_case(b, t, a, seed=b * 100 + t)
    if not with_entropy:
        ent = None
    flags = C.PPO_CLIPPED_VALUE_LOSS if clipped else 0
    losses, out, d_logp, d_values = ora.ppo_loss(lp_on, lp_off, v_on, v_off, adv, tgt, ent, kl_scale=kl_scale, flags=flags, **HYPER)
    tt = lambda x, g=False: None if x is None else torch.tensor(x, dtype=torch.float64, requires_grad=g)  # noqa: E731
    t_lp_off, t_v_off, t_ent = tt(lp_off, True), tt(v_off, True), tt(ent, True)
    want, loss = torch_ppo_loss(tt(lp_on), t_lp_off, tt(v_on), t_v_off, tt(adv), tt(tgt), t_ent, kl_scale=kl_scale, clipped=clipped, **HYPER)
    loss.backward()
    for i, name in enumerate(("policy", "value", "kl", "entropy")):
        if name in want:
            np.testing.assert_allclose(losses[i], want[name].detach().numpy(), rtol=1e-12, atol=1e-14, err_msg=name)
            np.testing.assert_allclose(out[1 + i], want[name].mean().item(), rtol=1e-12, atol=1e-15)
        else:
            assert not losses[i].any() and out[1 + i] == 0.0
    np.testing.assert_allclose(out[0], loss.item(), rtol=1e-12)
    g_lp = t_lp_off.grad.numpy()
    # the gradient is the same for every action dimension of a (b, t) element
    np.testing.assert_allclose(g_lp, np.broadcast_to(d_logp[..., None], g_lp.shape), rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(t_v_off.grad.numpy(), d_values, rtol=1e-10, atol=1e-16)
    if t_ent is not None:
        np.testing.assert_allclose(t_ent.grad.numpy(), -HYPER["entropy_coef"] / (b * t), rtol=1e-12)


def test_branches_are_all_exercised() -> None:
    """The random case must reach every branch the gradient distinguishes, or the parity tests prove less than they claim."""
    lp_on, lp_off, v_on, v_off, adv, tgt, _ = make_case(16, 24, 20, seed=1624)
    lr = (lp_off - lp_on).sum(-1)
    ratio = np.exp(np.clip(lr, -5, 5))
    assert (np.abs(lr) > 5).sum() >= 2
    for sign in (1, -1):
        assert ((ratio > 1.2) & (sign * adv > 0)).any() and ((ratio < 0.8) & (sign * adv > 0)).any()
    assert ((ratio > 0.8) & (ratio < 1.2)).any()
    d = v_off - v_on
    e2, c2 = (v_off - tgt) ** 2, (v_on + np.clip(d, -0.2, 0.2) - tgt) ** 2
    assert ((np.abs(d) > 0.2) & (c2 > e2)).any() and ((np.abs(d) > 0.2) & (e2 > c2)).any() and (np.abs(d) < 0.2).any()


def test_f32_oracle_close_to_f64(oracle_built, walking_program) -> None:  # noqa: ANN001
    from oracle import OracleEngine

    o32, o64 = OracleEngine(walking_program, "f32"), OracleEngine(walking_program, "f64")
    case = [x.astype(np.float32) for x in make_case(9, 24, 17, seed=5)]
    r32 = o32.ppo_loss(*case, kl_scale=1.0, flags=C.PPO_CLIPPED_VALUE_LOSS, **HYPER)
    r64 = o64.ppo_loss(*case, kl_scale=1.0, flags=C.PPO_CLIPPED_VALUE_LOSS, **HYPER)
    for a32, a64 in zip(r32, r64):
        np.testing.assert_allclose(a32, a64, rtol=2e-5, atol=1e-6)


def test_ppo_loss_has_no_cpu_path() -> None:
    from ksim_b200 import binding, ppo

    z3, z2 = torch.zeros(2, 3, 4), torch.zeros(2, 3)
    with pytest.raises(binding.B200SimError):
        ppo.ppo_loss_and_grads(ppo.PPOInputs(z2, z2), ppo.PPOVariables(z3, z2), ppo.PPOVariables(z3, z2))
