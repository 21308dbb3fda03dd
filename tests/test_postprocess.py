"""The ``postprocess_trajectory`` hook (ksim/task/rl.py:1590, amp.py:327-358, teacher.py:113-136): aux outputs
written between the rollout and the rewards pass, and ``AMPReward`` (amp.py:100-112) reading them."""

import numpy as np
import pytest

from ksim_b200 import rewards
from ksim_b200.program import build_program
from tests.helpers import tripod_spec


def _amp_spec():  # noqa: ANN202
    spec = tripod_spec(rewards={"amp": rewards.AMPReward(scale=2.0), "alive": rewards.StayAliveReward(scale=1.0)})
    spec.aux_outputs = {rewards.DISCRIMINATOR_OUTPUT_KEY: 1, "teacher": 3}
    return spec


def test_aux_region_layout_and_missing_declaration() -> None:
    prog = build_program(_amp_spec())
    assert prog["TI_AUX_SIZE"] == 4 and prog.aux_slices == {rewards.DISCRIMINATOR_OUTPUT_KEY: (0, 1), "teacher": (1, 3)}
    assert prog["TI_R_AUX"] + 4 <= prog.rec_size and prog["TI_R_AUX"] >= prog["TI_R_TERM"]
    base = build_program(tripod_spec())
    assert base["TI_AUX_SIZE"] == 0                      # tasks without aux outputs keep their record size
    with pytest.raises(ValueError, match="AMPReward auxiliary output is missing"):
        build_program(tripod_spec(rewards={"amp": rewards.AMPReward(scale=1.0)}))


def test_amp_reward_oracle(oracle_built) -> None:  # noqa: ANN001
    from oracle import OracleEngine

    prog = build_program(_amp_spec())
    ora = OracleEngine(prog, "f64")
    t = 9
    logits = np.linspace(-2.0, 4.0, t)
    rec = np.zeros((t, 1, prog.rec_size))
    rec[:, 0, prog["TI_R_AUX"]] = logits
    total, comps = ora.rewards(rec, np.zeros(1), np.zeros((1, max(prog["TI_REW_CARRY_SIZE"], 1))))
    want = np.maximum(0.0, 1.0 - 0.25 * (logits - 1.0) ** 2)          # amp.py:111
    np.testing.assert_allclose(comps[prog.reward_components.index("amp")][0], 2.0 * want, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_postprocess_hook_feeds_the_rewards_pass(oracle_built) -> None:  # noqa: ANN001
    """A toy discriminator over the trajectory's joint positions as the hook; the CUDA rewards pass equals the formula and
    the oracle on the same records."""
    import torch

    from ksim_b200.engine import B200Engine
    from oracle import OracleEngine

    n, t = 64, 12
    eng = B200Engine(_amp_spec(), n, device="cuda:0", seed=1)
    torch.manual_seed(0)
    disc = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.Tanh(), torch.nn.Linear(16, 1)).cuda()
    rec = eng.rollout(0.5 * torch.randn((t, n, eng.NA), device="cuda:0"))
    calls = []

    def hook(traj):  # noqa: ANN001, ANN202
        calls.append(tuple(traj.qpos.shape))
        with torch.no_grad():
            return {rewards.DISCRIMINATOR_OUTPUT_KEY: 3.0 * disc(traj.qpos[..., 7:])[..., 0]}

    total, comps = eng.postprocess_and_rewards(rec, postprocess_trajectory=hook)
    assert calls == [(t, n, 10)] and total.shape == (n, t)
    with torch.no_grad():
        logits = 3.0 * disc(eng.view(rec[:t]).qpos[..., 7:])[..., 0]      # [T, N]
    torch.testing.assert_close(eng.view(rec[:t]).aux_outputs[rewards.DISCRIMINATOR_OUTPUT_KEY][..., 0], logits, rtol=0, atol=0)
    want = 2.0 * torch.clamp(1.0 - 0.25 * (logits - 1.0) ** 2, min=0.0)
    i = eng.program.reward_components.index("amp")
    torch.testing.assert_close(comps[i], want.t(), rtol=1e-6, atol=1e-6)
    assert float(want.max()) > 0.1 and float((want == 0).float().mean()) > 0.01       # both branches of the max
    ora = OracleEngine(eng.program, "f32")
    o_total, o_comps = ora.rewards(rec[:t].cpu().numpy(), np.zeros(n, np.float32), np.zeros((n, max(eng.program["TI_REW_CARRY_SIZE"], 1)), np.float32))
    np.testing.assert_allclose(total.cpu().numpy(), o_total, rtol=1e-5, atol=1e-5)
    with pytest.raises(KeyError):
        eng.postprocess_and_rewards(rec, postprocess_trajectory=lambda traj: {"nope": torch.zeros(t, n, device="cuda:0")})
