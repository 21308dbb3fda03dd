"""GAE oracle (compute_ppo_inputs, ksim/task/ppo.py:54-121) against an independent vectorised numpy statement
of the same equations, plus the edge cases the path has: T = 1, all done, success bootstrap, normalisation."""

import numpy as np
import pytest

from ksim_b200 import codes as C


def numpy_gae(values, rewards, dones, successes, gamma, lam, normalize=False, monte_carlo=False):  # noqa: ANN001, ANN201
    values, rewards = np.asarray(values, np.float64), np.asarray(rewards, np.float64)
    shifted = np.concatenate([values[:, 1:], values[:, -1:]], axis=1)
    trunc = np.where(successes, 1.0, 0.0)
    boot = (1 - gamma) * rewards + gamma * values * trunc
    mask = np.where(dones, 0.0, 1.0)
    deltas = boot + gamma * shifted * mask - values
    ret, adv = np.zeros_like(values), np.zeros_like(values)
    r_next, a_next = np.zeros(values.shape[0]), np.zeros(values.shape[0])
    for t in range(values.shape[1] - 1, -1, -1):
        r_next = boot[:, t] + gamma * mask[:, t] * r_next
        a_next = deltas[:, t] + gamma * lam * mask[:, t] * a_next
        ret[:, t], adv[:, t] = r_next, a_next
    targets = ret if monte_carlo else adv + values
    if normalize:
        adv = (adv - adv.mean()) / max(adv.std(), 1e-6)
    return adv, targets, ret


@pytest.fixture(scope="module")
def ora(oracle_built, walking_program):  # noqa: ANN001, ANN201
    from oracle import OracleEngine

    return OracleEngine(walking_program, "f64")


@pytest.mark.parametrize("b,t", [(1, 1), (3, 2), (8, 24), (5, 100)])
@pytest.mark.parametrize("normalize,monte_carlo", [(False, False), (True, False), (False, True)])
def test_oracle_gae_matches_numpy(ora, b, t, normalize, monte_carlo) -> None:  # noqa: ANN001
    rs = np.random.RandomState(b * 100 + t)
    values, rewards = rs.randn(b, t), rs.randn(b, t)
    dones = rs.rand(b, t) < 0.15
    successes = dones & (rs.rand(b, t) < 0.4)
    flags = (C.GAE_NORMALIZE_ADVANTAGES if normalize else 0) | (C.GAE_MONTE_CARLO_RETURNS if monte_carlo else 0)
    adv, tgt, ret = ora.gae(values, rewards, dones, successes, 0.97, 0.95, flags)
    wa, wt, wr = numpy_gae(values, rewards, dones, successes, 0.97, 0.95, normalize, monte_carlo)
    np.testing.assert_allclose(ret, wr, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(adv, wa, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(tgt, wt, rtol=1e-12, atol=1e-12)


def test_gae_semantics(ora) -> None:  # noqa: ANN001
    # all done, no success: no bootstrap at all -> returns are the scaled rewards themselves
    v, r = np.full((2, 4), 3.0), np.arange(8.0).reshape(2, 4)
    ones, zeros = np.ones((2, 4), bool), np.zeros((2, 4), bool)
    adv, tgt, ret = ora.gae(v, r, ones, zeros, 0.9, 0.8, 0)
    np.testing.assert_allclose(ret, 0.1 * r)
    np.testing.assert_allclose(adv, 0.1 * r - v)
    # successful termination bootstraps one step with the state's own value (ppo.py:82-83)
    adv, tgt, ret = ora.gae(v, r, ones, ones, 0.9, 0.8, 0)
    np.testing.assert_allclose(ret, 0.1 * r + 0.9 * v)
    # the last step bootstraps with its own value: V_T = V_{T-1} (ppo.py:79)
    adv, _, _ = ora.gae(np.array([[2.0]]), np.array([[1.0]]), np.zeros((1, 1), bool), np.zeros((1, 1), bool), 0.5, 0.5, 0)
    np.testing.assert_allclose(adv, [[0.5 * 1.0 + 0.5 * 2.0 - 2.0]])
    # constant advantages normalise to zero with the 1e-6 std clip (ppo.py:118)
    adv, _, _ = ora.gae(np.zeros((3, 1)), np.ones((3, 1)), np.ones((3, 1), bool), np.zeros((3, 1), bool), 0.5, 0.5,
                        C.GAE_NORMALIZE_ADVANTAGES)
    np.testing.assert_allclose(adv, 0.0)
