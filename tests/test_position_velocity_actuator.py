"""``PositionVelocityActuator`` (ksim/actuators.py:241-293): action = [position targets | velocity targets], ``ctrl = kps (p* - q)
+ kds (v* - qdot)``, torque noise, clip.  CPU: properties of the oracle's implementation; GPU: CUDA == oracle through the fused
control step (generic instantiation: the action is 2 nu wide)."""

import numpy as np
import pytest

from ksim_b200 import commands, observation, resets
from ksim_b200.actuators import PositionActuators, PositionVelocityActuator
from ksim_b200.noise import AdditiveGaussianNoise
from ksim_b200.program import EngineConfig, TaskSpec, build_program
from ksim_b200.types import Metadata
from tests.helpers import tripod_model


def _spec(kp: float, kd: float, clip: float | None = 3.0, noisy: bool = False) -> TaskSpec:
    model = tripod_model()
    md = Metadata.from_model(model, kp=kp, kd=kd)
    if clip is not None:
        for jm in md.joint_name_to_metadata.values():
            jm.soft_torque_limit = clip
    noise = dict(pos_action_noise=AdditiveGaussianNoise(std=0.02), vel_action_noise=AdditiveGaussianNoise(std=0.1),
                 torque_noise=AdditiveGaussianNoise(std=0.05)) if noisy else {}
    act = PositionVelocityActuator(physics_model=model, metadata=md, **noise)
    return TaskSpec(model=model, config=EngineConfig(dt=0.004, ctrl_dt=0.02), actuators=act,
                    resets=[resets.RandomJointPositionReset.create(model)], events={},
                    observations={"joint_position": observation.JointPositionObservation(), "joint_velocity": observation.JointVelocityObservation()},
                    commands={"vec": commands.FloatVectorCommand(ranges=((0.0, 1.0),))}, rewards={}, terminations={})


def _run_oracle(spec: TaskSpec, actions: np.ndarray, level: float = 1.0):  # noqa: ANN202
    from ksim_b200.envinit import make_env_init
    from oracle import OracleEngine

    prog = build_program(spec)
    n = actions.shape[1]
    init = make_env_init(prog, {}, n, seed=2, level=level)
    ora = OracleEngine(prog, "f32")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((actions.shape[0] + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    ora.rollout(actions, init.rand, st, keys, rec)
    return prog, rec, init


def test_action_is_position_and_velocity_targets(oracle_built) -> None:  # noqa: ANN001
    spec = _spec(kp=1e-6, kd=1.0)
    prog = build_program(spec)
    assert prog.action_size == 2 * spec.model.nu == 6
    a = np.zeros((1, 2, 6), np.float32)
    a[0, 0, 3:] = 1e3          # env 0: huge positive velocity targets  -> +clip on every actuator
    a[0, 1, 3:] = -1e3         # env 1: huge negative                   -> -clip
    prog, rec, _ = _run_oracle(spec, a)
    c = prog["TI_R_CTRL"]
    np.testing.assert_allclose(rec[0, 0, c : c + 3], 3.0)
    np.testing.assert_allclose(rec[0, 1, c : c + 3], -3.0)
    # the position half drives the same actuators through kps
    spec = _spec(kp=50.0, kd=1e-6, clip=None)
    a = np.zeros((1, 1, 6), np.float32)
    a[0, 0, :3] = 0.4
    prog, rec, _ = _run_oracle(spec, a)
    q, c = prog["TI_R_QPOS"], prog["TI_R_CTRL"]
    # torque of the last sub-step = kp (target - q before that sub-step): same sign as (target - q after it) unless it overshot
    assert (np.sign(rec[0, 0, c : c + 3]) == np.sign(0.4 - rec[0, 0, q + 7 : q + 10])).all()
    assert np.abs(rec[0, 0, c : c + 3]).max() > 1.0
    # a PositionActuators task takes nu-wide actions: the two classes are told apart by the program
    pa = _spec(kp=50.0, kd=1.0)
    pa.actuators = PositionActuators(physics_model=pa.model, metadata=Metadata.from_model(pa.model, kp=50.0, kd=1.0))
    assert build_program(pa).action_size == 3


@pytest.mark.gpu
def test_position_velocity_actuator_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    spec = _spec(kp=20.0, kd=0.5, noisy=True)
    n, t = 48, 6
    actions = (0.4 * np.random.RandomState(3).randn(t, n, 6)).astype(np.float32)
    prog, rec_o, init = _run_oracle(spec, actions)
    eng = B200Engine(spec, n, device="cuda:0", seed=2, curriculum_level=1.0)
    assert eng.info("VARIANT") in (1, 2) and eng.NA == 6   # a run-time-sized instantiation (action 2 nu wide)
    rec = eng.rollout(torch.from_numpy(actions).cuda()).cpu().numpy()
    q, v, c, d = prog["TI_R_QPOS"], prog["TI_R_QVEL"], prog["TI_R_CTRL"], prog["TI_R_DONE"]
    m = spec.model
    np.testing.assert_array_equal(rec[:t, :, d], rec_o[:t, :, d])
    ok = np.ones(n, bool)
    for s in range(t):     # chained rollout: an env whose contact flips on an ulp drops out of the comparison from then on
        dq = np.abs(rec[s, :, q : q + m.nq] - rec_o[s, :, q : q + m.nq]).max(axis=1)
        ok &= dq < 5e-4
    assert ok.mean() >= 0.9, f"{(~ok).sum()} of {n} envs drift"
    np.testing.assert_allclose(rec[:2, :, c : c + m.nu][:, ok], rec_o[:2, :, c : c + m.nu][:, ok], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(rec[:2, :, v : v + m.nv][:, ok], rec_o[:2, :, v : v + m.nv][:, ok], rtol=5e-3, atol=2e-2)
