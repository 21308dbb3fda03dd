"""Host logic: task-program layout, blob packing, key tree, randomizers, config validation."""

import math

import numpy as np
import pytest

from ksim_b200 import codes as C
from ksim_b200 import randomization, rng as R
from ksim_b200.blob import build_model_arrays, pack_blob
from ksim_b200.envinit import make_env_init
from ksim_b200.program import EngineConfig, build_program
from tests.helpers import tripod_spec


def test_walking_layout_matches_survey(walking_program) -> None:  # noqa: ANN001
    p = walking_program
    # 16 observations, 389 floats incl. 5 noisy copies; commands 6 + 2 floats
    assert p.obs_size == 389 and len(p.spec.observations) == 16
    assert sorted(k for k in p.obs_slices if k.startswith("noisy_")) == [
        "noisy_imu_acc", "noisy_imu_gyro", "noisy_imu_projected_gravity", "noisy_joint_position", "noisy_joint_velocity"]
    assert p["TI_CMD_SIZE"] == 8 and p.action_size == 17 and p["TI_NSUB"] == 5
    assert len(p.reward_components) == 16  # 1 + 3 + 4 + 2 + 3 + 1 + 1 + 1
    # actor / critic inputs of examples/walking.py:560-561 are assembled from these slices
    o = p.obs_slices
    actor = sum(o[k][1] for k in ("noisy_joint_position", "noisy_joint_velocity", "noisy_imu_projected_gravity")) + 6
    assert actor == 43
    assert p.state_size % 4 == 0 and p.rec_size % 4 == 0 and p.key_size % 4 == 0
    # algorithmic bytes per env-step: read state + keys + rand + action, write state + keys + record
    m = p.spec.model
    assert p.rec_size >= 389 + 8 + 1 + m.nq + m.nv + 7 * m.nbody + m.nu + 2 + 17 + 3 + 3


def test_blob_roundtrip_header(walking_program) -> None:  # noqa: ANN001
    p = walking_program
    blob = pack_blob(p.spec.model, p.ti, p.tf)
    h = np.frombuffer(blob[: C.MH_SIZE * 4], dtype=np.int32)
    assert h[C.MH_MAGIC] == C.B2S_MODEL_MAGIC and h[C.MH_NV] == 23 and h[C.MH_NBODY] == 17
    assert len(blob) == 4 * (C.MH_SIZE + h[C.MH_NI] + h[C.MH_NF] + h[C.MH_NTI] + h[C.MH_NTF])
    assert h[C.MH_NCON] == 2 and h[C.MH_NEFC_DENSE] == 2 and h[C.MH_NLIMIT] == 17 and h[C.MH_NLEVEL] == 6
    header, mi, _ = build_model_arrays(p.spec.model)
    sub_end = mi[header[C.MH_BODY_SUBTREE_END]:][: 17]
    assert sub_end[1] == 17 and sub_end[0] == 17  # torso's subtree is every body


def test_engine_config_validation() -> None:
    with pytest.raises(ValueError):
        EngineConfig(dt=0.004, ctrl_dt=0.02, action_latency_range=(-1.0, 0.0))
    with pytest.raises(ValueError):
        EngineConfig(dt=0.004, ctrl_dt=0.02, action_latency_range=(0.2, 0.1))
    with pytest.raises(ValueError):
        EngineConfig(dt=0.004, ctrl_dt=0.02, drop_action_prob=1.0)
    cfg = EngineConfig(dt=0.004, ctrl_dt=0.02, action_latency_range=(0.003, 0.01), actuator_update_dt=0.008)
    assert cfg.phys_steps_per_ctrl_steps == 5 and cfg.phys_steps_per_actuator_step == 2
    assert cfg.max_action_latency_step == pytest.approx(2.5)


def test_randomizers_follow_reference_formulas(walking_spec) -> None:  # noqa: ANN001
    m = walking_spec.model
    keys = R.split(R.prng_key(0), 5)
    out = randomization.ArmatureRandomizer()(m, keys)["dof_armature"]
    assert out.shape == (5, m.nv) and out.dtype == np.float32
    np.testing.assert_array_equal(out[:, :6], np.broadcast_to(m.dof_armature[:6].astype(np.float32), (5, 6)))
    ratio = out[:, 6:] / m.dof_armature[6:]
    assert (ratio >= 0.95 - 1e-6).all() and (ratio <= 1.05 + 1e-6).all()
    # same key -> same value (tests/test_randomizations.py:52-104 checks exactly this determinism)
    np.testing.assert_array_equal(out, randomization.ArmatureRandomizer()(m, keys)["dof_armature"])
    mass = randomization.MassMultiplicationRandomizer.from_body_name(m, "torso")(m, keys)["body_mass"]
    assert (mass[:, 2:] == m.body_mass[2:].astype(np.float32)).all() and (mass[:, 1] <= m.body_mass[1]).all()
    with pytest.raises(ValueError):
        randomization.MassMultiplicationRandomizer.from_body_name(m, "nope")
    both = randomization.AllBodiesInertiaRandomizer()(m, keys)
    assert set(both) == {"body_mass", "body_inertia"}


def test_env_keys_are_invariant_to_sharding(walking_program, walking_randomizers) -> None:  # noqa: ANN001
    full = make_env_init(walking_program, walking_randomizers, 16, seed=5)
    lo = make_env_init(walking_program, walking_randomizers, 8, seed=5, env_offset=0, total_envs=16)
    hi = make_env_init(walking_program, walking_randomizers, 8, seed=5, env_offset=8, total_envs=16)
    np.testing.assert_array_equal(full.init_keys, np.concatenate([lo.init_keys, hi.init_keys]))
    np.testing.assert_array_equal(full.rand, np.concatenate([lo.rand, hi.rand]))
    assert len({tuple(k) for k in full.init_keys[:, 8:10]}) == 16


def test_reward_sign_rule_and_scale_encoding() -> None:
    from ksim_b200 import rewards
    from ksim_b200.scales import LinearScale

    assert rewards.ForcePenalty(force_obs="x", ctrl_dt=0.02, expected_weight=1.0, scale=1.0).is_penalty
    assert not rewards.StayAliveReward(scale=1.0).is_penalty

    class Oops(rewards.StayAliveReward):
        pass

    with pytest.raises(ValueError):
        _ = Oops(scale=1.0).is_penalty
    prog = build_program(tripod_spec(rewards={"alive": rewards.StayAliveReward(scale=LinearScale(scale=2.0, bias=1.0))}))
    io = prog.ti[prog["TI_REW_TAB"] + C.TAB_IOFF]
    assert prog.ti[io] == C.SCALE_LINEAR
    assert math.isclose(prog.tf[prog.ti[prog["TI_REW_TAB"] + C.TAB_FOFF]], 2.0)


def test_unknown_names_fail_loudly() -> None:
    from ksim_b200 import observation, rewards

    spec = tripod_spec(rewards={"f": rewards.ForcePenalty(force_obs="missing", ctrl_dt=0.02, expected_weight=1.0, scale=1.0)})
    with pytest.raises(KeyError):
        build_program(spec)
    with pytest.raises(ValueError):
        observation.SensorObservation.create(physics_model=spec.model, sensor_name="nope")
