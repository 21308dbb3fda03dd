"""K-Bot walking -- the physics, actuator, noise, randomization, event and reset set-up of the reference's
examples/kbot/train.py:1089-1273 (BASELINE.json configs 3 and 5), expressed with this package's mirror classes.

What is the same as the reference task: the robot model (examples/kbot/robot/kbot-headless/robot.mjcf + metadata.json,
compiled to ``assets/kbot.npz`` by scripts/compile_reference_models.py), ``PositionActuators`` with the per-joint
gains / soft torque limits from the metadata, action noise sigma 0.02, kp/kd scale 1.4, torque-limit scale 1.5
(train.py:1095-1103), latency 3-10 ms, 5 % action drop, zero-offset +-3 deg (train.py:1771-1774), the force-push event
(train.py:1132-1142), the reset list (train.py:1144-1151), the core observations with their noise (train.py:1153-1202),
``NotUpright`` / ``EpisodeLength`` terminations (train.py:1255-1266), curriculum pinned at level 1 (train.py:1268-1273).

What differs, and why:
* the floor is a plain plane (the reference takes its scene from the un-vendored ``mujoco_scenes`` package);
* ``FloorFrictionRandomizer`` is omitted: the foot capsules have ``priority=1`` over the floor's 0, so MuJoCo takes the
  capsule's friction for the pair and the floor's value never enters the contact model (robot.mjcf:23-25);
* ``CollisionBodyRandomizer`` (per-env capsule size / position) is not supported by the engine yet;
* the plugins that live in the example file itself rather than in ksim -- ``UnifiedCommand``, ``COMDistanceObservation``,
  ``TerrainBadZTermination`` and the eleven task-local rewards (train.py:128-834) -- are replaced by their closest ksim
  counterparts (``LinearVelocityCommand`` / ``AngularVelocityCommand``, ``BadZTermination``, core rewards).
"""

from __future__ import annotations

__all__ = ["KbotConfig", "JOINT_BIASES", "load_kbot", "make_spec", "RANDOMIZERS"]

import json
import math
from dataclasses import dataclass
from pathlib import Path

from ksim_b200 import commands, events, observation, randomization, resets, rewards, terminations
from ksim_b200.actuators import PositionActuators
from ksim_b200.mjcf import Model, load_model
from ksim_b200.noise import AdditiveGaussianNoise, AdditiveUniformNoise
from ksim_b200.program import EngineConfig, TaskSpec
from ksim_b200.types import Metadata

_ASSETS = Path(__file__).resolve().parent.parent / "assets"

# examples/kbot/train.py:26-47 (policy output order; neutral pose)
JOINT_BIASES: dict[str, float] = {
    "dof_left_hip_pitch_04": math.radians(20.0), "dof_left_hip_roll_03": 0.0, "dof_left_hip_yaw_03": 0.0,
    "dof_left_knee_04": math.radians(50.0), "dof_left_ankle_02": math.radians(-30.0),
    "dof_right_hip_pitch_04": math.radians(-20.0), "dof_right_hip_roll_03": 0.0, "dof_right_hip_yaw_03": 0.0,
    "dof_right_knee_04": math.radians(-50.0), "dof_right_ankle_02": math.radians(30.0),
    "dof_right_shoulder_pitch_03": 0.0, "dof_right_shoulder_roll_03": math.radians(-10.0), "dof_right_shoulder_yaw_02": 0.0,
    "dof_right_elbow_02": math.radians(90.0), "dof_right_wrist_00": 0.0,
    "dof_left_shoulder_pitch_03": 0.0, "dof_left_shoulder_roll_03": math.radians(10.0), "dof_left_shoulder_yaw_02": 0.0,
    "dof_left_elbow_02": math.radians(-90.0), "dof_left_wrist_00": 0.0,
}


@dataclass
class KbotConfig:
    """examples/kbot/train.py:1751-1775 (rollout-relevant fields only)."""

    num_envs: int = 4096
    rollout_length_seconds: float = 2.0
    dt: float = 0.004
    ctrl_dt: float = 0.02
    iterations: int = 8
    ls_iterations: int = 8
    tolerance: float = 1e-10
    action_latency_range: tuple[float, float] = (0.003, 0.01)
    drop_action_prob: float = 0.05
    zero_offset_mag: float | None = math.radians(3.0)
    action_clip_max: float = 1e3
    gamma: float = 0.94
    lam: float = 0.94

    @property
    def rollout_length_frames(self) -> int:
        return int(round(self.rollout_length_seconds / self.ctrl_dt))


def load_kbot(config: KbotConfig) -> tuple[Model, Metadata]:
    model = load_model(_ASSETS / "kbot.npz")
    o = model.opt
    o.timestep, o.iterations, o.ls_iterations, o.tolerance = config.dt, config.iterations, config.ls_iterations, config.tolerance
    o.integrator, o.solver, o.disable_eulerdamp = "implicitfast", "newton", True
    metadata = Metadata.from_json(json.loads((_ASSETS / "kbot_metadata.json").read_text()))
    return model, metadata


RANDOMIZERS = {
    "static_friction": lambda m: randomization.StaticFrictionRandomizer(),
    "armature": lambda m: randomization.ArmatureRandomizer(),
    "joint_damping": lambda m: randomization.JointDampingRandomizer(scale_lower=0.5, scale_upper=2.5),
    "all_body_COM": lambda m: randomization.AllBodiesCOMRandomizer(scale=0.05),
    "all_body_inertia": lambda m: randomization.AllBodiesInertiaRandomizer(scale=0.15),
}


def make_spec(config: KbotConfig | None = None) -> TaskSpec:
    config = KbotConfig() if config is None else config
    model, metadata = load_kbot(config)
    eng = EngineConfig(
        dt=config.dt, ctrl_dt=config.ctrl_dt, action_latency_range=config.action_latency_range,
        drop_action_prob=config.drop_action_prob, zero_offset_mag=config.zero_offset_mag, action_clip_max=config.action_clip_max,
    )
    left, right = "LFootBushing_GPF_1517_12", "RFootBushing_GPF_1517_12"
    obs = {
        "joint_position": observation.JointPositionObservation(noise=AdditiveGaussianNoise(std=math.radians(3))),
        "joint_velocity": observation.JointVelocityObservation(noise=AdditiveUniformNoise(mag=math.radians(15))),
        "actuator_force": observation.ActuatorForceObservation(),
        "center_of_mass_inertia": observation.CenterOfMassInertiaObservation(),
        "center_of_mass_velocity": observation.CenterOfMassVelocityObservation(),
        "base_position": observation.BasePositionObservation(),
        "base_orientation": observation.BaseOrientationObservation(),
        "base_linear_velocity": observation.BaseLinearVelocityObservation(),
        "base_angular_velocity": observation.BaseAngularVelocityObservation(),
        "base_linear_acceleration": observation.BaseLinearAccelerationObservation(),
        "base_angular_acceleration": observation.BaseAngularAccelerationObservation(),
        "actuator_acceleration": observation.ActuatorAccelerationObservation(),
        "imu_gyro": observation.SensorObservation.create(physics_model=model, sensor_name="imu_gyro",
                                                         noise=AdditiveGaussianNoise(std=math.radians(10))),
        "left_foot_touch": observation.SensorObservation.create(physics_model=model, sensor_name="left_foot_touch"),
        "right_foot_touch": observation.SensorObservation.create(physics_model=model, sensor_name="right_foot_touch"),
        "feet_position": observation.BodyPositionObservation.create(physics_model=model, body_names=(left, right)),
        "imu_projected_gravity": observation.ProjectedGravityObservation.create(
            physics_model=model, framequat_name="imu_site_quat", noise=AdditiveGaussianNoise(std=math.radians(3)),
            min_lag=0.0, max_lag=0.75, imu_bias=math.radians(4)),
        "projected_gravity": observation.ProjectedGravityObservation.create(physics_model=model, framequat_name="imu_site_quat"),
    }
    cmds = {
        "linvel": commands.LinearVelocityCommand(min_vel=0.0, max_vel=1.2, max_yaw=math.pi / 4, zero_prob=0.2,
                                                 switch_prob=config.ctrl_dt / 5, ctrl_dt=config.ctrl_dt, linear_accel=0.5,
                                                 angular_accel=0.5),
        "angvel": commands.AngularVelocityCommand(min_vel=0.0, max_vel=1.0, zero_prob=0.2, switch_prob=config.ctrl_dt / 5,
                                                  ctrl_dt=config.ctrl_dt, angular_accel=0.5),
    }
    rews = {
        "stay_alive": rewards.StayAliveReward(scale=1.0),
        "upright": rewards.UprightReward(scale=0.2),
        "linvel": rewards.LinearVelocityReward(cmd="linvel", scale=0.2),
        "angvel": rewards.AngularVelocityReward(cmd="angvel", scale=0.1),
    }
    terms = {
        "bad_z": terminations.BadZTermination(min_z=0.35, max_z=3.0),
        "not_upright": terminations.NotUprightTermination(max_radians=math.radians(45)),
        "episode_length": terminations.EpisodeLengthTermination(max_length_sec=24),
    }
    return TaskSpec(
        model=model, config=eng,
        actuators=PositionActuators(physics_model=model, metadata=metadata, action_noise=AdditiveGaussianNoise(std=0.02),
                                    torque_noise=AdditiveGaussianNoise(std=0.0), kp_scale=1.4, kd_scale=1.4, torque_limit_scale=1.5),
        resets=[
            resets.RandomJointPositionReset.create(model, JOINT_BIASES, scale=0.1),
            resets.RandomJointVelocityReset(scale=2.0),
            resets.RandomBaseVelocityXYReset(scale=0.2),
            resets.RandomHeadingReset(),
            # the example's own PlaneXYPositionReset (train.py:817-834) redraws x, y in +-0.1 m and leaves z alone: the core
            # class with the floor height set to qpos0's z and no clipping does exactly that
            resets.PlaneXYPositionReset(bounds=(0.0, 0.0, float(model.qpos0[2])), padded_bounds=(-10.0, 10.0, -10.0, 10.0),
                                        x_range=0.1, y_range=0.1, robot_base_height=0.0),
        ],
        events={
            "force_push": events.ForcePushEvent.from_body_name(model=model, body_name="base", max_force=200.0, max_torque=10.0,
                                                               duration_range=(0.1, 0.5), interval_range=(0.0, 6.0)),
        },
        observations=obs, commands=cmds, rewards=rews, terminations=terms,
        randomized_fields=("dof_frictionloss", "dof_armature", "dof_damping", "body_ipos", "body_mass", "body_inertia"),
    )
