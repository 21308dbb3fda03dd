"""K-Bot walking -- the task of the reference's examples/kbot/train.py:1089-1273 (BASELINE.json configs 3 and 5), expressed
with this package's mirror classes, INCLUDING the plugins the example file defines itself (train.py:128-834):
``UnifiedCommand``, ``FeetPositionObservation``, ``BaseHeightObservation``, ``COMDistanceObservation``,
``TerrainBadZTermination``, the local ``PlaneXYPositionReset`` and the eleven task rewards.  Their arithmetic runs in
the CUDA engine (``OBS_* / CMD_UNIFIED / TERM_TERRAIN_BAD_Z / REW_KBOT_*`` in include/b200sim_codes.h).

Same as the reference task: the robot model (examples/kbot/robot/kbot-headless/robot.mjcf + metadata.json, compiled to
``assets/kbot.npz`` by scripts/compile_reference_models.py), ``PositionActuators`` with the per-joint gains / soft torque
limits from the metadata, action noise sigma 0.02, kp/kd scale 1.4, torque-limit scale 1.5 (train.py:1095-1103), latency
3-10 ms, 5 % action drop, zero-offset +-3 deg (train.py:1771-1774), the force-push event (train.py:1132-1142), the reset
list (train.py:1144-1151), all twenty observations with their noise (train.py:1153-1202), the unified command
(train.py:1204-1221), the eleven rewards with their scales (train.py:1223-1253), the three terminations
(train.py:1255-1266), curriculum pinned at level 1 (train.py:1268-1273).

What differs, and why:
* the floor is a plain plane (the reference takes its scene from the un-vendored ``mujoco_scenes`` package);
* ``FloorFrictionRandomizer`` is accepted and checked to be a no-op: the foot capsules have ``priority=1`` over the floor's
  0, so MuJoCo takes the capsule's friction for the pair and the floor's value never enters the contact model
  (robot.mjcf:23-25; ``randomization.pair_friction_unaffected``).

``make_spec(core_only=True)`` keeps the earlier reduced task (core ksim commands / rewards / ``BadZTermination``) that
the round-1 measurements were taken with.
"""

from __future__ import annotations

__all__ = [
    "KbotConfig", "JOINT_BIASES", "load_kbot", "make_spec", "RANDOMIZERS", "UnifiedCommand", "FeetPositionObservation",
    "BaseHeightObservation", "COMDistanceObservation", "TerrainBadZTermination", "PlaneXYPositionReset",
    "SingleFootContactReward", "NoContactPenalty", "FeetAirtimeReward", "ArmPositionReward", "LinearVelocityTrackingReward",
    "AngularVelocityReward", "XYOrientationReward", "TerrainBaseHeightReward", "FeetOrientationReward", "COMDistanceReward",
    "BaseAccelerationReward",
]

import json
import math
from dataclasses import dataclass
from pathlib import Path

import attrs

from ksim_b200 import codes as C
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
    # the reference-shaped capped Newton iteration ships for this task (EngineConfig.solver); "fast" is the opt-in whose
    # deviation from the reference's iterate tests/test_gpu_parity.py::test_kbot_parity_resynchronised counts
    solver: str = "reference"

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



# ------------------------------------------------------------------------------------------------------------------
# plugins the reference keeps in the example file (examples/kbot/train.py:128-834)
# ------------------------------------------------------------------------------------------------------------------


@attrs.define(frozen=True, kw_only=True)
class UnifiedCommand(commands.Command):
    """train.py:701-775: [vx, vy, wz, base height, roll, pitch, 10 arm joints], one of six modes drawn per resample."""

    vx_range: tuple[float, float] = attrs.field()
    vy_range: tuple[float, float] = attrs.field()
    wz_range: tuple[float, float] = attrs.field()
    bh_range: tuple[float, float] = attrs.field()
    rx_range: tuple[float, float] = attrs.field()
    ry_range: tuple[float, float] = attrs.field()
    arms_range: tuple[tuple[float, ...], ...] = attrs.field()
    ctrl_dt: float = attrs.field()
    switch_prob: float = attrs.field()

    def encode(self) -> tuple[int, list[int], list[float], int]:
        lo, hi = self.arms_range
        if len(lo) != 10 or len(hi) != 10:
            raise ValueError("UnifiedCommand expects ten arm joints")
        floats = [*self.vx_range, *self.vy_range, *self.wz_range, *self.bh_range, *self.rx_range, *self.ry_range, *lo, *hi,
                  self.switch_prob]
        return C.CMD_UNIFIED, [], [float(v) for v in floats], 16


@attrs.define(frozen=True, kw_only=True)
class FeetPositionObservation(observation.Observation):
    """train.py:653-689: both feet relative to the base, rotated into the base's yaw frame."""

    base_idx: int = attrs.field()
    foot_left_idx: int = attrs.field()
    foot_right_idx: int = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, base_body_name: str, foot_left_body_name: str,
               foot_right_body_name: str) -> "FeetPositionObservation":
        return cls(base_idx=physics_model.id("body", base_body_name), foot_left_idx=physics_model.id("body", foot_left_body_name),
                   foot_right_idx=physics_model.id("body", foot_right_body_name))

    def encode(self, model: Model) -> observation.Encoded:
        return C.OBS_FEET_POSITION, [self.base_idx, self.foot_left_idx, self.foot_right_idx], [], 6, 0


@attrs.define(frozen=True, kw_only=True)
class BaseHeightObservation(observation.Observation):
    """train.py:693-697: ``xpos[1, 2:]``."""

    def encode(self, model: Model) -> observation.Encoded:
        return C.OBS_BASE_HEIGHT, [], [], 1, 0


@attrs.define(frozen=True, kw_only=True)
class COMDistanceObservation(observation.Observation):
    """train.py:500-649: distance between ``subtree_com[2]`` and the centroid of the contact points' convex hull, or -1 when
    fewer than three distinct ``geom2`` values appear in the contact list.  MJX keeps every candidate contact of every
    pair, so that count is a property of the model: it is evaluated here and handed to the kernel as a flag."""

    def encode(self, model: Model) -> observation.Encoded:
        distinct = len(set(int(g) for g in model.arrays["pair_geom2"]))
        return C.OBS_COM_DISTANCE, [int(distinct >= 3)], [], 1, 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return ()


@attrs.define(frozen=True, kw_only=True)
class TerrainBadZTermination(terminations.Termination):
    """train.py:779-813: base height above the lower foot below ``unhealthy_z`` -> -1."""

    base_idx: int = attrs.field()
    foot_left_idx: int = attrs.field()
    foot_right_idx: int = attrs.field()
    unhealthy_z: float = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, base_body_name: str, foot_left_body_name: str, foot_right_body_name: str,
               unhealthy_z: float) -> "TerrainBadZTermination":
        return cls(base_idx=physics_model.id("body", base_body_name), foot_left_idx=physics_model.id("body", foot_left_body_name),
                   foot_right_idx=physics_model.id("body", foot_right_body_name), unhealthy_z=unhealthy_z)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_TERRAIN_BAD_Z, [self.base_idx, self.foot_left_idx, self.foot_right_idx], [float(self.unhealthy_z)]


@attrs.define(frozen=True, kw_only=True)
class PlaneXYPositionReset(resets.Reset):
    """train.py:817-834: redraws x, y in +-range and leaves z alone.  No earlier reset of the task touches z, so this is
    the core ``PlaneXYPositionReset`` (resets.py:93-119) with the plane at ``qpos0``'s z and no clipping."""

    x_range: float = attrs.field(default=1.0)
    y_range: float = attrs.field(default=1.0)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        core = resets.PlaneXYPositionReset(bounds=(0.0, 0.0, float(model.qpos0[2])), padded_bounds=(-1e9, 1e9, -1e9, 1e9),
                                           x_range=self.x_range, y_range=self.y_range, robot_base_height=0.0)
        return core.encode(model)


def _touch(ctx: rewards.EncodeContext, name: str) -> int:
    return ctx.obs_offset(name)


@attrs.define(frozen=True, kw_only=True)
class SingleFootContactReward(rewards.StatefulReward):
    """train.py:128-156."""

    ctrl_dt: float = 0.02
    grace_period: float = 0.2
    left: str = "left_foot_touch"
    right: str = "right_foot_touch"
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return (C.REW_KBOT_SINGLE_FOOT_CONTACT, [_touch(ctx, self.left), _touch(ctx, self.right), ctx.cmd_offset(self.cmd)],
                [self.ctrl_dt, self.grace_period], [""], 1)


@attrs.define(frozen=True, kw_only=True)
class NoContactPenalty(rewards.Reward):
    """train.py:159-167."""

    left: str = "left_foot_touch"
    right: str = "right_foot_touch"
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_NO_CONTACT, [_touch(ctx, self.left), _touch(ctx, self.right), ctx.cmd_offset(self.cmd)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class FeetAirtimeReward(rewards.StatefulReward):
    """train.py:170-218.  Carry: airtime per foot and the previous contact flags (stored inverted so that the reference's
    initial carry -- airtime 0, contact True -- is all zeros)."""

    ctrl_dt: float = 0.02
    touchdown_penalty: float = 0.4
    left: str = "left_foot_touch"
    right: str = "right_foot_touch"
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return (C.REW_KBOT_FEET_AIRTIME, [_touch(ctx, self.left), _touch(ctx, self.right), ctx.cmd_offset(self.cmd)],
                [self.ctrl_dt, self.touchdown_penalty], [""], 4)


@attrs.define(frozen=True, kw_only=True)
class ArmPositionReward(rewards.Reward):
    """train.py:221-270."""

    joint_indices: tuple[int, ...] = attrs.field()
    joint_biases: tuple[float, ...] = attrs.field()
    error_scale: float = attrs.field(default=0.1)
    cmd: str = "unified_command"

    ARM_JOINTS = (
        "dof_right_shoulder_pitch_03", "dof_right_shoulder_roll_03", "dof_right_shoulder_yaw_02", "dof_right_elbow_02",
        "dof_right_wrist_00", "dof_left_shoulder_pitch_03", "dof_left_shoulder_roll_03", "dof_left_shoulder_yaw_02",
        "dof_left_elbow_02", "dof_left_wrist_00",
    )

    @classmethod
    def create_reward(cls, physics_model: Model, scale: float = 0.05, error_scale: float = 0.1) -> "ArmPositionReward":
        qadr = physics_model.arrays["jnt_qposadr"]
        idx = tuple(int(qadr[physics_model.id("joint", n)]) - 7 for n in cls.ARM_JOINTS)
        return cls(joint_indices=idx, joint_biases=tuple(JOINT_BIASES[n] for n in cls.ARM_JOINTS), error_scale=error_scale, scale=scale)

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return (C.REW_KBOT_ARM_POSITION, [ctx.cmd_offset(self.cmd), *[i + 7 for i in self.joint_indices]],
                [self.error_scale, *self.joint_biases], [""], 0)


@attrs.define(frozen=True, kw_only=True)
class LinearVelocityTrackingReward(rewards.Reward):
    """train.py:273-299."""

    error_scale: float = attrs.field(default=0.25)
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_LINVEL_TRACKING, [ctx.cmd_offset(self.cmd)], [self.error_scale], [""], 0


@attrs.define(frozen=True, kw_only=True)
class AngularVelocityReward(rewards.Reward):
    """train.py:302-313 (the example's own class, not ksim.AngularVelocityReward)."""

    error_scale: float = attrs.field(default=0.25)
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_ANGVEL_TRACKING, [ctx.cmd_offset(self.cmd)], [self.error_scale], [""], 0


@attrs.define(frozen=True, kw_only=True)
class XYOrientationReward(rewards.Reward):
    """train.py:316-343."""

    error_scale: float = attrs.field(default=0.03)
    error_scale_zero_cmd: float = attrs.field(default=0.003)
    cmd: str = "unified_command"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_XY_ORIENTATION, [ctx.cmd_offset(self.cmd)], [self.error_scale, self.error_scale_zero_cmd], [""], 0


@attrs.define(frozen=True, kw_only=True)
class TerrainBaseHeightReward(rewards.Reward):
    """train.py:346-401."""

    base_idx: int = attrs.field()
    foot_left_idx: int = attrs.field()
    foot_right_idx: int = attrs.field()
    error_scale: float = attrs.field(default=0.25)
    standard_height: float = attrs.field(default=0.9)
    foot_origin_height: float = attrs.field(default=0.0)
    cmd: str = "unified_command"

    @classmethod
    def create(cls, *, physics_model: Model, base_body_name: str, foot_left_body_name: str, foot_right_body_name: str,
               scale: float, error_scale: float, standard_height: float, foot_origin_height: float) -> "TerrainBaseHeightReward":
        return cls(base_idx=physics_model.id("body", base_body_name), foot_left_idx=physics_model.id("body", foot_left_body_name),
                   foot_right_idx=physics_model.id("body", foot_right_body_name), scale=scale, error_scale=error_scale,
                   standard_height=standard_height, foot_origin_height=foot_origin_height)

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return (C.REW_KBOT_TERRAIN_BASE_HEIGHT, [ctx.cmd_offset(self.cmd), self.base_idx, self.foot_left_idx, self.foot_right_idx],
                [self.error_scale, self.standard_height, self.foot_origin_height], [""], 0)


@attrs.define(frozen=True, kw_only=True)
class FeetOrientationReward(rewards.Reward):
    """train.py:404-467 (``is_rotating`` is one flag per trajectory there, and here)."""

    error_scale: float = attrs.field(default=0.25)
    foot_left_idx: int = attrs.field(default=0)
    foot_right_idx: int = attrs.field(default=0)
    cmd: str = "unified_command"

    @classmethod
    def create(cls, *, physics_model: Model, foot_left_body_name: str, foot_right_body_name: str, scale: float,
               error_scale: float) -> "FeetOrientationReward":
        return cls(foot_left_idx=physics_model.id("body", foot_left_body_name),
                   foot_right_idx=physics_model.id("body", foot_right_body_name), scale=scale, error_scale=error_scale)

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return (C.REW_KBOT_FEET_ORIENTATION, [ctx.cmd_offset(self.cmd), self.foot_left_idx, self.foot_right_idx],
                [self.error_scale], [""], 0)


@attrs.define(frozen=True, kw_only=True)
class COMDistanceReward(rewards.Reward):
    """train.py:470-484."""

    error_scale: float = attrs.field(default=0.25)
    cmd: str = "unified_command"
    obs: str = "com_distance"

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_COM_DISTANCE, [ctx.cmd_offset(self.cmd), ctx.obs_offset(self.obs)], [self.error_scale], [""], 0


@attrs.define(frozen=True, kw_only=True)
class BaseAccelerationReward(rewards.Reward):
    """train.py:487-497."""

    error_scale: float = attrs.field(default=1.0)

    def encode(self, ctx: rewards.EncodeContext) -> rewards.Encoded:
        return C.REW_KBOT_BASE_ACCELERATION, [], [self.error_scale], [""], 0


RANDOMIZERS = {
    "static_friction": lambda m: randomization.StaticFrictionRandomizer(),
    "armature": lambda m: randomization.ArmatureRandomizer(),
    "joint_damping": lambda m: randomization.JointDampingRandomizer(scale_lower=0.5, scale_upper=2.5),
    "floor_friction": lambda m: randomization.FloorFrictionRandomizer.from_geom_name(model=m, floor_geom_name="floor", scale_lower=0.5,
                                                                                     scale_upper=1.5),
    "all_body_COM": lambda m: randomization.AllBodiesCOMRandomizer(scale=0.05),
    "all_body_inertia": lambda m: randomization.AllBodiesInertiaRandomizer(scale=0.15),
    "collision_body": lambda m: randomization.CollisionBodyRandomizer.from_geom_names(
        model=m, geom_names=[f"{side}FootBushing_GPF_1517_12_collision_capsule_{k}" for side in "LR" for k in (0, 1)],
        radius_scale=0.01, length_scale=0.03, position_jitter_x=0.020, position_jitter_y=0.005, position_jitter_z=0.005),
}


def make_spec(config: KbotConfig | None = None, core_only: bool = False) -> TaskSpec:
    config = KbotConfig() if config is None else config
    model, metadata = load_kbot(config)
    eng = EngineConfig(
        dt=config.dt, ctrl_dt=config.ctrl_dt, action_latency_range=config.action_latency_range,
        drop_action_prob=config.drop_action_prob, zero_offset_mag=config.zero_offset_mag, action_clip_max=config.action_clip_max,
        solver=config.solver,
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
        "feet_position": (observation.BodyPositionObservation.create(physics_model=model, body_names=(left, right)) if core_only
                          else FeetPositionObservation.create(physics_model=model, base_body_name="base", foot_left_body_name=left,
                                                              foot_right_body_name=right)),
        **({} if core_only else {"base_height": BaseHeightObservation()}),
        "imu_projected_gravity": observation.ProjectedGravityObservation.create(
            physics_model=model, framequat_name="imu_site_quat", noise=AdditiveGaussianNoise(std=math.radians(3)),
            min_lag=0.0, max_lag=0.75, imu_bias=math.radians(4)),
        "projected_gravity": observation.ProjectedGravityObservation.create(physics_model=model, framequat_name="imu_site_quat"),
        **({} if core_only else {"com_distance": COMDistanceObservation()}),
    }
    if not core_only:
        return _full_task(config, model, metadata, eng, obs, left, right)
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
    return _spec(model, metadata, eng, obs, cmds, rews, terms)


def _spec(model: Model, metadata: Metadata, eng: EngineConfig, obs: dict, cmds: dict, rews: dict, terms: dict) -> TaskSpec:
    return TaskSpec(
        model=model, config=eng,
        actuators=PositionActuators(physics_model=model, metadata=metadata, action_noise=AdditiveGaussianNoise(std=0.02),
                                    torque_noise=AdditiveGaussianNoise(std=0.0), kp_scale=1.4, kd_scale=1.4, torque_limit_scale=1.5),
        resets=[
            resets.RandomJointPositionReset.create(model, JOINT_BIASES, scale=0.1),
            resets.RandomJointVelocityReset(scale=2.0),
            resets.RandomBaseVelocityXYReset(scale=0.2),
            resets.RandomHeadingReset(),
            PlaneXYPositionReset(x_range=0.1, y_range=0.1),
        ],
        events={
            "force_push": events.ForcePushEvent.from_body_name(model=model, body_name="base", max_force=200.0, max_torque=10.0,
                                                               duration_range=(0.1, 0.5), interval_range=(0.0, 6.0)),
        },
        observations=obs, commands=cmds, rewards=rews, terminations=terms,
        randomized_fields=("dof_frictionloss", "dof_armature", "dof_damping", "body_ipos", "body_mass", "body_inertia", "geom_size",
                           "geom_pos"),
    )


def _full_task(config: KbotConfig, model: Model, metadata: Metadata, eng: EngineConfig, obs: dict, left: str, right: str) -> TaskSpec:
    """Commands, rewards and terminations of examples/kbot/train.py:1204-1266."""
    arm_names = list(JOINT_BIASES)[10:20]
    rng = model.arrays["jnt_range"]
    arm_limits = tuple(zip(*[tuple(float(v) for v in rng[model.id("joint", n)]) for n in arm_names], strict=False))
    cmds = {
        "unified_command": UnifiedCommand(vx_range=(-0.5, 1.2), vy_range=(-0.5, 0.5), wz_range=(-1.0, 1.0), bh_range=(-0.25, 0.05),
                                          rx_range=(-0.25, 0.25), ry_range=(-0.25, 0.25), arms_range=arm_limits,
                                          ctrl_dt=config.ctrl_dt, switch_prob=config.ctrl_dt / 5),
    }
    feet = dict(physics_model=model, foot_left_body_name=left, foot_right_body_name=right)
    rews = {
        "linvel": LinearVelocityTrackingReward(scale=0.2, error_scale=0.2),
        "angvel": AngularVelocityReward(scale=0.1, error_scale=0.2),
        "roll_pitch": XYOrientationReward(scale=0.2, error_scale=0.03, error_scale_zero_cmd=0.01),
        "base_height": TerrainBaseHeightReward.create(base_body_name="base", scale=0.2, error_scale=0.02, standard_height=0.75,
                                                      foot_origin_height=0.06, **feet),
        "arm_pos": ArmPositionReward.create_reward(model, scale=0.2, error_scale=0.1),
        "single_contact": SingleFootContactReward(scale=0.1, ctrl_dt=config.ctrl_dt, grace_period=2.0),
        "no_contact_p": NoContactPenalty(scale=0.1),
        "feet_airtime": FeetAirtimeReward(scale=1.5, ctrl_dt=config.ctrl_dt, touchdown_penalty=0.4),
        "feet_orient": FeetOrientationReward.create(scale=0.1, error_scale=0.02, **feet),
        "com_distance": COMDistanceReward(scale=0.05, error_scale=0.04),
        "base_accel": BaseAccelerationReward(scale=0.1, error_scale=5.0),
    }
    terms = {
        "bad_z": TerrainBadZTermination.create(base_body_name="base", unhealthy_z=0.35, **feet),
        "not_upright": terminations.NotUprightTermination(max_radians=math.radians(45)),
        "episode_length": terminations.EpisodeLengthTermination(max_length_sec=24),
    }
    return _spec(model, metadata, eng, obs, cmds, rews, terms)
