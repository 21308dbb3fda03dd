"""Reward plugins.  Mirrors ksim/rewards.py:96-1240 for the rewards the B200 reward kernel implements.

Class names, constructor fields and the Reward/Penalty sign rule (rewards.py:118-124) follow the reference;
evaluation happens in ``ksim_b200/csrc/b2s_rewards.cuh`` over the (T, N) trajectory records.
``encode(ctx)`` -> (REW_* code, ints, floats, component names, carry floats); ``ctx`` resolves observation /
command names to record offsets.
"""

from __future__ import annotations

__all__ = ["AMPReward", "DISCRIMINATOR_OUTPUT_KEY", 
    "Reward", "StatefulReward", "StayAliveReward", "UprightReward", "LinearVelocityReward", "AngularVelocityReward",
    "FeetAirTimeReward", "SparseTargetHeightReward", "ForcePenalty", "MotionlessAtRestPenalty",
    "ActionVelocityPenalty", "ActionAccelerationPenalty", "ActionJerkPenalty", "BaseHeightReward",
    "BaseHeightRangeReward", "OffAxisVelocityReward", "SmallJointVelocityReward", "SmallJointAccelerationReward",
    "SmallJointJerkReward", "AvoidLimitsPenalty", "TorquePenalty", "EnergyPenalty", "JointDeviationPenalty", "FlatBodyReward",
    "NoRollReward", "LinkAccelerationPenalty", "LinkJerkPenalty", "FeetSlipPenalty", "ReachabilityPenalty", "SymmetryReward",
    "PairwiseSymmetryReward", "PositionTrackingReward", "SinusoidalGaitReward", "JointPositionReward", "IntersectionPenalty",
    "exp_kernel", "exp_kernel_with_penalty",
]

import math
from abc import ABC, abstractmethod
from typing import Protocol

import attrs

from ksim_b200 import codes as C
from ksim_b200.scales import Scale, convert_to_scale

Encoded = tuple[int, list[int], list[float], list[str], int]


def exp_kernel(x: float, scale: float) -> float:
    return math.exp(-(x * x) / (2 * scale**2))


def exp_kernel_with_penalty(x: float, scale: float, sq_scale: float, abs_scale: float) -> float:
    return abs(x) * -abs_scale + x * x * -sq_scale + math.exp(-(x * x) / (2 * scale**2))


class EncodeContext(Protocol):
    model: object  # ksim_b200.mjcf.Model

    def obs_offset(self, name: str) -> int: ...
    def obs_shape(self, name: str) -> tuple[int, ...]: ...
    def cmd_offset(self, name: str, field: str | None = None) -> int: ...
    def cmd_type(self, name: str) -> type: ...
    def aux_offset(self, name: str) -> int: ...


def _norm_flag(norm: str) -> int:
    if norm not in ("l1", "l2"):
        raise ValueError(f"Invalid norm: {norm}")
    return int(norm == "l2")


@attrs.define(frozen=True, kw_only=True)
class Reward(ABC):
    scale: Scale = attrs.field(converter=convert_to_scale)

    @abstractmethod
    def encode(self, ctx: EncodeContext) -> Encoded: ...

    @property
    def is_penalty(self) -> bool:
        reward_name = self.__class__.__name__
        if reward_name.lower().endswith("reward"):
            return False
        if reward_name.lower().endswith("penalty"):
            return True
        raise ValueError(f"Reward function {reward_name} does not end with 'Reward' or 'Penalty'")


@attrs.define(frozen=True, kw_only=True)
class StatefulReward(Reward):
    pass


@attrs.define(frozen=True, kw_only=True)
class StayAliveReward(Reward):
    balance: float = attrs.field(default=100.0)
    success_reward: float | None = attrs.field(default=None)

    def encode(self, ctx: EncodeContext) -> Encoded:
        has = self.success_reward is not None
        return C.REW_STAY_ALIVE, [int(has)], [float(self.balance), float(self.success_reward or 0.0)], [""], 0


DISCRIMINATOR_OUTPUT_KEY = "_discriminator_logits"  # ksim/task/amp.py


@attrs.define(frozen=True, kw_only=True)
class AMPReward(Reward):
    """``ksim.task.amp.AMPReward`` (amp.py:100-112): ``max(0, 1 - 0.25 (logit - 1)^2)`` of the discriminator logits the task's
    ``postprocess_trajectory`` hook wrote into ``trajectory.aux_outputs``; the task must declare that aux output
    (``TaskSpec.aux_outputs``), as the reference raises when it is missing."""

    aux_output: str = attrs.field(default=DISCRIMINATOR_OUTPUT_KEY)

    def encode(self, ctx: EncodeContext) -> Encoded:
        try:
            off = ctx.aux_offset(self.aux_output)
        except KeyError as e:
            raise ValueError("AMPReward auxiliary output is missing! Make sure you are using it within the context of an AMP task, "
                             "which populates the auxiliary output for you.") from e
        return C.REW_AMP, [off], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class LinearVelocityReward(Reward):
    cmd: str = attrs.field()
    vel_length_scale: float = attrs.field(default=0.25)
    yaw_length_scale: float = attrs.field(default=0.25)
    zero_threshold: float = attrs.field(default=0.01)
    vis_height: float = attrs.field(default=0.5)
    sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        floats = [self.vel_length_scale, self.yaw_length_scale, self.zero_threshold, self.sq_scale, self.abs_scale]
        return C.REW_LINEAR_VELOCITY, [ctx.cmd_offset(self.cmd)], [float(f) for f in floats], ["vel", "yaw", "x", "y"], 0


@attrs.define(frozen=True, kw_only=True)
class AngularVelocityReward(Reward):
    cmd: str = attrs.field()
    angvel_length_scale: float = attrs.field(default=0.25)
    sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    vis_height: float = attrs.field(default=0.5)
    zero_threshold: float = attrs.field(default=math.radians(5.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        floats = [self.angvel_length_scale, self.sq_scale, self.abs_scale, self.zero_threshold]
        return C.REW_ANGULAR_VELOCITY, [ctx.cmd_offset(self.cmd)], [float(f) for f in floats], ["mag", "dir"], 0


@attrs.define(frozen=True, kw_only=True)
class OffAxisVelocityReward(Reward):
    lin_length_scale: float = attrs.field(default=0.25)
    ang_length_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_OFF_AXIS_VELOCITY, [], [float(self.lin_length_scale), float(self.ang_length_scale)], ["linz", "angx", "angy"], 0


@attrs.define(frozen=True, kw_only=True)
class BaseHeightReward(Reward):
    height_target: float = attrs.field()
    norm: str = attrs.field(default="l1")
    kernel_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        _norm_flag(self.norm)
        return C.REW_BASE_HEIGHT, [], [float(self.height_target), float(self.kernel_scale)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class BaseHeightRangeReward(Reward):
    z_lower: float = attrs.field()
    z_upper: float = attrs.field()
    dropoff: float = attrs.field()

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_BASE_HEIGHT_RANGE, [], [float(self.z_lower), float(self.z_upper), float(self.dropoff)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class ActionVelocityPenalty(Reward):
    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_ACTION_VELOCITY, [_norm_flag(self.norm)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class ActionAccelerationPenalty(Reward):
    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_ACTION_ACCELERATION, [_norm_flag(self.norm)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class ActionJerkPenalty(Reward):
    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_ACTION_JERK, [_norm_flag(self.norm)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class SmallJointVelocityReward(Reward):
    kernel_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_SMALL_JOINT_VELOCITY, [], [float(self.kernel_scale)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class SmallJointAccelerationReward(Reward):
    """ksim/rewards.py:479-491."""

    kernel_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_SMALL_JOINT_ACCELERATION, [], [float(self.kernel_scale)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class SmallJointJerkReward(Reward):
    """ksim/rewards.py:495-508."""

    kernel_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_SMALL_JOINT_JERK, [], [float(self.kernel_scale)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class AvoidLimitsPenalty(Reward):
    """ksim/rewards.py:522-554; ``joint_limits`` is ``((lo, hi), ...)`` for the joints after the free joint (+-inf: unlimited)."""

    joint_limits: tuple[tuple[float, float], ...] = attrs.field()

    @joint_limits.validator
    def _check(self, _attr: object, value: tuple[tuple[float, float], ...]) -> None:
        if any(len(r) != 2 or r[0] > r[1] for r in value):
            raise ValueError(f"Joint range must be sorted (n_joints, 2), got {value}")

    def encode(self, ctx: EncodeContext) -> Encoded:
        n = len(self.joint_limits)
        if n != ctx.model.nq - 7:
            raise ValueError(f"AvoidLimitsPenalty has {n} joint ranges, the model has {ctx.model.nq - 7} joints after the free joint")
        return (C.REW_AVOID_LIMITS, [n], [float(r[0]) for r in self.joint_limits] + [float(r[1]) for r in self.joint_limits], [""], 0)

    @classmethod
    def create(cls, model: object, factor: float = 0.05, scale: float | Scale = 1.0) -> "AvoidLimitsPenalty":
        lims = []
        for j in range(1, model.njnt):
            lo, hi = (float(v) for v in model.jnt_range[j])
            diff = (hi - lo) * factor
            lims.append((lo - diff, hi + diff) if model.jnt_limited[j] else (-math.inf, math.inf))
        return cls(joint_limits=tuple(lims), scale=scale)


@attrs.define(frozen=True, kw_only=True)
class TorquePenalty(Reward):
    """ksim/rewards.py:557-582."""

    ctrl_scales: tuple[float, ...] = attrs.field()
    _code = C.REW_TORQUE

    def encode(self, ctx: EncodeContext) -> Encoded:
        if len(self.ctrl_scales) != ctx.model.nu:
            raise ValueError(f"{len(self.ctrl_scales)} ctrl scales for {ctx.model.nu} actuators")
        return self._code, [len(self.ctrl_scales)], [float(v) for v in self.ctrl_scales], [""], 0

    @classmethod
    def create(cls, model: object, scale: float | Scale = 1.0) -> "TorquePenalty":
        rng = model.actuator_ctrlrange
        return cls(ctrl_scales=tuple(float((hi - lo) / 2.0) for lo, hi in rng), scale=scale)


@attrs.define(frozen=True, kw_only=True)
class EnergyPenalty(TorquePenalty):
    """ksim/rewards.py:586-593: |ctrl / scale * qvel[6:]| (one actuator per joint dof, in dof order)."""

    _code = C.REW_ENERGY

    def encode(self, ctx: EncodeContext) -> Encoded:
        if ctx.model.nu != ctx.model.nv - 6:
            raise ValueError("EnergyPenalty multiplies ctrl and qvel[6:] element-wise: needs one actuator per joint dof")
        return super().encode(ctx)


@attrs.define(frozen=True, kw_only=True)
class JointDeviationPenalty(Reward):
    """ksim/rewards.py:596-662; ``joint_indices`` count from the first joint after the free joint."""

    norm: str = attrs.field(default="l2")
    joint_indices: tuple[int, ...] = attrs.field()
    joint_targets: tuple[float, ...] = attrs.field()
    joint_weights: tuple[float, ...] | None = attrs.field(default=None)

    def encode(self, ctx: EncodeContext) -> Encoded:
        n = len(self.joint_indices)
        if len(self.joint_targets) != n or (self.joint_weights is not None and len(self.joint_weights) != n):
            raise ValueError("joint indices, targets and weights must have the same length")
        if any(i < 0 or i >= ctx.model.nq - 7 for i in self.joint_indices):
            raise ValueError(f"joint index outside [0, {ctx.model.nq - 7})")
        w = [1.0] * n if self.joint_weights is None else [float(v) for v in self.joint_weights]
        return (C.REW_JOINT_DEVIATION, [_norm_flag(self.norm), n, int(self.joint_weights is not None), *self.joint_indices],
                [float(v) for v in self.joint_targets] + w, [""], 0)

    @classmethod
    def create(cls, physics_model: object, joint_names: tuple[str, ...], joint_targets: tuple[float, ...],
               scale: float | Scale = 1.0, joint_weights: tuple[float, ...] | None = None) -> "JointDeviationPenalty":
        if len(joint_names) != len(joint_targets):
            raise ValueError(f"Joint names and targets must match, got {len(joint_names)} and {len(joint_targets)}")
        idx = tuple(int(physics_model.jnt_qposadr[physics_model.id("joint", n)]) - 7 for n in joint_names)
        return cls(joint_indices=idx, joint_targets=tuple(float(v) for v in joint_targets),
                   joint_weights=None if joint_weights is None else tuple(float(v) for v in joint_weights), scale=scale)


@attrs.define(frozen=True, kw_only=True)
class FlatBodyReward(Reward):
    """ksim/rewards.py:666-700 (``norm`` is accepted and, as in the reference, unused)."""

    body_indices: tuple[int, ...] = attrs.field()
    plane: tuple[float, float, float] = attrs.field(default=(0.0, 0.0, 1.0))
    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        if not self.body_indices or any(b < 0 or b >= ctx.model.nbody for b in self.body_indices):
            raise ValueError("FlatBodyReward needs body indices inside the model")
        return C.REW_FLAT_BODY, [len(self.body_indices), *self.body_indices], [float(v) for v in self.plane], [""], 0

    @classmethod
    def create(cls, physics_model: object, body_names: tuple[str, ...], plane: tuple[float, float, float] = (0.0, 0.0, 1.0),
               norm: str = "l2", scale: float | Scale = 1.0) -> "FlatBodyReward":
        return cls(body_indices=tuple(physics_model.id("body", n) for n in body_names), plane=plane, norm=norm, scale=scale)


@attrs.define(frozen=True, kw_only=True)
class NoRollReward(Reward):
    """ksim/rewards.py:781-805."""

    angvel_scale: float = attrs.field(default=0.25)
    roll_scale: float = attrs.field(default=0.25)
    angvel_sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    angvel_abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    roll_sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    roll_abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        floats = [self.angvel_scale, self.roll_scale, self.angvel_sq_scale, self.angvel_abs_scale, self.roll_sq_scale, self.roll_abs_scale]
        return C.REW_NO_ROLL, [], [float(f) for f in floats], ["angvel", "pose"], 0


@attrs.define(frozen=True, kw_only=True)
class LinkAccelerationPenalty(Reward):
    """ksim/rewards.py:809-822."""

    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_LINK_ACCELERATION, [_norm_flag(self.norm), int(ctx.model.nbody)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class LinkJerkPenalty(Reward):
    """ksim/rewards.py:826-841."""

    norm: str = attrs.field(default="l2")

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_LINK_JERK, [_norm_flag(self.norm), int(ctx.model.nbody)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class UprightReward(Reward):
    angvel_scale: float = attrs.field(default=0.25)
    linvel_scale: float = attrs.field(default=0.25)
    pose_scale: float = attrs.field(default=0.25)
    angvel_sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    angvel_abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    linvel_sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    linvel_abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    pose_sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    pose_abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        floats = [self.angvel_scale, self.linvel_scale, self.pose_scale, self.angvel_sq_scale, self.angvel_abs_scale,
                  self.linvel_sq_scale, self.linvel_abs_scale, self.pose_sq_scale, self.pose_abs_scale]
        return C.REW_UPRIGHT, [], [float(f) for f in floats], ["angvel", "linvel", "pose"], 0


def _contact_rows(ctx: EncodeContext, contact_obs: str, n: int) -> int:
    shape = ctx.obs_shape(contact_obs)
    if len(shape) != 2 or shape[-1] != n:
        raise ValueError(f"contact observation {contact_obs} has shape {shape}, expected (rows, {n})")
    return shape[0]


@attrs.define(frozen=True, kw_only=True)
class FeetAirTimeReward(StatefulReward):
    gait_period: Scale = attrs.field(converter=convert_to_scale)
    air_time_percent: float = attrs.field()
    ctrl_dt: float = attrs.field()
    contact_obs: str = attrs.field()
    num_feet: int = attrs.field(default=2)
    bias: float = attrs.field(default=0.0)
    alpha: float = attrs.field(default=0.1)
    ground_penalty: float = attrs.field(default=0.1)
    linvel_moving_threshold: float = attrs.field(default=0.1)
    angvel_moving_threshold: float = attrs.field(default=math.radians(10))
    linvel_cmd: str | None = attrs.field(default=None)
    angvel_cmd: str | None = attrs.field(default=None)

    def encode(self, ctx: EncodeContext) -> Encoded:
        rows = _contact_rows(ctx, self.contact_obs, self.num_feet)
        lcmd = -1 if self.linvel_cmd is None else ctx.cmd_offset(self.linvel_cmd, "vel")
        acmd = -1 if self.angvel_cmd is None else ctx.cmd_offset(self.angvel_cmd, "vel")
        gi, gf = self.gait_period.encode()
        ints = [ctx.obs_offset(self.contact_obs), self.num_feet, rows, lcmd, acmd, gi[0]]
        floats = [self.air_time_percent, self.ctrl_dt, self.bias, self.alpha, self.ground_penalty,
                  self.linvel_moving_threshold, self.angvel_moving_threshold, *gf]
        return C.REW_FEET_AIR_TIME, ints, [float(f) for f in floats], ["air", "gnd", "stationary_gnd"], 3 * self.num_feet


@attrs.define(frozen=True, kw_only=True)
class SparseTargetHeightReward(StatefulReward):
    contact_obs: str = attrs.field()
    position_obs: str = attrs.field()
    height: float = attrs.field()
    num_bodies: int = attrs.field(default=2)
    linvel_moving_threshold: float = attrs.field(default=0.5)
    angvel_moving_threshold: float = attrs.field(default=math.radians(30))

    def encode(self, ctx: EncodeContext) -> Encoded:
        rows = _contact_rows(ctx, self.contact_obs, self.num_bodies)
        if ctx.obs_shape(self.position_obs) != (self.num_bodies, 3):
            raise ValueError(f"position observation {self.position_obs} must have shape ({self.num_bodies}, 3)")
        ints = [ctx.obs_offset(self.contact_obs), ctx.obs_offset(self.position_obs), self.num_bodies, rows]
        floats = [self.height, self.linvel_moving_threshold, self.angvel_moving_threshold]
        return C.REW_SPARSE_TARGET_HEIGHT, ints, [float(f) for f in floats], [""], self.num_bodies


@attrs.define(frozen=True, kw_only=True)
class MotionlessAtRestPenalty(Reward):
    linvel_moving_threshold: float = attrs.field(default=0.5)
    angvel_moving_threshold: float = attrs.field(default=math.radians(30))

    def encode(self, ctx: EncodeContext) -> Encoded:
        return C.REW_MOTIONLESS_AT_REST, [], [float(self.linvel_moving_threshold), float(self.angvel_moving_threshold)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class ForcePenalty(Reward):
    force_obs: str = attrs.field()
    ctrl_dt: float = attrs.field()
    expected_weight: float = attrs.field()
    bias: float = attrs.field(default=0.0)

    def encode(self, ctx: EncodeContext) -> Encoded:
        shape = ctx.obs_shape(self.force_obs)
        if len(shape) != 2:
            raise ValueError(f"force observation {self.force_obs} must be (rows, dim), got {shape}")
        ints = [ctx.obs_offset(self.force_obs), shape[0], shape[1]]
        return C.REW_FORCE_PENALTY, ints, [float(self.ctrl_dt), float(self.expected_weight), float(self.bias)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class FeetSlipPenalty(Reward):
    """ksim/rewards.py:1089-1106: speed of the feet that are in contact; ``contact_obs`` is ``(rows, n_feet)``,
    ``velocity_obs`` ``(n_feet, 6)`` (e.g. ``BodyVelocityObservation`` of the foot bodies)."""

    contact_obs: str = attrs.field()
    velocity_obs: str = attrs.field()

    def encode(self, ctx: EncodeContext) -> Encoded:
        cs, vs = ctx.obs_shape(self.contact_obs), ctx.obs_shape(self.velocity_obs)
        if len(cs) != 2 or len(vs) != 2 or vs[1] != 6 or vs[0] != cs[1]:
            raise ValueError(f"FeetSlipPenalty needs contact (rows, n) and velocity (n, 6) observations, got {cs} and {vs}")
        return C.REW_FEET_SLIP, [ctx.obs_offset(self.contact_obs), cs[0], cs[1], ctx.obs_offset(self.velocity_obs)], [], [""], 0


@attrs.define(frozen=True, kw_only=True)
class ReachabilityPenalty(Reward):
    """ksim/rewards.py:970-989: |action - qpos[7:]| beyond the per-joint envelope ``delta_max_j``."""

    delta_max_j: tuple[float, ...] = attrs.field()
    squared: bool = attrs.field(default=True)

    def encode(self, ctx: EncodeContext) -> Encoded:
        n = len(self.delta_max_j)
        if n != ctx.model.nq - 7 or n != ctx.model.nu:
            raise ValueError("ReachabilityPenalty compares action and qpos[7:] element-wise: delta_max_j needs one entry per joint")
        return C.REW_REACHABILITY, [int(self.squared), n], [float(v) for v in self.delta_max_j], [""], 0


@attrs.define(frozen=True, kw_only=True)
class SymmetryReward(StatefulReward):
    """ksim/rewards.py:844-897 (carry: an EMA over rollouts of the mean joint offsets, initially the targets)."""

    joint_indices: tuple[int, ...] = attrs.field()
    joint_targets: tuple[float, ...] = attrs.field()
    alpha: float = attrs.field(default=0.1)

    def encode(self, ctx: EncodeContext) -> Encoded:
        n = len(self.joint_indices)
        if n == 0 or len(self.joint_targets) != n or any(i < 0 or i >= ctx.model.nq - 7 for i in self.joint_indices):
            raise ValueError("SymmetryReward needs matching joint indices / targets inside the model")
        return C.REW_SYMMETRY, [n, *self.joint_indices], [float(self.alpha)] + [float(v) for v in self.joint_targets], [""], n

    @classmethod
    def create(cls, physics_model: object, joint_names: tuple[str, ...], joint_targets: tuple[float, ...], alpha: float = 0.1,
               scale: float | Scale = 1.0) -> "SymmetryReward":
        if len(joint_names) != len(joint_targets):
            raise ValueError(f"Joint names and targets must match, got {len(joint_names)} and {len(joint_targets)}")
        idx = tuple(int(physics_model.jnt_qposadr[physics_model.id("joint", n)]) - 7 for n in joint_names)
        return cls(joint_indices=idx, joint_targets=tuple(float(v) for v in joint_targets), alpha=alpha, scale=scale)


@attrs.define(frozen=True, kw_only=True)
class PairwiseSymmetryReward(StatefulReward):
    """ksim/rewards.py:901-966."""

    left_joint_index: int = attrs.field()
    right_joint_index: int = attrs.field()
    left_zero: float = attrs.field(default=0.0)
    right_zero: float = attrs.field(default=0.0)
    flipped: bool = attrs.field(default=False)
    alpha: float = attrs.field(default=0.1)
    kernel_scale: float = attrs.field(default=0.25)
    sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        if any(i < 0 or i >= ctx.model.nq - 7 for i in (self.left_joint_index, self.right_joint_index)):
            raise ValueError("PairwiseSymmetryReward joint index outside the model")
        floats = [self.left_zero, self.right_zero, self.alpha, self.kernel_scale, self.sq_scale, self.abs_scale]
        return (C.REW_PAIRWISE_SYMMETRY, [self.left_joint_index, self.right_joint_index, int(self.flipped)],
                [float(f) for f in floats], ["lr", "self"], 2)

    @classmethod
    def create(cls, physics_model: object, left_joint_name: str, right_joint_name: str, **kw: object) -> "PairwiseSymmetryReward":
        adr = physics_model.jnt_qposadr
        return cls(left_joint_index=int(adr[physics_model.id("joint", left_joint_name)]) - 7,
                   right_joint_index=int(adr[physics_model.id("joint", right_joint_name)]) - 7, **kw)


@attrs.define(frozen=True, kw_only=True)
class PositionTrackingReward(Reward):
    """Closeness of a body (relative to a base body) to the current point of a ``CartesianCoordinateCommand``
    (ksim/rewards.py:704-744)."""

    tracked_body_idx: int = attrs.field()
    base_body_idx: int = attrs.field()
    command_name: str = attrs.field()
    body_name: str = attrs.field(default="")
    norm: str = attrs.field(default="l1")
    kernel_scale: float = attrs.field(default=0.25)

    def encode(self, ctx: EncodeContext) -> Encoded:
        ints = [int(self.tracked_body_idx), int(self.base_body_idx), ctx.cmd_offset(self.command_name), _norm_flag(self.norm)]
        return C.REW_POSITION_TRACKING, ints, [float(self.kernel_scale)], [""], 0

    @classmethod
    def create(cls, model, command_name: str, tracked_body_name: str, base_body_name: str, norm: str = "l1", kernel_scale: float = 0.25,  # noqa: ANN001, ANN206
               scale: float | Scale = 1.0):
        return cls(tracked_body_idx=model.id("body", tracked_body_name), base_body_idx=model.id("body", base_body_name), norm=norm,
                   command_name=command_name, body_name=tracked_body_name, scale=scale, kernel_scale=kernel_scale)


@attrs.define(frozen=True, kw_only=True)
class SinusoidalGaitReward(Reward):
    """``1 - sum_f |foot height - commanded height| / max_height`` (ksim/rewards.py:1320-1348)."""

    ctrl_dt: float = attrs.field()
    max_height: float = attrs.field()
    pos_obs: str = attrs.field()
    pos_cmd: str = attrs.field()
    num_feet: int = attrs.field(default=2)

    def encode(self, ctx: EncodeContext) -> Encoded:
        if ctx.obs_shape(self.pos_obs) != (self.num_feet, 3):
            raise ValueError(f"position observation {self.pos_obs} must have shape ({self.num_feet}, 3)")
        ints = [ctx.obs_offset(self.pos_obs), int(self.num_feet), ctx.cmd_offset(self.pos_cmd, "height")]
        return C.REW_SINUSOIDAL_GAIT, ints, [float(self.max_height)], [""], 0


@attrs.define(frozen=True, kw_only=True)
class JointPositionReward(Reward):
    """Tracking of a ``JointPositionCommand``'s interpolated position and constant velocity (ksim/rewards.py:1352-1414)."""

    command_name: str = attrs.field()
    indices: tuple[int, ...] = attrs.field()
    pos_scale: float = attrs.field(default=0.25)
    vel_scale: float = attrs.field(default=0.25)
    sq_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))
    abs_scale: float = attrs.field(default=0.1, validator=attrs.validators.gt(0.0))

    def encode(self, ctx: EncodeContext) -> Encoded:
        ints = [len(self.indices), ctx.cmd_offset(self.command_name), *[int(i) for i in self.indices]]
        floats = [self.pos_scale, self.vel_scale, self.sq_scale, self.abs_scale]
        return C.REW_JOINT_POSITION, ints, [float(f) for f in floats], ["pos", "vel"], 0

    @classmethod
    def create(cls, physics_model, joint_names, command_name: str, pos_scale: float = 0.25, vel_scale: float = 0.25,  # noqa: ANN001, ANN206
               sq_scale: float = 0.1, abs_scale: float = 0.1, scale: float | Scale = 1.0):
        all_names = physics_model.names["joint"][1:]
        for name in joint_names:
            if name not in all_names:
                raise ValueError(f"Joint {name} not found in the model! Options are: {all_names}")
        return cls(indices=tuple(all_names.index(n) for n in joint_names), command_name=command_name, pos_scale=pos_scale,
                   vel_scale=vel_scale, sq_scale=sq_scale, abs_scale=abs_scale, scale=scale)


@attrs.define(frozen=True, kw_only=True)
class IntersectionPenalty(Reward):
    """1 when any two of the observed points are closer than ``min_distance`` (ksim/rewards.py:1418-1430)."""

    position_obs: str = attrs.field()
    min_distance: float = attrs.field()

    def encode(self, ctx: EncodeContext) -> Encoded:
        shape = ctx.obs_shape(self.position_obs)
        if len(shape) != 2 or shape[1] != 3:
            raise ValueError(f"position observation {self.position_obs} must have shape (n, 3)")
        return C.REW_INTERSECTION, [ctx.obs_offset(self.position_obs), int(shape[0])], [float(self.min_distance)], [""], 0
