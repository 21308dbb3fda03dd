"""Command generators.  Mirrors ksim/commands.py:42-508 for the commands the rollout kernel carries per env.

State layout inside the per-env command block (floats):
  LinearVelocityCommand   [vel, yaw, xvel, yvel, target_vel, target_yaw]   (LinearVelocityCommandValue)
  AngularVelocityCommand  [vel, target_vel]                                (AngularVelocityCommandValue)
  FloatVectorCommand      [x_0 .. x_{n-1}]
  IntVectorCommand        [x_0 .. x_{n-1}]  (integers stored as floats, like the reference's jnp.where(mask, 0.0, ints))
  StartPositionCommand    [x, y, z];  StartQuaternionCommand [w, x, y, z];  BaseHeightCommand [height]
  CartesianCoordinateCommand  [current xyz, target xyz, speed]
  SinusoidalGaitCommand   [moving_flag, phase, height_0 .. height_{num_feet-1}]     (SinusoidalGaitCommandValue)
  JointPositionCommand    [start_0.., target_0.., total_time, num_steps, step_id]   (JointPositionCommandValue)
``encode`` -> (CMD_* code, ints, floats, state floats).
"""

from __future__ import annotations

__all__ = ["Command", "FloatVectorCommand", "IntVectorCommand", "StartPositionCommand", "StartQuaternionCommand", "BaseHeightCommand",
           "CartesianCoordinateCommand", "SinusoidalGaitCommand", "JointPositionCommand", "LinearVelocityCommand", "AngularVelocityCommand"]

from abc import ABC, abstractmethod

import attrs

from ksim_b200 import codes as C
from ksim_b200.scales import ConstantScale, Scale, convert_to_scale


@attrs.define(frozen=True, kw_only=True)
class Command(ABC):
    @abstractmethod
    def encode(self) -> tuple[int, list[int], list[float], int]: ...

    def fields(self) -> dict[str, int]:
        """Named sub-values -> offset inside this command's block (what rewards address)."""
        return {}


@attrs.define(frozen=True, kw_only=True)
class FloatVectorCommand(Command):
    ranges: tuple[tuple[float, float], ...] = attrs.field()
    switch_prob: float = attrs.field(default=0.0)
    zero_prob: float = attrs.field(default=0.0)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        floats: list[float] = []
        for lo, hi in self.ranges:
            floats += [float(lo), float(hi)]
        floats += [float(self.switch_prob), float(self.zero_prob)]
        if len(self.ranges) > 16:
            raise ValueError("FloatVectorCommand supports at most 16 dimensions")
        return C.CMD_FLOAT_VECTOR, [], floats, len(self.ranges)


@attrs.define(frozen=True, kw_only=True)
class IntVectorCommand(Command):
    """Integer vector drawn uniformly in inclusive ranges, zeroed with ``zero_prob`` (ksim/commands.py:137-169)."""

    ranges: tuple[tuple[int, int], ...] = attrs.field()
    switch_prob: float = attrs.field(default=0.0)
    zero_prob: float = attrs.field(default=0.0)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        if len(self.ranges) > 16:
            raise ValueError("IntVectorCommand supports at most 16 dimensions")
        ints: list[int] = []
        for lo, hi in self.ranges:
            if int(hi) < int(lo):
                raise ValueError("IntVectorCommand ranges must satisfy lo <= hi")
            ints += [int(lo), int(hi)]
        return C.CMD_INT_VECTOR, ints, [float(self.switch_prob), float(self.zero_prob)], len(self.ranges)


@attrs.define(frozen=True, kw_only=True)
class StartPositionCommand(Command):
    """The base position right after the reset, held for the episode (ksim/commands.py:173-192)."""

    def encode(self) -> tuple[int, list[int], list[float], int]:
        return C.CMD_START_POSITION, [], [], 3


@attrs.define(frozen=True, kw_only=True)
class StartQuaternionCommand(Command):
    """The base quaternion right after the reset, held for the episode (ksim/commands.py:195-214)."""

    def encode(self) -> tuple[int, list[int], list[float], int]:
        return C.CMD_START_QUATERNION, [], [], 4


@attrs.define(frozen=True, kw_only=True)
class BaseHeightCommand(Command):
    """Target base height, redrawn with ``switch_prob`` per step (ksim/commands.py:773-796)."""

    min_height: float = attrs.field()
    max_height: float = attrs.field()
    switch_prob: float = attrs.field(default=0.005)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        return C.CMD_BASE_HEIGHT, [], [float(self.min_height), float(self.max_height), float(self.switch_prob)], 1


@attrs.define(frozen=True, kw_only=True)
class CartesianCoordinateCommand(Command):
    """A point moving at a random speed towards random targets inside a (curriculum-scaled) box (ksim/commands.py:537-692)."""

    box_min: tuple[float, float, float] = attrs.field()
    box_max: tuple[float, float, float] = attrs.field()
    dt: float = attrs.field()
    base_name: str | None = attrs.field(default=None)
    vis_radius: float = attrs.field(default=0.05)
    vis_color: tuple[float, float, float, float] = attrs.field(default=(1.0, 0.0, 0.0, 0.8))
    min_speed: float = attrs.field(default=0.5)
    max_speed: float = attrs.field(default=3.0)
    switch_prob: float = attrs.field(default=0.0)
    jump_prob: float = attrs.field(default=0.0)
    unique_name: str | None = attrs.field(default=None)
    curriculum_scale: float = attrs.field(default=0.1)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        floats = [*self.box_min, *self.box_max, self.dt, self.min_speed, self.max_speed, self.switch_prob, self.jump_prob, self.curriculum_scale]
        return C.CMD_CARTESIAN_COORDINATE, [], [float(f) for f in floats], 7

    def fields(self) -> dict[str, int]:
        return {"current": 0, "target": 3, "speed": 6}

    @classmethod
    def create(cls, model, box_min, box_max, vis_target_name=None, vis_radius=0.05, vis_color=(1.0, 0.0, 0.0, 0.8), unique_name=None,  # noqa: ANN001, ANN206
               min_speed=0.5, max_speed=3.0, switch_prob=1.0, jump_prob=0.0, curriculum_scale=1.0):
        return cls(base_name=vis_target_name, box_min=tuple(box_min), box_max=tuple(box_max), dt=float(model.opt.timestep),
                   vis_radius=vis_radius, vis_color=vis_color, unique_name=unique_name, min_speed=min_speed, max_speed=max_speed,
                   switch_prob=switch_prob, jump_prob=jump_prob, curriculum_scale=curriculum_scale)


@attrs.define(frozen=True, kw_only=True)
class SinusoidalGaitCommand(Command):
    """Per-foot target heights of a phase-shifted half-sine swing (ksim/commands.py:703-770)."""

    gait_period: float = attrs.field()
    ctrl_dt: float = attrs.field()
    max_height: float = attrs.field()
    height_offset: float = attrs.field(default=0.0)
    num_feet: int = attrs.field(default=2)
    stance_ratio: float = attrs.field(default=0.6)
    moving_threshold: float = attrs.field(default=0.05)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        if not 1 <= self.num_feet <= 8:
            raise ValueError("SinusoidalGaitCommand supports 1..8 feet")
        floats = [self.gait_period, self.ctrl_dt, self.max_height, self.height_offset, self.stance_ratio, self.moving_threshold]
        return C.CMD_SINUSOIDAL_GAIT, [int(self.num_feet)], [float(f) for f in floats], 2 + int(self.num_feet)

    def fields(self) -> dict[str, int]:
        return {"moving_flag": 0, "phase": 1, "height": 2}


@attrs.define(frozen=True, kw_only=True)
class JointPositionCommand(Command):
    """Linear joint-space moves to random targets in random times (ksim/commands.py:799-900)."""

    indices: tuple[int, ...] = attrs.field()
    ranges: tuple[tuple[float, float], ...] = attrs.field()
    ctrl_dt: float = attrs.field()
    min_time: float = attrs.field(validator=attrs.validators.gt(0.0))
    max_time: float = attrs.field(validator=attrs.validators.gt(0.0))

    def encode(self) -> tuple[int, list[int], list[float], int]:
        n = len(self.indices)
        if n != len(self.ranges) or not 1 <= n <= 10:
            raise ValueError("JointPositionCommand needs one range per index, 1..10 joints")
        floats: list[float] = []
        for lo, hi in self.ranges:
            floats += [float(lo), float(hi)]
        floats += [float(self.ctrl_dt), float(self.min_time), float(self.max_time)]
        return C.CMD_JOINT_POSITION, [n, *[int(i) for i in self.indices]], floats, 2 * n + 3

    def fields(self) -> dict[str, int]:
        n = len(self.indices)
        return {"start_position": 0, "target_position": n, "total_time": 2 * n, "num_steps": 2 * n + 1, "step_id": 2 * n + 2}

    @classmethod
    def create(cls, physics_model, ctrl_dt: float, joint_names, min_time: float = 0.1, max_time: float = 1.0):  # noqa: ANN001, ANN206
        all_names = physics_model.names["joint"][1:]  # without the floating base
        for name in joint_names:
            if name not in all_names:
                raise ValueError(f"Joint {name} not found in the model! Options are: {all_names}")
        indices = tuple(all_names.index(name) for name in joint_names)
        rng = physics_model.jnt_range[1:]
        ranges = tuple((float(rng[i][0]), float(rng[i][1])) for i in indices)
        return cls(indices=indices, ranges=ranges, ctrl_dt=ctrl_dt, min_time=min_time, max_time=max_time)


def _scales(*scales: Scale) -> tuple[list[int], list[float]]:
    ints: list[int] = []
    floats: list[float] = []
    for s in scales:
        i, f = s.encode()
        ints += i
        floats += f
    return ints, floats


@attrs.define(frozen=True, kw_only=True)
class LinearVelocityCommand(Command):
    min_vel: Scale = attrs.field(converter=convert_to_scale)
    max_vel: Scale = attrs.field(converter=convert_to_scale)
    ctrl_dt: float = attrs.field()
    linear_accel: float = attrs.field()
    angular_accel: float = attrs.field()
    max_yaw: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    zero_prob: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    backward_prob: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    switch_prob: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    vis_height: float = attrs.field(default=0.5)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        ints, floats = _scales(self.min_vel, self.max_vel, self.max_yaw, self.zero_prob, self.backward_prob, self.switch_prob)
        floats += [float(self.ctrl_dt), float(self.linear_accel), float(self.angular_accel)]
        return C.CMD_LINEAR_VELOCITY, ints, floats, 6

    def fields(self) -> dict[str, int]:
        return {"vel": 0, "yaw": 1, "xvel": 2, "yvel": 3, "target_vel": 4, "target_yaw": 5}


@attrs.define(frozen=True, kw_only=True)
class AngularVelocityCommand(Command):
    max_vel: Scale = attrs.field(converter=convert_to_scale)
    ctrl_dt: float = attrs.field()
    angular_accel: float = attrs.field()
    min_vel: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    zero_prob: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    switch_prob: Scale = attrs.field(default=ConstantScale(scale=0.0), converter=convert_to_scale)
    vis_height: float = attrs.field(default=0.5)

    def encode(self) -> tuple[int, list[int], list[float], int]:
        ints, floats = _scales(self.min_vel, self.max_vel, self.zero_prob, self.switch_prob)
        floats += [float(self.ctrl_dt), float(self.angular_accel)]
        return C.CMD_ANGULAR_VELOCITY, ints, floats, 2

    def fields(self) -> dict[str, int]:
        return {"vel": 0, "target_vel": 1}
