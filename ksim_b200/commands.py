"""Command generators.  Mirrors ksim/commands.py:42-508 for the commands the rollout kernel carries per env.

State layout inside the per-env command block (floats):
  LinearVelocityCommand   [vel, yaw, xvel, yvel, target_vel, target_yaw]   (LinearVelocityCommandValue)
  AngularVelocityCommand  [vel, target_vel]                                (AngularVelocityCommandValue)
  FloatVectorCommand      [x_0 .. x_{n-1}]
``encode`` -> (CMD_* code, ints, floats, state floats).
"""

from __future__ import annotations

__all__ = ["Command", "FloatVectorCommand", "LinearVelocityCommand", "AngularVelocityCommand"]

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
