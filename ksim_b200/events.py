"""Per-physics-substep perturbation events.  Mirrors ksim/events.py:25-365 (names, fields)."""

from __future__ import annotations

__all__ = ["Event", "LinearPushEvent", "AngularPushEvent", "JumpEvent", "ForcePushEvent"]

from abc import ABC, abstractmethod

import attrs

from ksim_b200 import codes as C
from ksim_b200.mjcf import Model
from ksim_b200.scales import ConstantScale, Scale, convert_to_scale


@attrs.define(frozen=True, kw_only=True)
class Event(ABC):
    @abstractmethod
    def encode(self, model: Model) -> tuple[int, list[int], list[float]]: ...

    def state_size(self) -> int:
        return 1


@attrs.define(frozen=True, kw_only=True)
class AngularPushEvent(Event):
    angvel: float = attrs.field()
    vel_range: tuple[float, float] = attrs.field(default=(0.0, 1.0))
    interval_range: tuple[float, float] = attrs.field()
    scale: Scale = attrs.field(default=ConstantScale(scale=1.0), converter=convert_to_scale)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        si, sf = self.scale.encode()
        return C.EVENT_ANGULAR_PUSH, si, [self.angvel, *self.vel_range, *self.interval_range, *sf]


@attrs.define(frozen=True, kw_only=True)
class LinearPushEvent(Event):
    linvel: float = attrs.field()
    vel_range: tuple[float, float] = attrs.field(default=(0.0, 1.0))
    interval_range: tuple[float, float] = attrs.field()
    scale: Scale = attrs.field(default=ConstantScale(scale=1.0), converter=convert_to_scale)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        si, sf = self.scale.encode()
        return C.EVENT_LINEAR_PUSH, si, [self.linvel, *self.vel_range, *self.interval_range, *sf]


@attrs.define(frozen=True, kw_only=True)
class JumpEvent(Event):
    jump_height_range: tuple[float, float] = attrs.field()
    interval_range: tuple[float, float] = attrs.field()
    scale: Scale = attrs.field(default=ConstantScale(scale=1.0), converter=convert_to_scale)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        si, sf = self.scale.encode()
        return C.EVENT_JUMP, si, [*self.jump_height_range, *self.interval_range, *sf]


@attrs.define(frozen=True, kw_only=True)
class ForcePushEvent(Event):
    max_force: float = attrs.field()
    max_torque: float = attrs.field()
    force_range: tuple[float, float] = attrs.field(default=(0.0, 1.0))
    duration_range: tuple[float, float] = attrs.field()
    interval_range: tuple[float, float] = attrs.field()
    body_id: int = attrs.field()
    scale: Scale = attrs.field(default=ConstantScale(scale=1.0), converter=convert_to_scale)

    @classmethod
    def from_body_name(cls, model: Model, body_name: str, *, max_force: float, max_torque: float,
                       duration_range: tuple[float, float], interval_range: tuple[float, float],
                       force_range: tuple[float, float] = (0.0, 1.0), scale: float | int | Scale = 1.0) -> "ForcePushEvent":
        if body_name not in model.names["body"]:
            raise ValueError(f"Body name {body_name} not found in model")
        return cls(max_force=max_force, max_torque=max_torque, duration_range=duration_range,
                   interval_range=interval_range, force_range=force_range, body_id=model.id("body", body_name), scale=scale)

    def state_size(self) -> int:
        return 8

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        si, sf = self.scale.encode()
        return (C.EVENT_FORCE_PUSH, si + [self.body_id],
                [self.max_force, self.max_torque, *self.force_range, *self.duration_range, *self.interval_range, *sf])
