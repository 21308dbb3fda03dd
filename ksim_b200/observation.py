"""Observation plugins.  Mirrors ksim/observation.py:89-747 (class names, constructor arguments, ``create``).

An observation here is a descriptor: the gather itself (and the noise / bias of ``Observation.add_noise``,
observation.py:110-142) runs inside the fused CUDA step kernel.  ``encode`` returns the task-program entry:
``(OBS_* code, ints, floats, output dim, carry floats)``.
"""

from __future__ import annotations

__all__ = [
    "Observation", "StatefulObservation", "BasePositionObservation", "BaseOrientationObservation",
    "BaseLinearVelocityObservation", "BaseAngularVelocityObservation", "JointPositionObservation",
    "JointVelocityObservation", "CenterOfMassInertiaObservation", "CenterOfMassVelocityObservation",
    "ActuatorForceObservation", "SensorObservation", "BaseLinearAccelerationObservation",
    "BaseAngularAccelerationObservation", "ActuatorAccelerationObservation", "ProjectedGravityObservation",
    "FeetContactObservation", "BodyPositionObservation", "BodyVelocityObservation", "FeetForceObservation",
    "FeetTorqueObservation", "TimestepObservation",
]

from abc import ABC, abstractmethod
from typing import Collection, Iterable

import attrs

from ksim_b200 import codes as C
from ksim_b200.mjcf import Model
from ksim_b200.noise import Bias, Noise

Encoded = tuple[int, list[int], list[float], int, int]


@attrs.define(frozen=True, kw_only=True)
class Observation(ABC):
    noise: Noise | None = attrs.field(default=None)
    bias: Bias | None = attrs.field(default=None)

    @abstractmethod
    def encode(self, model: Model) -> Encoded: ...

    def shape(self, model: Model) -> tuple[int, ...]:
        """Shape the reference's ``observe`` returns (the flat record stores it row-major)."""
        return (self.encode(model)[3],)

    @property
    def has_noisy_copy(self) -> bool:
        return self.noise is not None or self.bias is not None


@attrs.define(frozen=True, kw_only=True)
class StatefulObservation(Observation):
    pass


def _simple(code: int, dim_fn):  # noqa: ANN001, ANN202
    def encode(self, model: Model) -> Encoded:  # noqa: ANN001
        return code, [], [], dim_fn(model), 0

    return encode


@attrs.define(frozen=True, kw_only=True)
class BasePositionObservation(Observation):
    encode = _simple(C.OBS_BASE_POSITION, lambda m: 3)


@attrs.define(frozen=True, kw_only=True)
class BaseOrientationObservation(Observation):
    encode = _simple(C.OBS_BASE_ORIENTATION, lambda m: 4)


@attrs.define(frozen=True, kw_only=True)
class BaseLinearVelocityObservation(Observation):
    in_robot_frame: bool = attrs.field(default=False)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_BASE_LINEAR_VELOCITY, [int(self.in_robot_frame)], [], 3, 0


@attrs.define(frozen=True, kw_only=True)
class BaseAngularVelocityObservation(Observation):
    in_robot_frame: bool = attrs.field(default=False)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_BASE_ANGULAR_VELOCITY, [int(self.in_robot_frame)], [], 3, 0


@attrs.define(frozen=True, kw_only=True)
class JointPositionObservation(Observation):
    encode = _simple(C.OBS_JOINT_POSITION, lambda m: m.nq - 7)


@attrs.define(frozen=True, kw_only=True)
class JointVelocityObservation(Observation):
    encode = _simple(C.OBS_JOINT_VELOCITY, lambda m: m.nv - 6)


@attrs.define(frozen=True, kw_only=True)
class CenterOfMassInertiaObservation(Observation):
    encode = _simple(C.OBS_CENTER_OF_MASS_INERTIA, lambda m: 10 * (m.nbody - 1))


@attrs.define(frozen=True, kw_only=True)
class CenterOfMassVelocityObservation(Observation):
    encode = _simple(C.OBS_CENTER_OF_MASS_VELOCITY, lambda m: 6 * (m.nbody - 1))


@attrs.define(frozen=True, kw_only=True)
class ActuatorForceObservation(Observation):
    encode = _simple(C.OBS_ACTUATOR_FORCE, lambda m: m.nu)


@attrs.define(frozen=True, kw_only=True)
class BaseLinearAccelerationObservation(Observation):
    encode = _simple(C.OBS_BASE_LINEAR_ACCELERATION, lambda m: 3)


@attrs.define(frozen=True, kw_only=True)
class BaseAngularAccelerationObservation(Observation):
    encode = _simple(C.OBS_BASE_ANGULAR_ACCELERATION, lambda m: 3)


@attrs.define(frozen=True, kw_only=True)
class ActuatorAccelerationObservation(Observation):
    encode = _simple(C.OBS_ACTUATOR_ACCELERATION, lambda m: m.nv - 6)


@attrs.define(frozen=True, kw_only=True)
class TimestepObservation(Observation):
    encode = _simple(C.OBS_TIMESTEP, lambda m: 1)


def _sensor_range(model: Model, name: str) -> tuple[int, int]:
    if name not in model.names["sensor"]:
        options = "\n".join(sorted(model.names["sensor"]))
        raise ValueError(f"{name} not found in model. Available:\n{options}")
    return model.sensor_range(name)


@attrs.define(frozen=True, kw_only=True)
class SensorObservation(Observation):
    sensor_name: str = attrs.field()
    sensor_idx_range: tuple[int, int | None] = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, sensor_name: str, noise: Noise | None = None) -> "SensorObservation":
        return cls(noise=noise, sensor_name=sensor_name, sensor_idx_range=_sensor_range(physics_model, sensor_name))

    def encode(self, model: Model) -> Encoded:
        start, end = self.sensor_idx_range
        end = model.nsensordata if end is None else end
        return C.OBS_SENSOR, [int(start)], [], int(end - start), 0


@attrs.define(frozen=True, kw_only=True)
class ProjectedGravityObservation(StatefulObservation):
    """Lagged, biased projected-gravity reading from a framequat sensor (ksim/observation.py:363-442)."""

    framequat_idx_range: tuple[int, int | None] = attrs.field()
    gravity: tuple[float, float, float] = attrs.field()
    min_lag: float = attrs.field(validator=attrs.validators.and_(attrs.validators.ge(0.0), attrs.validators.lt(1.0)))
    max_lag: float = attrs.field(validator=attrs.validators.and_(attrs.validators.ge(0.0), attrs.validators.lt(1.0)))
    imu_bias: float = attrs.field(validator=attrs.validators.ge(0.0))

    @classmethod
    def create(cls, *, physics_model: Model, framequat_name: str, min_lag: float = 0.0, max_lag: float = 0.0,
               imu_bias: float = 0.0, noise: Noise | None = None) -> "ProjectedGravityObservation":
        gx, gy, gz = (float(v) for v in physics_model.opt.gravity)
        return cls(framequat_idx_range=_sensor_range(physics_model, framequat_name), gravity=(gx, gy, gz),
                   min_lag=min_lag, max_lag=max_lag, noise=noise, imu_bias=imu_bias)

    def encode(self, model: Model) -> Encoded:
        floats = [*self.gravity, float(self.min_lag), float(self.max_lag), float(self.imu_bias)]
        return C.OBS_PROJECTED_GRAVITY, [int(self.framequat_idx_range[0])], floats, 3, 7


def _as_list(v: str | Collection[str]) -> list[str]:
    return [v] if isinstance(v, str) else list(v)


@attrs.define(frozen=True, kw_only=True)
class FeetContactObservation(Observation):
    foot_left: tuple[int, ...] = attrs.field()
    foot_right: tuple[int, ...] = attrs.field()
    floor_geom: tuple[int, ...] = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, foot_left_geom_names: str | Collection[str],
               foot_right_geom_names: str | Collection[str], floor_geom_names: str | Collection[str],
               noise: Noise | None = None) -> "FeetContactObservation":
        left = [physics_model.id("geom", n) for n in _as_list(foot_left_geom_names)]
        right = [physics_model.id("geom", n) for n in _as_list(foot_right_geom_names)]
        floor = [physics_model.id("geom", n) for n in _as_list(floor_geom_names)]
        return cls(foot_left=tuple(left), foot_right=tuple(right), floor_geom=tuple(floor), noise=noise)

    def encode(self, model: Model) -> Encoded:
        if len(self.foot_left) != len(self.foot_right):
            raise ValueError("foot_left and foot_right must list the same number of geoms (jnp.stack in the reference)")
        ints = [len(self.foot_left), len(self.foot_right), len(self.floor_geom), *self.foot_left, *self.foot_right, *self.floor_geom]
        return C.OBS_FEET_CONTACT, ints, [], 2 * len(self.foot_left), 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (len(self.foot_left), 2)


@attrs.define(frozen=True, kw_only=True)
class BodyPositionObservation(Observation):
    body_idxs: tuple[int, ...] = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, body_names: Iterable[str], noise: Noise | None = None) -> "BodyPositionObservation":
        return cls(body_idxs=tuple(physics_model.id("body", n) for n in body_names), noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_BODY_POSITION, [len(self.body_idxs), *self.body_idxs], [], 3 * len(self.body_idxs), 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (len(self.body_idxs), 3)


@attrs.define(frozen=True, kw_only=True)
class BodyVelocityObservation(Observation):
    body_idxs: tuple[int, ...] = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, body_names: Iterable[str], noise: Noise | None = None) -> "BodyVelocityObservation":
        return cls(body_idxs=tuple(physics_model.id("body", n) for n in body_names), noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_BODY_VELOCITY, [len(self.body_idxs), *self.body_idxs], [], 6 * len(self.body_idxs), 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (len(self.body_idxs), 6)


@attrs.define(frozen=True, kw_only=True)
class FeetForceObservation(Observation):
    foot_left: int = attrs.field()
    foot_right: int = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, foot_left_body_name: str, foot_right_body_name: str,
               noise: Noise | None = None) -> "FeetForceObservation":
        return cls(foot_left=physics_model.id("body", foot_left_body_name),
                   foot_right=physics_model.id("body", foot_right_body_name), noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_FEET_FORCE, [self.foot_left, self.foot_right], [], 6, 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (2, 3)


@attrs.define(frozen=True, kw_only=True)
class FeetTorqueObservation(FeetForceObservation):
    def encode(self, model: Model) -> Encoded:
        return C.OBS_FEET_TORQUE, [self.foot_left, self.foot_right], [], 6, 0
