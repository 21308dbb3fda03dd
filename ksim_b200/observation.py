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
    "FeetTorqueObservation", "TimestepObservation", "DelayedJointPositionObservation", "DelayedJointVelocityObservation",
    "ContactObservation", "BodyOrientationObservation", "SiteOrientationObservation", "FeetOrientationObservation",
    "ActPosObservation",
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
class DelayedJointPositionObservation(StatefulObservation):
    """``qpos[7:] - zero_offset`` as it was ``delay_steps`` observations ago (ksim/observation.py:229-252): the carry is a ring of
    ``delay_steps`` rows, every row initialised to the joint positions right after the reset."""

    delay_steps: int = attrs.field(default=1, validator=attrs.validators.ge(1))

    def encode(self, model: Model) -> Encoded:
        n = model.nq - 7
        return C.OBS_DELAYED_JOINT_POSITION, [int(self.delay_steps)], [], n, int(self.delay_steps) * n


@attrs.define(frozen=True, kw_only=True)
class JointVelocityObservation(Observation):
    encode = _simple(C.OBS_JOINT_VELOCITY, lambda m: m.nv - 6)


@attrs.define(frozen=True, kw_only=True)
class DelayedJointVelocityObservation(StatefulObservation):
    """``qvel[6:]`` as it was ``delay_steps`` observations ago (ksim/observation.py:261-283)."""

    delay_steps: int = attrs.field(default=1, validator=attrs.validators.ge(1))

    def encode(self, model: Model) -> Encoded:
        n = model.nv - 6
        return C.OBS_DELAYED_JOINT_VELOCITY, [int(self.delay_steps)], [], n, int(self.delay_steps) * n


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


@attrs.define(frozen=True, kw_only=True)
class ContactObservation(Observation):
    """Per listed geom: is it in contact with any other listed geom (ksim/observation.py:452-478)."""

    geom_idxs: tuple[int, ...] = attrs.field()
    contact_group: str | None = attrs.field(default=None)

    @classmethod
    def create(cls, *, physics_model: Model, geom_names: str | Collection[str], contact_group: str | None = None,
               noise: Noise | None = None) -> "ContactObservation":
        return cls(geom_idxs=tuple(physics_model.id("geom", n) for n in _as_list(geom_names)), contact_group=contact_group, noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_CONTACT, [len(self.geom_idxs), *self.geom_idxs], [], len(self.geom_idxs), 0


@attrs.define(frozen=True, kw_only=True)
class BodyOrientationObservation(Observation):
    """``xquat`` of the listed bodies (ksim/observation.py:619-642)."""

    body_ids: tuple[int, ...] = attrs.field()
    name: str = attrs.field(default="body_orientation")

    @classmethod
    def create(cls, *, physics_model: Model, body_names: Collection[str], noise: Noise | None = None):  # noqa: ANN206
        ids = tuple(physics_model.id("body", n) for n in body_names)
        snake = "".join("_" + ch.lower() if ch.isupper() else ch for ch in cls.__name__).lstrip("_")
        return cls(body_ids=ids, name=f"{snake}_{'_'.join(body_names)}", noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_BODY_ORIENTATION, [len(self.body_ids), *self.body_ids], [], 4 * len(self.body_ids), 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (len(self.body_ids), 4)


@attrs.define(frozen=True, kw_only=True)
class FeetOrientationObservation(BodyOrientationObservation):
    """``BodyOrientationObservation`` of exactly two bodies, left then right (ksim/observation.py:672-703)."""

    @classmethod
    def create(cls, *, physics_model: Model, body_names: Collection[str], noise: Noise | None = None):  # noqa: ANN206
        if len(body_names) != 2:
            raise ValueError("FeetOrientationObservation expects exactly two body names (left and right foot).")
        return super().create(physics_model=physics_model, body_names=body_names, noise=noise)

    @classmethod
    def create_from_feet(cls, *, physics_model: Model, foot_left_body_name: str, foot_right_body_name: str,
                         noise: Noise | None = None):  # noqa: ANN206
        return super().create(physics_model=physics_model, body_names=(foot_left_body_name, foot_right_body_name), noise=noise)


@attrs.define(frozen=True, kw_only=True)
class SiteOrientationObservation(Observation):
    """6-D rotation of the listed sites (ksim/observation.py:645-669).  ``xax.rotation_matrix_to_rotation6d`` is un-vendored; it
    is restated from memory as the first two ROWS of the matrix, flattened (the Zhou et al. convention pytorch3d also uses)."""

    site_ids: tuple[int, ...] = attrs.field()
    name: str = attrs.field(default="site_orientation")

    @classmethod
    def create(cls, *, physics_model: Model, site_names: tuple[str, ...], noise: Noise | None = None) -> "SiteOrientationObservation":
        ids = tuple(physics_model.id("site", n) for n in site_names)
        return cls(site_ids=ids, name=f"site_orientation_observation_{'_'.join(site_names)}", noise=noise)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_SITE_ORIENTATION, [len(self.site_ids), *self.site_ids], [], 6 * len(self.site_ids), 0

    def shape(self, model: Model) -> tuple[int, ...]:
        return (len(self.site_ids), 6)


@attrs.define(frozen=True)
class ActPosObservation(Observation):
    """(last control of one actuator, position of its joint): the reference's debugging observation (ksim/observation.py:717-747)."""

    joint_name: str = attrs.field()
    ctrl_idx: int = attrs.field()
    qpos_idx: int = attrs.field()

    @classmethod
    def create(cls, *, physics_model: Model, joint_name: str) -> "ActPosObservation":
        qpos_idx = int(physics_model.jnt_qposadr[physics_model.id("joint", joint_name)])
        return cls(joint_name=joint_name, ctrl_idx=physics_model.id("actuator", f"{joint_name}_ctrl"), qpos_idx=qpos_idx)

    def encode(self, model: Model) -> Encoded:
        return C.OBS_ACT_POS, [int(self.ctrl_idx), int(self.qpos_idx)], [], 2, 0
