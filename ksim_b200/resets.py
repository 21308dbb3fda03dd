"""Reset plugins applied to a fresh state before the reset-time forward pass.  Mirrors ksim/resets.py:36-354."""

from __future__ import annotations

__all__ = [
    "Reset", "PlaneXYPositionReset", "HFieldXYPositionReset", "InitialMotionStateReset", "MotionReferenceData", "RandomJointPositionReset", "RandomJointVelocityReset",
    "RandomBaseVelocityXYReset", "RandomHeadingReset", "RandomHeightReset", "RandomPitchRollReset",
    "get_xy_position_reset",
]

from abc import ABC, abstractmethod

import attrs

from ksim_b200 import codes as C
from ksim_b200.mjcf import GEOM, JNT, Model


@attrs.define(frozen=True, kw_only=True)
class Reset(ABC):
    @abstractmethod
    def encode(self, model: Model) -> tuple[int, list[int], list[float]]: ...


@attrs.define(frozen=True, kw_only=True)
class PlaneXYPositionReset(Reset):
    bounds: tuple[float, float, float]
    padded_bounds: tuple[float, float, float, float]
    x_range: float = attrs.field(default=5.0)
    y_range: float = attrs.field(default=5.0)
    robot_base_height: float = attrs.field(default=0.0)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_PLANE_XY_POSITION, [], [self.bounds[2] + self.robot_base_height, self.x_range, self.y_range, *self.padded_bounds]


@attrs.define(frozen=True, kw_only=True, eq=False)
class HFieldXYPositionReset(Reset):
    """XY drawn like ``PlaneXYPositionReset``; z is the height-field cell under (x, y) plus ``z_top`` plus the base height
    (ksim/resets.py:45-89).  ``hfield_data`` is the ``[nrow, ncol]`` elevation array the reference reads from
    ``physics_model.hfield_data``; it travels in the task program, so the reset works whatever the collision floor is (height
    -field *collision* is outside the supported model subset)."""

    bounds: tuple[float, float, float, float]
    padded_bounds: tuple[float, float, float, float]
    x_range: float = attrs.field(default=5.0)
    y_range: float = attrs.field(default=5.0)
    hfield_data: object = attrs.field()
    robot_base_height: float = attrs.field(default=0.0)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        import numpy as np

        data = np.asarray(getattr(self.hfield_data, "array", self.hfield_data), dtype=np.float32)
        if data.ndim != 2 or data.size == 0:
            raise ValueError("hfield_data must be a non-empty [nrow, ncol] array")
        x_bound, y_bound, z_top, _ = self.bounds
        floats = [x_bound, y_bound, z_top, self.robot_base_height, self.x_range, self.y_range, *self.padded_bounds]
        return C.RESET_HFIELD_XY_POSITION, [int(data.shape[0]), int(data.shape[1])], [float(v) for v in floats] + data.reshape(-1).tolist()


@attrs.define(frozen=True, kw_only=True, eq=False)
class MotionReferenceData:
    """The fields of ``ksim.utils.priors.MotionReferenceData`` (priors.py:53-91) the reset reads: ``qpos`` and ``qvel`` are both
    ``[T, nq]`` (the reference stores the velocities nq wide), ``ctrl_dt`` the frame period."""

    qpos: object = attrs.field()
    qvel: object = attrs.field()
    ctrl_dt: float = attrs.field()

    @property
    def num_frames(self) -> int:
        import numpy as np

        return int(np.asarray(getattr(self.qpos, "array", self.qpos)).shape[0])


@attrs.define(frozen=True, kw_only=True, eq=False)
class InitialMotionStateReset(Reset):
    """Joint positions / velocities of a random frame of a motion bank, ``time = frame * ctrl_dt`` (ksim/resets.py:284-304).
    ``freejoint=True`` is refused: the reference then writes the bank's nq-wide ``qvel`` row into the nv-wide ``data.qvel``,
    which does not trace."""

    reference_motion: MotionReferenceData = attrs.field()
    freejoint: bool = attrs.field(default=False)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        import numpy as np

        if self.freejoint:
            raise NotImplementedError("InitialMotionStateReset(freejoint=True) is shape-inconsistent in the reference (qvel is [T, nq])")
        qpos = np.asarray(getattr(self.reference_motion.qpos, "array", self.reference_motion.qpos), dtype=np.float32)
        qvel = np.asarray(getattr(self.reference_motion.qvel, "array", self.reference_motion.qvel), dtype=np.float32)
        if qpos.ndim != 2 or qpos.shape != qvel.shape or qpos.shape[1] != model.nq or qpos.shape[0] < 1:
            raise ValueError(f"reference motion must hold qpos and qvel of shape [T, nq = {model.nq}]")
        floats = [float(self.reference_motion.ctrl_dt)] + qpos.reshape(-1).tolist() + qvel.reshape(-1).tolist()
        return C.RESET_INITIAL_MOTION_STATE, [int(qpos.shape[0]), int(model.nq)], floats


@attrs.define(frozen=True, kw_only=True)
class RandomJointPositionReset(Reset):
    scale: float = attrs.field(default=0.01)
    scale_by_curriculum: bool = attrs.field(default=True)
    zeros: tuple[float, ...] | None = attrs.field(default=None)
    mins: tuple[float, ...] | None = attrs.field(default=None)
    maxs: tuple[float, ...] | None = attrs.field(default=None)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        nj = model.nq - 7
        def vec(v: tuple[float, ...] | None) -> list[float]:
            if v is None:
                return [0.0] * nj
            if len(v) != nj:
                raise ValueError(f"expected {nj} joint values, got {len(v)}")
            return [float(x) for x in v]
        ints = [int(self.scale_by_curriculum), int(self.zeros is not None), int(self.mins is not None), int(self.maxs is not None)]
        return C.RESET_RANDOM_JOINT_POSITION, ints, [float(self.scale)] + vec(self.zeros) + vec(self.mins) + vec(self.maxs)

    @classmethod
    def create(cls, physics_model: Model, zeros: dict[str, float] | None = None, scale: float = 0.01,
               scale_by_curriculum: bool = True) -> "RandomJointPositionReset":
        joint_names = physics_model.names["joint"][1:]  # remove the root joint (resets.py:156)
        rng = physics_model.jnt_range[1:]
        zeros_values = [0.0] * len(joint_names)
        mins, maxs = [float(r[0]) for r in rng], [float(r[1]) for r in rng]
        if zeros is not None:
            for name, value in zeros.items():
                if name not in joint_names:
                    raise ValueError(f"Joint {name} not found in model. Choices are: {joint_names}")
                index = joint_names.index(name)
                if value < mins[index] or value > maxs[index]:
                    raise ValueError(f"Zero value {value} for joint {name} is out of range {mins[index]}, {maxs[index]}")
                zeros_values[index] = value
        return cls(scale=scale, zeros=tuple(zeros_values), mins=tuple(mins), maxs=tuple(maxs), scale_by_curriculum=scale_by_curriculum)


@attrs.define(frozen=True, kw_only=True)
class RandomJointVelocityReset(Reset):
    scale: float = attrs.field(default=0.01)
    scale_by_curriculum: bool = attrs.field(default=True)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_RANDOM_JOINT_VELOCITY, [int(self.scale_by_curriculum)], [float(self.scale)]


@attrs.define(frozen=True, kw_only=True)
class RandomBaseVelocityXYReset(Reset):
    """A no-op for the dynamics, exactly as on the reference's MJX path (ksim/resets.py:204-206)."""

    scale: float = attrs.field(default=0.01)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_RANDOM_BASE_VELOCITY_XY, [], [float(self.scale)]


@attrs.define(frozen=True, kw_only=True)
class RandomHeadingReset(Reset):
    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_RANDOM_HEADING, [], []


@attrs.define(frozen=True, kw_only=True)
class RandomHeightReset(Reset):
    range: tuple[float, float] = attrs.field(default=(0.0, 0.1))

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_RANDOM_HEIGHT, [], [*self.range]


@attrs.define(frozen=True, kw_only=True)
class RandomPitchRollReset(Reset):
    pitch_range: tuple[float, float] = attrs.field(default=(-0.1, 0.1))
    roll_range: tuple[float, float] = attrs.field(default=(-0.1, 0.1))

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.RESET_RANDOM_PITCH_ROLL, [], [*self.pitch_range, *self.roll_range]


def get_xy_position_reset(physics_model: Model, x_range: float = 5.0, y_range: float = 5.0, x_edge_padding: float = 5.0,
                          y_edge_padding: float = 5.0, robot_base_height: float = 0.0) -> Reset:
    """ksim/resets.py:210-280: a height-field reset when the model carries ``hfield_size`` / ``hfield_data`` (this repo's MJCF
    compiler never produces them; a caller may attach the arrays to the model), else the plane reset."""
    hsize = getattr(physics_model, "hfield_size", None)
    if hsize is not None and len(hsize) > 0:
        import numpy as np

        x_bound, y_bound, z_top, z_bottom = (float(v) for v in np.asarray(hsize).reshape(-1)[:4])
        nx, ny = int(np.asarray(physics_model.hfield_nrow).reshape(-1)[0]), int(np.asarray(physics_model.hfield_ncol).reshape(-1)[0])
        padded = (-x_bound + x_edge_padding, x_bound - x_edge_padding, -y_bound + y_edge_padding, y_bound - y_edge_padding)
        return HFieldXYPositionReset(bounds=(x_bound, y_bound, z_top, z_bottom), padded_bounds=padded, x_range=x_range, y_range=y_range,
                                     hfield_data=np.asarray(physics_model.hfield_data, dtype=np.float32).reshape(nx, ny),
                                     robot_base_height=robot_base_height)
    planes = [i for i, t in enumerate(physics_model.geom_type) if t == GEOM.PLANE]
    if not planes:
        raise ValueError("No heightfield or plane geom found in the model. MuJoCo scene missing floor!")
    x_bound = y_bound = 5.0
    z_pos = float(physics_model.geom_pos[planes[0]][2])
    padded = (-x_bound + x_edge_padding, x_bound - x_edge_padding, -y_bound + y_edge_padding, y_bound - y_edge_padding)
    return PlaneXYPositionReset(bounds=(x_bound, y_bound, z_pos), padded_bounds=padded, x_range=x_range, y_range=y_range,
                                robot_base_height=robot_base_height)


_ = JNT  # re-exported for plugin authors
