"""Per-environment physics randomizers.  Mirrors ksim/randomization.py:39-487 (class names, fields, key usage).

They run once per environment at initialisation (``apply_randomizations``, ksim/task/rl.py:361-372), so they
are evaluated on the host with the numpy threefry restatement (``ksim_b200.rng``), vectorised over
environments; the resulting model fields travel to the GPU as the per-env override block ``rand[N][RAND]``
the step kernel reads (``RAND_*`` codes).  Derived constants (invweight0, subtreemass) stay nominal, exactly
as in the reference (randomization.py:351-355).
"""

from __future__ import annotations

__all__ = [
    "PhysicsRandomizer", "StaticFrictionRandomizer", "ArmatureRandomizer", "MassAdditionRandomizer",
    "MassMultiplicationRandomizer", "AllBodiesMassMultiplicationRandomizer", "JointDampingRandomizer",
    "COMRandomizer", "AllBodiesCOMRandomizer", "AllBodiesInertiaRandomizer", "FloorFrictionRandomizer",
    "CollisionBodyRandomizer", "IMUAlignmentRandomizer", "FIELD_CODES", "pair_friction_unaffected",
]

import math
from abc import ABC, abstractmethod

import attrs
import numpy as np

from ksim_b200 import codes as C
from ksim_b200 import rng as R
from ksim_b200.mjcf import Model
from ksim_b200.terminations import camelcase_to_snakecase

FIELD_CODES = {
    "dof_frictionloss": C.RAND_DOF_FRICTIONLOSS,
    "dof_armature": C.RAND_DOF_ARMATURE,
    "dof_damping": C.RAND_DOF_DAMPING,
    "body_mass": C.RAND_BODY_MASS,
    "body_inertia": C.RAND_BODY_INERTIA,
    "body_ipos": C.RAND_BODY_IPOS,
    "geom_size": C.RAND_GEOM_SIZE,
    "geom_pos": C.RAND_GEOM_POS,
    "site_quat": C.RAND_SITE_QUAT,
    "site_pos": C.RAND_SITE_POS,
}

_F32 = np.float32


def _u(keys: np.ndarray, n: int, lo: float | np.ndarray, hi: float | np.ndarray) -> np.ndarray:
    """jax.random.uniform(key, (n,), lo, hi) for a batch of keys (E, 2) -> (E, n) float32."""
    return R.uniform(keys, n, lo, hi)


@attrs.define(frozen=True, kw_only=True)
class PhysicsRandomizer(ABC):
    @abstractmethod
    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        """rng: (E, 2) uint32 keys, one per environment -> {field: (E, *field_shape) float32}."""

    def get_name(self) -> str:
        return camelcase_to_snakecase(self.__class__.__name__)

    @property
    def randomization_name(self) -> str:
        return self.get_name()


def _scale_tail(base: np.ndarray, rng: np.ndarray, lo: float, hi: float) -> np.ndarray:
    base = base.astype(_F32)
    e = rng.shape[0]
    tail = base[None, 6:] * _u(rng, base.shape[0] - 6, lo, hi)
    return np.concatenate([np.broadcast_to(base[None, :6], (e, 6)), tail], axis=1).astype(_F32)


@attrs.define(frozen=True, kw_only=True)
class StaticFrictionRandomizer(PhysicsRandomizer):
    scale_lower: float = attrs.field(default=0.5)
    scale_upper: float = attrs.field(default=2.0)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        return {"dof_frictionloss": _scale_tail(model.dof_frictionloss, rng, self.scale_lower, self.scale_upper)}


@attrs.define(frozen=True, kw_only=True)
class ArmatureRandomizer(PhysicsRandomizer):
    scale_lower: float = attrs.field(default=0.95)
    scale_upper: float = attrs.field(default=1.05)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        return {"dof_armature": _scale_tail(model.dof_armature, rng, self.scale_lower, self.scale_upper)}


@attrs.define(frozen=True, kw_only=True)
class JointDampingRandomizer(PhysicsRandomizer):
    scale_lower: float = attrs.field(default=0.9)
    scale_upper: float = attrs.field(default=1.1)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        return {"dof_damping": _scale_tail(model.dof_damping, rng, self.scale_lower, self.scale_upper)}


def _body_id(model: Model, body_name: str) -> int:
    if body_name not in model.names["body"]:
        raise ValueError(f"Body name {body_name} not found in model")
    return model.id("body", body_name)


@attrs.define(frozen=True, kw_only=True)
class MassAdditionRandomizer(PhysicsRandomizer):
    body_id: int = attrs.field()
    scale_lower: float = attrs.field(default=-1.0)
    scale_upper: float = attrs.field(default=1.0)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        mass = np.broadcast_to(model.body_mass.astype(_F32)[None], (rng.shape[0], model.nbody)).copy()
        mass[:, self.body_id] = mass[:, self.body_id] + _u(rng, 1, self.scale_lower, self.scale_upper)[:, 0]
        return {"body_mass": mass}

    @classmethod
    def from_body_name(cls, model: Model, body_name: str, scale_lower: float = 0.0, scale_upper: float = 1.0) -> "MassAdditionRandomizer":
        return cls(body_id=_body_id(model, body_name), scale_lower=scale_lower, scale_upper=scale_upper)


@attrs.define(frozen=True, kw_only=True)
class MassMultiplicationRandomizer(PhysicsRandomizer):
    body_id: int = attrs.field()
    scale_lower: float = attrs.field(default=0.95)
    scale_upper: float = attrs.field(default=1.05)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        mass = np.broadcast_to(model.body_mass.astype(_F32)[None], (rng.shape[0], model.nbody)).copy()
        mass[:, self.body_id] = mass[:, self.body_id] * _u(rng, 1, self.scale_lower, self.scale_upper)[:, 0]
        return {"body_mass": mass}

    @classmethod
    def from_body_name(cls, model: Model, body_name: str, scale_lower: float = 0.0, scale_upper: float = 1.0) -> "MassMultiplicationRandomizer":
        # the reference's from_body_name defaults are (0.0, 1.0), not the class defaults (randomization.py:186-204)
        return cls(body_id=_body_id(model, body_name), scale_lower=scale_lower, scale_upper=scale_upper)


@attrs.define(frozen=True, kw_only=True)
class AllBodiesMassMultiplicationRandomizer(PhysicsRandomizer):
    scale_lower: float = attrs.field(default=0.98)
    scale_upper: float = attrs.field(default=1.02)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        return {"body_mass": (model.body_mass.astype(_F32)[None] * _u(rng, model.nbody, self.scale_lower, self.scale_upper)).astype(_F32)}


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product, wxyz, float32 (xax.quat_mul)."""
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=-1).astype(_F32)


@attrs.define(frozen=True, kw_only=True)
class IMUAlignmentRandomizer(PhysicsRandomizer):
    """Small random rotation (and optional translation) of one site's frame in its body (ksim/randomization.py:245-283):
    roll / pitch ~ N(0, tilt_std), optional yaw, composed as Z Y X of small-angle quaternions and normalised."""

    site_name: str = attrs.field(default="imu_site")
    tilt_std_rad: float = attrs.field(default=math.radians(2))
    yaw_std_rad: float | None = attrs.field(default=None)
    translate_std_m: float | None = attrs.field(default=None)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        n = rng.shape[0]
        pair = R.split(rng, 2)
        rng, sub = pair[:, 0], pair[:, 1]
        rxy = R.normal(sub, 2) * _F32(self.tilt_std_rad)
        rz = np.zeros(n, _F32)
        if self.yaw_std_rad is not None:
            pair = R.split(rng, 2)
            rng, sub = pair[:, 0], pair[:, 1]
            rz = R.normal(sub, 1)[:, 0] * _F32(self.yaw_std_rad)
        h, one, zero = _F32(0.5), np.ones(n, _F32), np.zeros(n, _F32)
        qx = np.stack([one, rxy[:, 0] * h, zero, zero], axis=-1)
        qy = np.stack([one, zero, rxy[:, 1] * h, zero], axis=-1)
        qz = np.stack([one, zero, zero, rz * h], axis=-1)
        q = _quat_mul(qz, _quat_mul(qy, qx))
        q = (q / np.sqrt((q * q).sum(axis=-1, dtype=_F32), dtype=_F32)[:, None]).astype(_F32)
        site = model.id("site", self.site_name)
        site_quat = np.broadcast_to(model.site_quat.astype(_F32)[None], (n, model.nsite, 4)).copy()
        site_quat[:, site] = _quat_mul(q, site_quat[:, site])
        out = {"site_quat": site_quat}
        if self.translate_std_m is not None and self.translate_std_m > 0.0:
            pair = R.split(rng, 2)
            rng, sub = pair[:, 0], pair[:, 1]
            site_pos = np.broadcast_to(model.site_pos.astype(_F32)[None], (n, model.nsite, 3)).copy()
            site_pos[:, site] = site_pos[:, site] + R.normal(sub, 3) * _F32(self.translate_std_m)
            out["site_pos"] = site_pos
        return out


@attrs.define(frozen=True, kw_only=True)
class COMRandomizer(PhysicsRandomizer):
    body_id: int = attrs.field()
    scale: float = attrs.field(default=0.01)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        sub = R.split(rng, 2)[:, 1]
        ipos = np.broadcast_to(model.body_ipos.astype(_F32)[None], (rng.shape[0], model.nbody, 3)).copy()
        ipos[:, self.body_id] = ipos[:, self.body_id] + _u(sub, 3, -self.scale, self.scale)
        return {"body_ipos": ipos}

    @classmethod
    def from_body_name(cls, model: Model, body_name: str, scale: float = 0.01) -> "COMRandomizer":
        return cls(body_id=_body_id(model, body_name), scale=scale)


@attrs.define(frozen=True, kw_only=True)
class AllBodiesCOMRandomizer(PhysicsRandomizer):
    scale: float = attrs.field(default=0.01)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        sub = R.split(rng, 2)[:, 1]
        off = _u(sub, model.nbody * 3, -self.scale, self.scale).reshape(rng.shape[0], model.nbody, 3)
        return {"body_ipos": (model.body_ipos.astype(_F32)[None] + off).astype(_F32)}


@attrs.define(frozen=True, kw_only=True)
class AllBodiesInertiaRandomizer(PhysicsRandomizer):
    scale: float = attrs.field(default=0.02)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        sub = R.split(rng, 2)[:, 1]
        one = _F32(1.0)
        # same key for both draws (randomization.py:361-364)
        mass_scaling = one + _u(sub, model.nbody, -self.scale, self.scale)
        pert = one + _u(sub, model.nbody * 3, -self.scale, self.scale).reshape(rng.shape[0], model.nbody, 3)
        mass = model.body_mass.astype(_F32)[None] * mass_scaling
        inertia = model.body_inertia.astype(_F32)[None] * mass_scaling[:, :, None] * pert
        return {"body_mass": mass.astype(_F32), "body_inertia": inertia.astype(_F32)}


@attrs.define(frozen=True, kw_only=True)
class FloorFrictionRandomizer(PhysicsRandomizer):
    """randomization.py:75-104: scales nothing, *sets* the floor geom's sliding friction to U(lower, upper).

    The engine's contact pairs carry their own friction, mixed from the two geoms at model-compile time by MuJoCo's rule
    (higher ``priority`` wins, else the maximum).  ``make_env_init`` therefore accepts this randomizer only when the rule
    makes it a no-op for every pair of the model (``pair_friction_unaffected``) -- the K-Bot's foot capsules have
    ``priority=1`` over the floor's 0 (robot.mjcf:23-25) -- and raises otherwise."""

    floor_geom_id: int = attrs.field()
    scale_lower: float = attrs.field(default=0.4)
    scale_upper: float = attrs.field(default=1.0)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        fr = np.broadcast_to(model.arrays["geom_friction"].astype(_F32)[None], (rng.shape[0], model.ngeom, 3)).copy()
        fr[:, self.floor_geom_id, 0] = _u(rng, 1, self.scale_lower, self.scale_upper)[:, 0]
        return {"geom_friction": fr}

    @classmethod
    def from_geom_name(cls, model: Model, floor_geom_name: str, scale_lower: float = 0.4, scale_upper: float = 1.0) -> "FloorFrictionRandomizer":
        return cls(floor_geom_id=model.id("geom", floor_geom_name), scale_lower=scale_lower, scale_upper=scale_upper)


def pair_friction_unaffected(model: Model, geom_friction: np.ndarray) -> bool:
    """True when per-env ``geom_friction`` (E, ngeom, 3) cannot change any contact pair's friction: every changed geom
    only meets geoms of strictly higher priority."""
    base = model.arrays["geom_friction"].astype(_F32)
    changed = np.nonzero((geom_friction != base[None]).any(axis=(0, 2)))[0]
    prio = model.arrays["geom_priority"]
    for g1, g2 in zip(model.arrays["pair_geom1"], model.arrays["pair_geom2"]):
        for a, b in ((g1, g2), (g2, g1)):
            if a in changed and not prio[b] > prio[a]:
                return False
    return True


@attrs.define(frozen=True, kw_only=True)
class CollisionBodyRandomizer(PhysicsRandomizer):
    """randomization.py:380-487: per-env capsule radius / half-length scaling and position jitter of the listed geoms."""

    geom_ids: tuple[int, ...] = attrs.field()
    radius_scale: float = attrs.field(default=0.05)
    length_scale: float = attrs.field(default=0.05)
    position_jitter_x: float = attrs.field(default=0.001)
    position_jitter_y: float = attrs.field(default=0.001)
    position_jitter_z: float = attrs.field(default=0.001)

    def __call__(self, model: Model, rng: np.ndarray) -> dict[str, np.ndarray]:
        e, n = rng.shape[0], len(self.geom_ids)
        pair = R.split(rng, 2); rng, radius_rng = pair[:, 0], pair[:, 1]
        pair = R.split(rng, 2); rng, length_rng = pair[:, 0], pair[:, 1]
        pos_rng = R.split(rng, 2)[:, 1]
        rs = _u(radius_rng, n, 1.0 - self.radius_scale, 1.0 + self.radius_scale)
        ls = _u(length_rng, n, 1.0 - self.length_scale, 1.0 + self.length_scale)
        jit = np.array([self.position_jitter_x, self.position_jitter_y, self.position_jitter_z], dtype=_F32)
        off = _u(pos_rng, 3 * n, np.tile(-jit, n), np.tile(jit, n)).reshape(e, n, 3)
        size = np.broadcast_to(model.arrays["geom_size"].astype(_F32)[None], (e, model.ngeom, 3)).copy()
        pos = np.broadcast_to(model.arrays["geom_pos"].astype(_F32)[None], (e, model.ngeom, 3)).copy()
        for i, g in enumerate(self.geom_ids):
            size[:, g, 0] = size[:, g, 0] * rs[:, i]
            size[:, g, 1] = size[:, g, 1] * ls[:, i]
            pos[:, g] = pos[:, g] + off[:, i]
        return {"geom_size": size, "geom_pos": pos}

    @classmethod
    def from_geom_names(cls, model: Model, geom_names: list[str] | tuple[str, ...], radius_scale: float = 0.05,
                        length_scale: float = 0.05, position_jitter_x: float = 0.001, position_jitter_y: float = 0.001,
                        position_jitter_z: float = 0.001) -> "CollisionBodyRandomizer":
        return cls(geom_ids=tuple(model.id("geom", n) for n in geom_names), radius_scale=radius_scale, length_scale=length_scale,
                   position_jitter_x=position_jitter_x, position_jitter_y=position_jitter_y, position_jitter_z=position_jitter_z)
