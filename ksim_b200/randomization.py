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
    "COMRandomizer", "AllBodiesCOMRandomizer", "AllBodiesInertiaRandomizer", "FIELD_CODES",
]

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
