"""Termination plugins.  Mirrors ksim/terminations.py:33-192 (class names and fields).

Values are ternary (-1 failed / 0 running / +1 success); the predicate itself is evaluated inside the fused
CUDA step kernel (ksim_b200/csrc/b2s_engine.cuh ``termination``).  ``encode`` -> (TERM_* code, ints, floats).
"""

from __future__ import annotations

__all__ = [
    "Termination", "NotUprightTermination", "MinimumHeightTermination", "IllegalContactTermination",
    "BadZTermination", "BadVelocityTermination", "FarFromOriginTermination", "EpisodeLengthTermination",
]

import re
from abc import ABC, abstractmethod
from typing import Collection

import attrs

from ksim_b200 import codes as C
from ksim_b200.mjcf import Model


def camelcase_to_snakecase(name: str) -> str:
    return re.sub(r"(?<!^)(?=[A-Z])", "_", name).lower()


@attrs.define(frozen=True, kw_only=True)
class Termination(ABC):
    @abstractmethod
    def encode(self, model: Model) -> tuple[int, list[int], list[float]]: ...

    def get_name(self) -> str:
        return camelcase_to_snakecase(self.__class__.__name__)

    @property
    def termination_name(self) -> str:
        return self.get_name()


@attrs.define(frozen=True, kw_only=True)
class NotUprightTermination(Termination):
    max_radians: float = attrs.field(validator=attrs.validators.gt(0.0))

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_NOT_UPRIGHT, [], [float(self.max_radians)]


@attrs.define(frozen=True, kw_only=True)
class MinimumHeightTermination(Termination):
    min_height: float = attrs.field(validator=attrs.validators.gt(0.0))

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_MINIMUM_HEIGHT, [], [float(self.min_height)]


@attrs.define(frozen=True, kw_only=True)
class IllegalContactTermination(Termination):
    illegal_geom_idxs: tuple[int, ...] = attrs.field(converter=lambda v: tuple(int(x) for x in v))
    contact_eps: float = -0.001

    @classmethod
    def create(cls, physics_model: Model, geom_names: Collection[str], contact_eps: float = -1e-3) -> "IllegalContactTermination":
        remaining = set(geom_names)
        idxs = []
        for idx, name in enumerate(physics_model.names["geom"]):
            if name in remaining:
                idxs.append(idx)
                remaining.remove(name)
        if remaining:
            raise ValueError(f"Geoms {remaining} not found in model. Choices are: {sorted(physics_model.names['geom'])}")
        return cls(contact_eps=contact_eps, illegal_geom_idxs=tuple(idxs))

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_ILLEGAL_CONTACT, [len(self.illegal_geom_idxs), *self.illegal_geom_idxs], [float(self.contact_eps)]


@attrs.define(frozen=True, kw_only=True)
class BadZTermination(Termination):
    min_z: float = attrs.field()
    final_min_z: float | None = attrs.field(default=None)
    max_z: float = attrs.field()
    final_max_z: float | None = attrs.field(default=None)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        fmin = self.min_z if self.final_min_z is None else self.final_min_z
        fmax = self.max_z if self.final_max_z is None else self.final_max_z
        return C.TERM_BAD_Z, [], [float(self.min_z), float(fmin), float(self.max_z), float(fmax)]


@attrs.define(frozen=True, kw_only=True)
class BadVelocityTermination(Termination):
    max_vel: float = attrs.field()

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_BAD_VELOCITY, [], [float(self.max_vel)]


@attrs.define(frozen=True, kw_only=True)
class FarFromOriginTermination(Termination):
    max_dist: float = attrs.field(validator=attrs.validators.gt(0.0))
    pos_termination: bool = attrs.field(default=True)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        return C.TERM_FAR_FROM_ORIGIN, [1 if self.pos_termination else -1], [float(self.max_dist)]


@attrs.define(frozen=True, kw_only=True)
class EpisodeLengthTermination(Termination):
    max_length_sec: float = attrs.field(validator=attrs.validators.gt(0.0))
    disable_at_curriculum_level: int | None = attrs.field(default=None)
    pos_termination: bool = attrs.field(default=True)

    def encode(self, model: Model) -> tuple[int, list[int], list[float]]:
        dis = -1 if self.disable_at_curriculum_level is None else int(self.disable_at_curriculum_level)
        return C.TERM_EPISODE_LENGTH, [1 if self.pos_termination else -1, dis], [float(self.max_length_sec)]
