"""Curriculum-level -> scalar maps.  Mirrors ksim/scales.py:22-115 (same class names and fields)."""

from __future__ import annotations

__all__ = ["Scale", "ConstantScale", "LinearScale", "QuadraticScale", "SquareRootScale", "NonZeroScale", "convert_to_scale"]

import math
from abc import ABC, abstractmethod

import attrs

from ksim_b200 import codes as C


@attrs.define(frozen=True, kw_only=True)
class Scale(ABC):
    @abstractmethod
    def get_scale(self, curriculum_level: float) -> float: ...

    @abstractmethod
    def encode(self) -> tuple[list[int], list[float]]:
        """(ints[SCALE_NI], floats[SCALE_NF]) of the task program."""


@attrs.define(frozen=True, kw_only=True)
class ConstantScale(Scale):
    scale: float = attrs.field(validator=attrs.validators.ge(0.0))

    def get_scale(self, curriculum_level: float) -> float:
        return self.scale

    def encode(self) -> tuple[list[int], list[float]]:
        return [C.SCALE_CONST], [float(self.scale), 0.0]


def _nonneg_sum(inst: object, attribute: object, value: float) -> None:
    if value + inst.bias < 0.0:  # type: ignore[attr-defined]
        raise ValueError(f"scale + bias must be non-negative, got {value} + {inst.bias}")  # type: ignore[attr-defined]


@attrs.define(frozen=True, kw_only=True)
class LinearScale(Scale):
    scale: float = attrs.field(validator=_nonneg_sum)
    bias: float = attrs.field(validator=attrs.validators.ge(0.0), default=0.0)

    def get_scale(self, curriculum_level: float) -> float:
        return self.scale * curriculum_level + self.bias

    def encode(self) -> tuple[list[int], list[float]]:
        return [C.SCALE_LINEAR], [float(self.scale), float(self.bias)]

    @classmethod
    def from_endpoints(cls, start: float, end: float) -> "LinearScale":
        return cls(scale=(end - start), bias=start)


@attrs.define(frozen=True, kw_only=True)
class QuadraticScale(Scale):
    scale: float = attrs.field(validator=_nonneg_sum)
    bias: float = attrs.field(validator=attrs.validators.ge(0.0), default=0.0)

    def get_scale(self, curriculum_level: float) -> float:
        return self.scale * curriculum_level**2 + self.bias

    def encode(self) -> tuple[list[int], list[float]]:
        return [C.SCALE_QUADRATIC], [float(self.scale), float(self.bias)]

    @classmethod
    def from_endpoints(cls, start: float, end: float) -> "QuadraticScale":
        return cls(scale=(end - start), bias=start)


@attrs.define(frozen=True, kw_only=True)
class SquareRootScale(Scale):
    scale: float = attrs.field(validator=_nonneg_sum)
    bias: float = attrs.field(validator=attrs.validators.ge(0.0), default=0.0)

    def get_scale(self, curriculum_level: float) -> float:
        return self.scale * math.sqrt(curriculum_level) + self.bias

    def encode(self) -> tuple[list[int], list[float]]:
        return [C.SCALE_SQRT], [float(self.scale), float(self.bias)]

    @classmethod
    def from_endpoints(cls, start: float, end: float) -> "SquareRootScale":
        return cls(scale=(end - start), bias=start)


@attrs.define(frozen=True, kw_only=True)
class NonZeroScale(Scale):
    scale: float = attrs.field()
    bias: float = attrs.field(validator=attrs.validators.ge(0.0), default=0.0)

    def get_scale(self, curriculum_level: float) -> float:
        return self.scale if curriculum_level > 0.0 else self.bias

    def encode(self) -> tuple[list[int], list[float]]:
        return [C.SCALE_NONZERO], [float(self.scale), float(self.bias)]


def convert_to_scale(value: float | int | Scale) -> Scale:
    if isinstance(value, (float, int)):
        return ConstantScale(scale=value)
    if isinstance(value, Scale):
        return value
    raise ValueError(f"Invalid scale: {value}")
