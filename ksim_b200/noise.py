"""Random variables, biases and noises.  Mirrors ksim/noise.py:51-229 (same class names and fields).

Instances are descriptors: sampling happens inside the fused CUDA kernel from the per-env threefry key
chain; ``encode`` turns a descriptor into the (kind, mean, std, mag) block of the task program.
"""

from __future__ import annotations

__all__ = [
    "RandomVariable", "ZeroRandomVariable", "UniformRandomVariable", "GaussianRandomVariable",
    "UniformGaussianRandomVariable", "Bias", "AdditiveBias", "MultiplicativeBias", "AdditiveGaussianBias",
    "MultiplicativeGaussianBias", "AdditiveUniformBias", "MultiplicativeUniformBias", "NoBias", "Noise",
    "AdditiveNoise", "MultiplicativeNoise", "AdditiveGaussianNoise", "MultiplicativeGaussianNoise",
    "AdditiveUniformNoise", "MultiplicativeUniformNoise", "NoNoise", "ChainedNoise", "encode_noise",
    "encode_bias", "encode_rv",
]

from abc import ABC

import attrs

from ksim_b200 import codes as C


@attrs.define(frozen=True, kw_only=True)
class RandomVariable(ABC):
    def rv_code(self) -> tuple[int, list[float]]:
        """(RV_* kind, [mean, std, mag])."""
        raise NotImplementedError


@attrs.define(frozen=True, kw_only=True)
class ZeroRandomVariable(RandomVariable):
    def rv_code(self) -> tuple[int, list[float]]:
        return C.RV_ZERO, [0.0, 0.0, 0.0]


@attrs.define(frozen=True, kw_only=True)
class UniformRandomVariable(RandomVariable):
    mean: float = attrs.field()
    mag: float = attrs.field(validator=attrs.validators.ge(0.0))

    def rv_code(self) -> tuple[int, list[float]]:
        return C.RV_UNIFORM, [float(self.mean), 0.0, float(self.mag)]


@attrs.define(frozen=True, kw_only=True)
class GaussianRandomVariable(RandomVariable):
    mean: float = attrs.field()
    std: float = attrs.field(validator=attrs.validators.ge(0.0))

    def rv_code(self) -> tuple[int, list[float]]:
        return C.RV_GAUSSIAN, [float(self.mean), float(self.std), 0.0]


@attrs.define(frozen=True, kw_only=True)
class UniformGaussianRandomVariable(RandomVariable):
    mean: float = attrs.field()
    std: float = attrs.field(validator=attrs.validators.ge(0.0))
    mag: float = attrs.field(validator=attrs.validators.ge(0.0))

    def rv_code(self) -> tuple[int, list[float]]:
        return C.RV_UNIFORM_GAUSSIAN, [float(self.mean), float(self.std), float(self.mag)]


# ---- bias ------------------------------------------------------------------------------------------


@attrs.define(frozen=True, kw_only=True)
class Bias(RandomVariable):
    @property
    def multiplicative(self) -> bool:
        return False


@attrs.define(frozen=True, kw_only=True)
class AdditiveBias(Bias):
    pass


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeBias(Bias):
    @property
    def multiplicative(self) -> bool:
        return True


@attrs.define(frozen=True, kw_only=True)
class AdditiveGaussianBias(AdditiveBias, GaussianRandomVariable):
    mean: float = attrs.field(default=0.0)


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeGaussianBias(MultiplicativeBias, GaussianRandomVariable):
    mean: float = attrs.field(default=1.0)


@attrs.define(frozen=True, kw_only=True)
class AdditiveUniformBias(AdditiveBias, UniformRandomVariable):
    mean: float = attrs.field(default=0.0)


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeUniformBias(MultiplicativeBias, UniformRandomVariable):
    mean: float = attrs.field(default=1.0)


@attrs.define(frozen=True, kw_only=True)
class NoBias(Bias):
    def rv_code(self) -> tuple[int, list[float]]:
        return C.RV_ZERO, [0.0, 0.0, 0.0]


# ---- noise -----------------------------------------------------------------------------------------


@attrs.define(frozen=True, kw_only=True)
class Noise(ABC):
    def links(self) -> list[tuple[int, int, list[float]]]:
        """[(multiplicative, RV kind, [mean, std, mag])]."""
        raise NotImplementedError


@attrs.define(frozen=True, kw_only=True)
class AdditiveNoise(Noise, RandomVariable):
    def links(self) -> list[tuple[int, int, list[float]]]:
        kind, f = self.rv_code()
        return [(0, kind, f)]


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeNoise(Noise, RandomVariable):
    def links(self) -> list[tuple[int, int, list[float]]]:
        kind, f = self.rv_code()
        return [(1, kind, f)]


@attrs.define(frozen=True, kw_only=True)
class AdditiveGaussianNoise(AdditiveNoise, GaussianRandomVariable):
    mean: float = attrs.field(default=0.0)


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeGaussianNoise(MultiplicativeNoise, GaussianRandomVariable):
    mean: float = attrs.field(default=1.0)


@attrs.define(frozen=True, kw_only=True)
class AdditiveUniformNoise(AdditiveNoise, UniformRandomVariable):
    mean: float = attrs.field(default=0.0)


@attrs.define(frozen=True, kw_only=True)
class MultiplicativeUniformNoise(MultiplicativeNoise, UniformRandomVariable):
    mean: float = attrs.field(default=1.0)


@attrs.define(frozen=True, kw_only=True)
class NoNoise(Noise):
    def links(self) -> list[tuple[int, int, list[float]]]:
        return []


@attrs.define(frozen=True, kw_only=True)
class ChainedNoise(Noise):
    noises: tuple[Noise, ...] = attrs.field()

    def links(self) -> list[tuple[int, int, list[float]]]:
        out: list[tuple[int, int, list[float]]] = []
        for n in self.noises:
            out.extend(n.links())
        return out


def encode_noise(noise: Noise | None) -> tuple[list[int], list[float]]:
    links = [] if noise is None else noise.links()
    if len(links) > C.NOISE_MAX_LINKS:
        raise ValueError(f"ChainedNoise with more than {C.NOISE_MAX_LINKS} links is not supported")
    ints = [len(links)]
    floats: list[float] = []
    for mul, kind, f in links:
        ints += [mul, kind]
        floats += f
    ints += [0, 0] * (C.NOISE_MAX_LINKS - len(links))
    floats += [0.0] * (3 * (C.NOISE_MAX_LINKS - len(links)))
    return ints, floats


def encode_bias(bias: Bias | None) -> tuple[list[int], list[float]]:
    if bias is None:
        return [C.RV_ZERO, 0], [0.0, 0.0, 0.0]
    kind, f = bias.rv_code()
    return [kind, int(bias.multiplicative)], f


def encode_rv(rv: RandomVariable | None) -> tuple[int, list[float]]:
    """kind is -1 when the variable is absent (caller-specific default applies)."""
    if rv is None:
        return -1, [0.0, 0.0, 0.0]
    return rv.rv_code()
