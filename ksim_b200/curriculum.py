"""Curriculum policies: host mirror of ``ksim/curriculum.py`` (``SURVEY.md 8f.2``).

The reference evaluates a curriculum once per training iteration on the whole ``(N, T)`` trajectory
(``RLTask._single_update`` -> ``curriculum(trajectory, rewards, training_state, prev_state)``, rl.py:1740-1746).  The
trajectory reductions each policy needs -- mean ``Trajectory.episode_length()`` (curriculum.py:125), the largest distance
from the origin (:187), mean deaths per environment (:262) -- come out of ``b200sim_metrics`` (``B200Engine.metrics``); what
is left is a few scalar operations per iteration, done here in fp32 like the reference.  Same class names, fields, defaults
and state layout; ``__call__`` takes the metrics dict instead of the trajectory.
"""

from __future__ import annotations

__all__ = ["CurriculumState", "Curriculum", "ConstantCurriculum", "LinearCurriculum", "EpisodeLengthCurriculum",
           "DistanceFromOriginCurriculum", "StepWhenSaturated"]

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Any, Mapping

import attrs
import numpy as np

f32 = np.float32


def _scalar(x: Any) -> np.float32:  # noqa: ANN401
    """A 0-d device tensor / numpy scalar / float -> fp32 host scalar (one device-to-host read per iteration)."""
    if hasattr(x, "item"):
        x = x.item()
    return f32(x)


@dataclass(frozen=True)
class CurriculumState:
    """``ksim.curriculum.CurriculumState`` (curriculum.py:29-31)."""

    level: np.float32
    state: Any = None


@attrs.define(frozen=True, kw_only=True)
class Curriculum(ABC):
    """Curricula return a level between 0 and 1 (curriculum.py:35-52)."""

    @abstractmethod
    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState: ...

    @abstractmethod
    def get_initial_state(self) -> CurriculumState: ...


@attrs.define(frozen=True, kw_only=True)
class ConstantCurriculum(Curriculum):
    level: float = attrs.field(default=1.0, validator=attrs.validators.ge(0.0))

    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState:
        return prev_state

    def get_initial_state(self) -> CurriculumState:
        return CurriculumState(level=f32(self.level))


@attrs.define(frozen=True, kw_only=True)
class LinearCurriculum(Curriculum):
    """curriculum.py:74-96: ``(num_steps - delay_steps) // step_every_n_epochs * step_size`` clipped to [min_level, 1]."""

    step_size: float = attrs.field(default=0.01, validator=attrs.validators.ge(0.0))
    step_every_n_epochs: int = attrs.field(default=1, validator=attrs.validators.ge(1))
    min_level: float = attrs.field(default=0.0, validator=attrs.validators.ge(0.0))
    delay_steps: int = attrs.field(default=0, validator=attrs.validators.ge(0))

    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState:
        level = f32((int(num_steps) - self.delay_steps) // self.step_every_n_epochs) * f32(self.step_size)
        return CurriculumState(level=f32(np.clip(level, f32(self.min_level), f32(1.0))))

    def get_initial_state(self) -> CurriculumState:
        return CurriculumState(level=f32(self.min_level))


@dataclass(frozen=True)
class EpisodeLengthCurriculumState:
    step_counter: int
    ema_episode_length: np.float32


@attrs.define(frozen=True, kw_only=True)
class EpisodeLengthCurriculum(Curriculum):
    """curriculum.py:106-162; reads ``metrics["mean_episode_length"]`` (seconds, ``Trajectory.episode_length().mean()``)."""

    num_levels: int = attrs.field(default=100, validator=attrs.validators.ge(1))
    increase_threshold: float = attrs.field(default=3.0, validator=attrs.validators.ge(0.0))
    decrease_threshold: float = attrs.field(default=3.0, validator=attrs.validators.ge(0.0))
    min_level_steps: int = attrs.field(default=1, validator=attrs.validators.ge(0))
    min_level: float = attrs.field(default=0.0, validator=attrs.validators.ge(0.0))
    ema_decay: float = attrs.field(default=0.9, validator=attrs.validators.ge(0.0))

    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState:
        step_size = f32(1 / self.num_levels)
        episode_length = _scalar(metrics["mean_episode_length"])
        ema = f32(self.ema_decay) * prev_state.state.ema_episode_length + f32(1 - self.ema_decay) * episode_length
        steps, level = prev_state.state.step_counter, prev_state.level
        can_step = steps == 0
        should_inc = bool(ema > f32(self.increase_threshold)) and can_step
        should_dec = bool(ema < f32(self.decrease_threshold)) and can_step
        next_level = level + step_size if should_inc else (level - step_size if should_dec else level)
        next_level = f32(np.clip(next_level, f32(self.min_level), f32(1.0)))
        next_steps = self.min_level_steps if (should_inc or should_dec) else max(steps - 1, 0)
        return CurriculumState(level=next_level, state=EpisodeLengthCurriculumState(step_counter=next_steps, ema_episode_length=f32(ema)))

    def get_initial_state(self) -> CurriculumState:
        return CurriculumState(level=f32(self.min_level),
                               state=EpisodeLengthCurriculumState(step_counter=self.min_level_steps, ema_episode_length=f32(0.0)))


@dataclass(frozen=True)
class DistanceFromOriginCurriculumState:
    step_counter: int
    ema_distance: np.float32


@attrs.define(frozen=True, kw_only=True)
class DistanceFromOriginCurriculum(Curriculum):
    """curriculum.py:172-227; reads ``metrics["max_distance_from_origin"]`` (max over envs and steps of |qpos[:3]|)."""

    num_levels: int = attrs.field(default=100, validator=attrs.validators.ge(1))
    increase_threshold: float = attrs.field(default=3.0, validator=attrs.validators.ge(0.0))
    decrease_threshold: float = attrs.field(default=3.0, validator=attrs.validators.ge(0.0))
    min_level_steps: int = attrs.field(default=1, validator=attrs.validators.ge(1))
    min_level: float = attrs.field(default=0.0, validator=attrs.validators.ge(0.0))
    ema_decay: float = attrs.field(default=0.9, validator=attrs.validators.ge(0.0))

    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState:
        distance = _scalar(metrics["max_distance_from_origin"])
        ema = f32(self.ema_decay) * prev_state.state.ema_distance + f32(1 - self.ema_decay) * distance
        current = prev_state.level
        steps = max(prev_state.state.step_counter - 1, 0)
        delta = f32(1.0 / self.num_levels)
        if ema > f32(self.increase_threshold):
            candidate = min(current + delta, f32(1.0))
        elif ema < f32(self.decrease_threshold):
            candidate = max(current - delta, f32(self.min_level))
        else:
            candidate = current
        new_level = f32(candidate) if steps == 0 else current
        if new_level != current:
            steps = self.min_level_steps
        return CurriculumState(level=new_level, state=DistanceFromOriginCurriculumState(step_counter=steps, ema_distance=f32(ema)))

    def get_initial_state(self) -> CurriculumState:
        return CurriculumState(level=f32(self.min_level),
                               state=DistanceFromOriginCurriculumState(step_counter=self.min_level_steps, ema_distance=f32(0.0)))


@attrs.define(frozen=True, kw_only=True)
class StepWhenSaturated(Curriculum):
    """curriculum.py:231-277; reads ``metrics["mean_terminations"]`` (``trajectory.done.sum(-1).mean()``: deaths per env).
    The state is the number of iterations left before the level may change again."""

    num_levels: int = attrs.field()
    increase_threshold: float = attrs.field()
    decrease_threshold: float = attrs.field()
    min_level_steps: int = attrs.field()
    min_level: float = attrs.field(default=0.0, validator=attrs.validators.ge(0.0))

    def __call__(self, metrics: Mapping[str, Any], num_steps: int, prev_state: CurriculumState) -> CurriculumState:
        level, steps = prev_state.level, int(prev_state.state)
        if steps > 0:
            return CurriculumState(level=level, state=steps - 1)
        deaths = _scalar(metrics["mean_terminations"])
        delta = f32(1.0 / self.num_levels)
        if deaths < f32(self.increase_threshold):
            level = level + delta
        elif deaths > f32(self.decrease_threshold):
            level = level - delta
        return CurriculumState(level=f32(np.clip(level, f32(self.min_level), f32(1.0))), state=self.min_level_steps)

    def get_initial_state(self) -> CurriculumState:
        return CurriculumState(level=f32(self.min_level), state=self.min_level_steps)
