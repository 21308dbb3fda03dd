"""ksim_b200.curriculum against a literal numpy transcription of ksim/curriculum.py's jnp expressions (np.where for jnp.where,
fp32 scalars), driven by random metric sequences; plus hand-computed known answers.  The reference's tests hold no
curriculum vectors."""

import numpy as np
import pytest

from ksim_b200 import curriculum as cur

f32 = np.float32


def ref_episode_length(c, ema_prev, steps, level, episode_length):  # noqa: ANN001, ANN201
    """curriculum.py:117-149."""
    step_size = f32(1 / c.num_levels)
    ema = f32(c.ema_decay) * ema_prev + f32(1 - c.ema_decay) * episode_length
    can_step = steps == 0
    should_inc = (ema > c.increase_threshold) & can_step
    should_dec = (ema < c.decrease_threshold) & can_step
    next_steps = np.where(should_inc | should_dec, c.min_level_steps, np.maximum(steps - 1, 0))
    next_level = np.where(should_inc, level + step_size, np.where(should_dec, level - step_size, level))
    next_level = np.clip(next_level, f32(c.min_level), f32(1.0))
    next_steps = np.where(should_inc | should_dec, c.min_level_steps, next_steps)
    return f32(next_level), int(next_steps), f32(ema)


def ref_distance(c, ema_prev, counter, level, distance):  # noqa: ANN001, ANN201
    """curriculum.py:185-217."""
    ema = f32(c.ema_decay) * ema_prev + f32(1 - c.ema_decay) * distance
    steps = np.maximum(counter - 1, 0)
    new_if = np.where(ema > c.increase_threshold, np.minimum(level + f32(1.0 / c.num_levels), f32(1.0)),
                      np.where(ema < c.decrease_threshold, np.maximum(level - f32(1.0 / c.num_levels), f32(c.min_level)), level))
    new_level = np.where(steps == 0, new_if, level)
    steps = np.where(new_level != level, c.min_level_steps, steps)
    return f32(new_level), int(steps), f32(ema)


def ref_saturated(c, level, steps, deaths):  # noqa: ANN001, ANN201
    """curriculum.py:251-270."""
    if steps <= 0:
        delta = f32(1.0 / c.num_levels)
        level = np.where(deaths < c.increase_threshold, level + delta, np.where(deaths > c.decrease_threshold, level - delta, level))
        return f32(np.clip(level, f32(c.min_level), f32(1.0))), c.min_level_steps
    return f32(level), steps - 1


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stateful_curricula_follow_the_reference_expressions(seed) -> None:  # noqa: ANN001
    rs = np.random.RandomState(seed)
    el = cur.EpisodeLengthCurriculum(num_levels=10, increase_threshold=2.0, decrease_threshold=1.0, min_level_steps=2, min_level=0.1)
    do = cur.DistanceFromOriginCurriculum(num_levels=7, increase_threshold=2.5, decrease_threshold=1.5, min_level_steps=3, ema_decay=0.5)
    sw = cur.StepWhenSaturated(num_levels=5, increase_threshold=0.5, decrease_threshold=1.5, min_level_steps=1, min_level=0.2)
    s_el, s_do, s_sw = el.get_initial_state(), do.get_initial_state(), sw.get_initial_state()
    assert s_el.level == f32(0.1) and s_el.state.step_counter == 2 and s_do.state.step_counter == 3 and s_sw.state == 1
    r_el = (s_el.level, s_el.state.step_counter, s_el.state.ema_episode_length)
    r_do = (s_do.level, s_do.state.step_counter, s_do.state.ema_distance)
    r_sw = (s_sw.level, s_sw.state)
    levels = set()
    for it in range(200):
        m = {"mean_episode_length": f32(rs.uniform(1, 6) if it % 40 < 25 else rs.uniform(0, 0.8)),
             "max_distance_from_origin": f32(rs.uniform(0, 5)), "mean_terminations": f32(rs.uniform(0, 2.5))}
        s_el, s_do, s_sw = el(m, it, s_el), do(m, it, s_do), sw(m, it, s_sw)
        r_el = ref_episode_length(el, r_el[2], r_el[1], r_el[0], m["mean_episode_length"])
        r_do = ref_distance(do, r_do[2], r_do[1], r_do[0], m["max_distance_from_origin"])
        r_sw = ref_saturated(sw, r_sw[0], r_sw[1], m["mean_terminations"])
        assert (s_el.level, s_el.state.step_counter, s_el.state.ema_episode_length) == r_el
        assert (s_do.level, s_do.state.step_counter, s_do.state.ema_distance) == r_do
        assert (s_sw.level, s_sw.state) == r_sw
        assert all(0.0 <= float(s.level) <= 1.0 for s in (s_el, s_do, s_sw))
        levels.add(float(s_el.level))
    assert len(levels) > 3, "the sequence must move the level up and down"


def test_known_answers() -> None:
    c = cur.ConstantCurriculum(level=0.7)
    s = c.get_initial_state()
    assert c({}, 5, s) is s and s.level == f32(0.7)
    lin = cur.LinearCurriculum(step_size=0.25, step_every_n_epochs=2, min_level=0.1, delay_steps=3)
    s = lin.get_initial_state()
    assert s.level == f32(0.1)
    # (num_steps - 3) // 2 * 0.25 clipped to [0.1, 1]: floor division also for the negative numerator
    want = {0: 0.1, 3: 0.1, 5: 0.25, 7: 0.5, 10: 0.75, 11: 1.0, 50: 1.0}
    for steps, level in want.items():
        assert lin({}, steps, s).level == f32(level), steps
    sw = cur.StepWhenSaturated(num_levels=4, increase_threshold=1.0, decrease_threshold=2.0, min_level_steps=2)
    s = sw.get_initial_state()                       # level 0, two iterations before the first change
    s = sw({"mean_terminations": 0.0}, 0, s); assert (s.level, s.state) == (f32(0.0), 1)
    s = sw({"mean_terminations": 0.0}, 1, s); assert (s.level, s.state) == (f32(0.0), 0)
    s = sw({"mean_terminations": 0.0}, 2, s); assert (s.level, s.state) == (f32(0.25), 2)    # few deaths: up
    s = sw({"mean_terminations": 9.0}, 3, s); s = sw({"mean_terminations": 9.0}, 4, s)
    s = sw({"mean_terminations": 9.0}, 5, s); assert (s.level, s.state) == (f32(0.0), 2)     # many deaths: down, clipped at min
    s = cur.CurriculumState(level=f32(0.5), state=0)
    assert sw({"mean_terminations": 1.5}, 6, s).level == f32(0.5)                             # between the thresholds: stays


def test_metrics_may_be_tensors() -> None:
    import torch

    el = cur.EpisodeLengthCurriculum(min_level_steps=0)
    s = el({"mean_episode_length": torch.tensor(40.0)}, 0, el.get_initial_state())
    assert s.level == f32(0.01) and s.state.ema_episode_length == f32(0.1) * f32(40.0)
