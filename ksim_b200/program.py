"""Task-program builder: plugin lists -> the flat ``ti[]`` / ``tf[]`` arrays the kernels interpret.

This is the host-side image of ``RLTask._get_constants`` (ksim/task/rl.py: the ``RolloutConstants`` that hold
engine, observations, commands, rewards, terminations) plus ``EngineConfig`` (ksim/engine.py:50-139) and the
rollout-relevant ``RLConfig`` fields (ksim/task/rl.py:377-602: action_clip_max, reward_clip_min/max).
Encoding: ``include/b200sim_codes.h``.
"""

from __future__ import annotations

__all__ = ["EngineConfig", "TaskSpec", "TaskProgram", "build_program"]

import math
from dataclasses import dataclass, field
from typing import Mapping

import numpy as np

from ksim_b200 import codes as C
from ksim_b200.actuators import Actuators
from ksim_b200.commands import Command
from ksim_b200.events import Event
from ksim_b200.mjcf import Model
from ksim_b200.noise import (
    GaussianRandomVariable, RandomVariable, UniformGaussianRandomVariable, UniformRandomVariable, ZeroRandomVariable,
    encode_bias, encode_noise,
)
from ksim_b200.observation import Observation
from ksim_b200.randomization import FIELD_CODES
from ksim_b200.resets import Reset
from ksim_b200.rewards import Reward
from ksim_b200.terminations import Termination


@dataclass
class EngineConfig:
    """ksim/engine.py:50-139 (same field names, same validation)."""

    dt: float
    ctrl_dt: float
    action_latency_range: tuple[float, float] = (0.0, 0.0)
    drop_action_prob: float = 0.0
    zero_offset_std: float | None = None
    zero_offset_mag: float | None = None
    actuator_update_dt: float | None = None
    # RLConfig fields the rollout consumes (ksim/task/rl.py)
    action_clip_max: float = 1e6
    reward_clip_min: float | None = None
    reward_clip_max: float | None = None
    # Which Newton iteration the CUDA constraint solver runs (b200sim_model_set_option, B2S_OPT_SOLVER):
    #   "reference": MJX's shape -- primal Newton in dof space, bracketing line search, both capped at the model's
    #                iterations / ls_iterations: reproduces the reference's (possibly unconverged) iterate;
    #   "fast":      constraint-space / factor-reusing Newton with an exact one-pass line search: converges further within the
    #                same caps, so it departs from the reference's iterate wherever that one stops at its cap (K-Bot's 40-70
    #                rows: ~1.5 % of env-steps beyond the stated tolerance; humanoid: within it, tests/test_gpu_parity.py).
    solver: str = "fast"

    def __post_init__(self) -> None:
        if self.solver not in ("fast", "reference"):
            raise ValueError("`solver` must be 'fast' or 'reference'")
        lo, hi = self.action_latency_range
        if lo < 0:
            raise ValueError("`min_action_latency` must be non-negative")
        if lo > hi:
            raise ValueError("`min_action_latency` must be less than or equal to `max_action_latency`")
        if self.drop_action_prob < 0 or self.drop_action_prob >= 1:
            raise ValueError("`drop_action_prob` must be between 0 and 1")

    @property
    def min_action_latency_step(self) -> float:
        return self.action_latency_range[0] / self.dt

    @property
    def max_action_latency_step(self) -> float:
        return self.action_latency_range[1] / self.dt

    @property
    def phys_steps_per_ctrl_steps(self) -> int:
        return round(self.ctrl_dt / self.dt)

    @property
    def phys_steps_per_actuator_step(self) -> int:
        if self.actuator_update_dt is not None:
            return max(1, round(self.actuator_update_dt / self.dt))
        return 1

    def zero_offset_random_variable(self) -> RandomVariable:
        if self.zero_offset_mag is not None and self.zero_offset_std is not None:
            return UniformGaussianRandomVariable(mean=0.0, std=self.zero_offset_std, mag=self.zero_offset_mag)
        if self.zero_offset_std is not None:
            return GaussianRandomVariable(mean=0.0, std=self.zero_offset_std)
        if self.zero_offset_mag is not None:
            return UniformRandomVariable(mean=0.0, mag=self.zero_offset_mag)
        return ZeroRandomVariable()


@dataclass
class TaskSpec:
    """The plugin lists a task supplies (the ``get_*`` hooks of RLTask, ksim/task/rl.py:774-960)."""

    model: Model
    config: EngineConfig
    actuators: Actuators
    resets: list[Reset] = field(default_factory=list)
    events: Mapping[str, Event] = field(default_factory=dict)
    observations: Mapping[str, Observation] = field(default_factory=dict)
    commands: Mapping[str, Command] = field(default_factory=dict)
    rewards: Mapping[str, Reward] = field(default_factory=dict)
    terminations: Mapping[str, Termination] = field(default_factory=dict)
    randomized_fields: tuple[str, ...] = ()
    # Trajectory.aux_outputs (types.py:48-70): name -> floats per (env, step).  The slots live in every record and are filled
    # by the postprocess_trajectory hook (rl.py:1590; AMP's discriminator logits amp.py:327-358, a teacher's outputs
    # teacher.py:113-136) between the rollout and the rewards pass; rewards read them by name (AMPReward).
    aux_outputs: Mapping[str, int] = field(default_factory=dict)


@dataclass
class TaskProgram:
    ti: np.ndarray
    tf: np.ndarray
    spec: TaskSpec
    obs_slices: dict[str, tuple[int, int, tuple[int, ...]]]   # name (incl. noisy_*) -> (offset in obs block, dim, shape)
    cmd_slices: dict[str, tuple[int, int]]
    event_slices: dict[str, tuple[int, int]]
    reward_components: list[str]
    termination_names: list[str]
    rand_slices: dict[str, tuple[int, int]]
    aux_slices: dict[str, tuple[int, int]] = field(default_factory=dict)

    def __getitem__(self, key: str) -> int:
        return int(self.ti[getattr(C, key)])

    @property
    def state_size(self) -> int:
        return self["TI_STATE_SIZE"]

    @property
    def rec_size(self) -> int:
        return self["TI_REC_SIZE"]

    @property
    def key_size(self) -> int:
        return self["TI_KEY_SIZE"]

    @property
    def rand_size(self) -> int:
        return self["TI_RAND_SIZE"]

    @property
    def action_size(self) -> int:
        return self["TI_ACTION_SIZE"]

    @property
    def obs_size(self) -> int:
        return self["TI_OBS_SIZE"]


class _Ctx:
    def __init__(self, model: Model, obs: dict[str, tuple[int, int, tuple[int, ...]]], cmds: Mapping[str, Command],
                 cmd_slices: dict[str, tuple[int, int]], aux_slices: dict[str, tuple[int, int]] | None = None) -> None:
        self.model, self.obs, self.cmds, self.cmd_slices, self.aux_slices = model, obs, cmds, cmd_slices, aux_slices or {}
        self.used_obs: set[tuple[int, int]] = set()  # (offset, dim) of every observation a reward asked for

    def aux_offset(self, name: str) -> int:
        """Float offset of an aux output inside a record's aux region."""
        if name not in self.aux_slices:
            raise KeyError(f"aux output '{name}' not found; the task's aux_outputs are {sorted(self.aux_slices)}")
        return self.aux_slices[name][0]

    def obs_offset(self, name: str) -> int:
        if name not in self.obs:
            raise KeyError(f"observation '{name}' not found; choices: {sorted(self.obs)}")
        self.used_obs.add((int(self.obs[name][0]), int(self.obs[name][1])))
        return self.obs[name][0]

    def obs_shape(self, name: str) -> tuple[int, ...]:
        self.obs_offset(name)
        return self.obs[name][2]

    def cmd_offset(self, name: str, field_name: str | None = None) -> int:
        if name not in self.cmd_slices:
            raise KeyError(f"command '{name}' not found; choices: {sorted(self.cmd_slices)}")
        off = self.cmd_slices[name][0]
        if field_name is not None:
            off += self.cmds[name].fields()[field_name]
        return off

    def cmd_type(self, name: str) -> type:
        return type(self.cmds[name])


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def build_program(spec: TaskSpec) -> TaskProgram:
    m, cfg = spec.model, spec.config
    ints: list[int] = [0] * C.TI_HEADER_SIZE
    floats: list[float] = []

    def put(i: list[int], f: list[float]) -> tuple[int, int]:
        io, fo = len(ints), len(floats)
        ints.extend(int(v) for v in i)
        floats.extend(float(v) for v in f)
        return io, fo

    def table(n: int) -> int:
        base = len(ints)
        ints.extend([0] * (n * C.TAB_STRIDE))
        return base

    def entry(base: int, idx: int, **kw: int) -> None:
        for k, v in kw.items():
            ints[base + idx * C.TAB_STRIDE + getattr(C, "TAB_" + k)] = int(v)

    ints[C.TI_MAGIC] = C.B2S_TASK_MAGIC
    ints[C.TI_NSUB] = cfg.phys_steps_per_ctrl_steps
    ints[C.TI_ACT_PERIOD] = cfg.phys_steps_per_actuator_step
    ints[C.TI_ACT_TYPE] = spec.actuators.act_type
    na = spec.actuators.action_size(m)
    ints[C.TI_ACTION_SIZE] = na
    ai, af = spec.actuators.encode(m)
    ints[C.TI_ACT_IOFF], ints[C.TI_ACT_FOFF] = put(ai, af)
    act_state_size = spec.actuators.state_size(m)
    ints[C.TI_ACT_STATE_SIZE] = act_state_size

    zkind, zf = cfg.zero_offset_random_variable().rv_code()
    ints[C.TI_ZERO_OFFSET_KIND] = zkind
    nan = float("nan")
    _, fo = put([], [cfg.min_action_latency_step, cfg.max_action_latency_step, cfg.drop_action_prob, *zf,
                     cfg.action_clip_max, nan if cfg.reward_clip_min is None else cfg.reward_clip_min,
                     nan if cfg.reward_clip_max is None else cfg.reward_clip_max])
    ints[C.TI_ENGINE_FOFF] = fo

    # events
    ev_names = list(spec.events)
    ints[C.TI_N_EVENT] = len(ev_names)
    ints[C.TI_EVENT_TAB] = base = table(len(ev_names))
    event_slices: dict[str, tuple[int, int]] = {}
    eoff = 0
    for k, name in enumerate(ev_names):
        ev = spec.events[name]
        code, ei, ef = ev.encode(m)
        io, fo = put(ei, ef)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo, OUT_OFF=eoff, OUT_DIM=ev.state_size())
        event_slices[name] = (eoff, ev.state_size())
        eoff += ev.state_size()
    ints[C.TI_EVENT_SIZE] = eoff

    # resets
    ints[C.TI_N_RESET] = len(spec.resets)
    ints[C.TI_RESET_TAB] = base = table(len(spec.resets))
    for k, rs in enumerate(spec.resets):
        code, ri, rf = rs.encode(m)
        io, fo = put(ri, rf)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo)

    # observations: first the plain values, then the noisy copies
    obs_names = list(spec.observations)
    ints[C.TI_N_OBS] = len(obs_names)
    ints[C.TI_OBS_TAB] = base = table(len(obs_names))
    encoded = [spec.observations[n].encode(m) for n in obs_names]
    obs_slices: dict[str, tuple[int, int, tuple[int, ...]]] = {}
    ooff = 0
    for name, (code, _, _, dim, _) in zip(obs_names, encoded):
        obs_slices[name] = (ooff, dim, spec.observations[name].shape(m))
        ooff += dim
    carry_off = 0
    for k, (name, (code, oi, of, dim, carry)) in enumerate(zip(obs_names, encoded)):
        ob = spec.observations[name]
        noff = -1
        if ob.has_noisy_copy:
            noff = ooff
            obs_slices["noisy_" + name] = (ooff, dim, ob.shape(m))
            ooff += dim
        ni, nf = encode_noise(ob.noise)
        bi, bf = encode_bias(ob.bias)
        io, fo = put(ni + bi + oi, nf + bf + of)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo, OUT_OFF=obs_slices[name][0], OUT_DIM=dim, AUX0=noff,
              AUX1=(carry_off if carry > 0 else -1), AUX2=int(ob.bias is not None))
        carry_off += carry
    ints[C.TI_OBS_SIZE] = ooff
    ints[C.TI_OBS_CARRY_SIZE] = carry_off

    # terminations
    term_names = list(spec.terminations)
    ints[C.TI_N_TERM] = len(term_names)
    ints[C.TI_TERM_TAB] = base = table(len(term_names))
    for k, name in enumerate(term_names):
        code, xi, xf = spec.terminations[name].encode(m)
        io, fo = put(xi, xf)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo)

    # commands
    cmd_names = list(spec.commands)
    ints[C.TI_N_CMD] = len(cmd_names)
    ints[C.TI_CMD_TAB] = base = table(len(cmd_names))
    cmd_slices: dict[str, tuple[int, int]] = {}
    coff = 0
    for k, name in enumerate(cmd_names):
        code, ci, cf, dim = spec.commands[name].encode()
        io, fo = put(ci, cf)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo, OUT_OFF=coff, OUT_DIM=dim)
        cmd_slices[name] = (coff, dim)
        coff += dim
    ints[C.TI_CMD_SIZE] = coff

    # aux outputs (filled by the postprocess_trajectory hook; read by rewards)
    aux_slices: dict[str, tuple[int, int]] = {}
    aoff = 0
    for name, dim in spec.aux_outputs.items():
        if int(dim) < 1:
            raise ValueError(f"aux output '{name}' needs at least one float per step")
        aux_slices[name] = (aoff, int(dim))
        aoff += int(dim)

    # rewards
    ctx = _Ctx(m, obs_slices, spec.commands, cmd_slices, aux_slices)
    rew_names = list(spec.rewards)
    ints[C.TI_N_REW] = len(rew_names)
    ints[C.TI_REW_TAB] = base = table(len(rew_names))
    comp_names: list[str] = []
    rcarry = 0
    for k, name in enumerate(rew_names):
        rw = spec.rewards[name]
        code, xi, xf, comps, carry = rw.encode(ctx)
        si, sf = rw.scale.encode()
        io, fo = put(si + xi, sf + xf)
        entry(base, k, TYPE=code, IOFF=io, FOFF=fo, OUT_OFF=len(comp_names), OUT_DIM=len(comps),
              AUX0=(rcarry if carry > 0 else -1), AUX1=carry, AUX2=int(rw.is_penalty))
        comp_names += [name if c == "" else f"{name}/{c}" for c in comps]
        rcarry += carry
    ints[C.TI_N_REW_COMP] = len(comp_names)
    ints[C.TI_REW_CARRY_SIZE] = rcarry

    # per-env model overrides
    field_sizes = {"dof_frictionloss": m.nv, "dof_armature": m.nv, "dof_damping": m.nv, "body_mass": m.nbody,
                   "body_inertia": 3 * m.nbody, "body_ipos": 3 * m.nbody, "geom_size": 3 * m.ngeom, "geom_pos": 3 * m.ngeom,
                   "site_quat": 4 * m.nsite, "site_pos": 3 * m.nsite}
    ints[C.TI_N_RAND] = len(spec.randomized_fields)
    ints[C.TI_RAND_TAB] = base = table(len(spec.randomized_fields))
    rand_slices: dict[str, tuple[int, int]] = {}
    roff = 0
    for k, fname in enumerate(spec.randomized_fields):
        if fname not in FIELD_CODES:
            raise NotImplementedError(f"per-env override of model field '{fname}' is not supported by the CUDA engine")
        entry(base, k, AUX0=FIELD_CODES[fname], AUX1=field_sizes[fname])
        rand_slices[fname] = (roff, field_sizes[fname])
        roff += field_sizes[fname]
    ints[C.TI_RAND_SIZE] = _pad4(roff)

    # state layout
    off = 0

    def alloc(slot: int, n: int) -> None:
        nonlocal off
        ints[slot] = off
        off += n

    alloc(C.TI_S_QPOS, m.nq); alloc(C.TI_S_QVEL, m.nv); alloc(C.TI_S_QACC, m.nv); alloc(C.TI_S_WARM, m.nv)
    alloc(C.TI_S_TIME, 1); alloc(C.TI_S_CTRL, m.nu); alloc(C.TI_S_LAST_CTRL, na); alloc(C.TI_S_LATENCY, 1)
    alloc(C.TI_S_ZERO_OFF, m.nu); alloc(C.TI_S_EVENT, eoff); alloc(C.TI_S_ACT_STATE, act_state_size)
    alloc(C.TI_S_LEVEL, 1); alloc(C.TI_S_CMD, coff); alloc(C.TI_S_OBS_CARRY, carry_off); alloc(C.TI_S_NONFINITE, 1)
    ints[C.TI_S_XFRC] = -1
    ints[C.TI_STATE_SIZE] = _pad4(off)
    ints[C.TI_KEY_SIZE] = _pad4(2 + 2 * len(obs_names))

    # record layout: [pre-step block | post-step block]
    off = 0
    alloc(C.TI_R_OBS, ooff); alloc(C.TI_R_CMD, coff); alloc(C.TI_R_LEVEL, 1)
    ints[C.TI_REC_PRE_SIZE] = off
    alloc(C.TI_R_QPOS, m.nq); alloc(C.TI_R_QVEL, m.nv); alloc(C.TI_R_XPOS, 3 * m.nbody); alloc(C.TI_R_XQUAT, 4 * m.nbody)
    alloc(C.TI_R_CTRL, m.nu); alloc(C.TI_R_EVENT, eoff); alloc(C.TI_R_ACTION, na); alloc(C.TI_R_DONE, 1)
    alloc(C.TI_R_SUCCESS, 1); alloc(C.TI_R_TIME, 1); alloc(C.TI_R_TERM, len(term_names))
    alloc(C.TI_R_AUX, aoff)
    ints[C.TI_AUX_SIZE] = aoff
    ints[C.TI_REC_SIZE] = _pad4(off)

    # Scoring read-set (b200sim_score stages these columns of an env's T records in shared memory with bulk copies): the row from
    # the command block to its end (commands, level, qpos, qvel, ..., action, flags, aux outputs) plus every observation a reward
    # named through ctx.obs_offset, as 16-byte aligned, merged segments.  A column outside the set is still read correctly (from
    # global memory); the set only decides what is fast.
    rsize = ints[C.TI_REC_SIZE]
    spans = sorted([(ints[C.TI_R_CMD] & ~3, rsize)] + [((ints[C.TI_R_OBS] + o) & ~3, min(_pad4(ints[C.TI_R_OBS] + o + d), rsize))
                                                      for o, d in ctx.used_obs])
    segs: list[list[int]] = []
    for lo, hi in spans:
        if segs and lo <= segs[-1][1]:
            segs[-1][1] = max(segs[-1][1], hi)
        else:
            segs.append([lo, hi])
    width = sum(hi - lo for lo, hi in segs)
    ints[C.TI_SCORE_NSEG] = len(segs)
    ints[C.TI_SCORE_SEG], _ = put([v for lo, hi in segs for v in (lo, hi - lo)], [])
    ints[C.TI_SCORE_WIDTH] = width if (width // 4) % 2 == 1 else width + 4  # width / 4 odd: 4-way instead of 8-way bank conflicts

    tf = np.array(floats, dtype=np.float64)
    tf32 = tf.astype(np.float32)
    tf32[np.isinf(tf)] = np.sign(tf[np.isinf(tf)]) * np.float32(np.finfo(np.float32).max)
    assert not any(math.isinf(v) for v in tf32.tolist())
    return TaskProgram(ti=np.array(ints, dtype=np.int32), tf=tf32, spec=spec, obs_slices=obs_slices, cmd_slices=cmd_slices,
                       event_slices=event_slices, reward_components=comp_names, termination_names=term_names,
                       rand_slices=rand_slices, aux_slices=aux_slices)
