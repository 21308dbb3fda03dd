"""Test helpers: a tiny task around the TRIPOD model whose rewards / terminations can be chosen per test."""

from __future__ import annotations

import numpy as np

from ksim_b200 import commands, observation, resets
from ksim_b200.actuators import TorqueActuators
from ksim_b200.mjcf import load_mjcf_string
from ksim_b200.program import EngineConfig, TaskProgram, TaskSpec, build_program
from tests.models import TRIPOD


def tripod_model():  # noqa: ANN201
    return load_mjcf_string(TRIPOD)


def tripod_spec(rewards=None, terminations=None, events=None, observations=None, cmds=None, **cfg) -> TaskSpec:  # noqa: ANN001, ANN003
    model = tripod_model()
    config = EngineConfig(dt=0.004, ctrl_dt=0.02, **cfg)
    obs = observations
    if obs is None:
        obs = {
            "joint_position": observation.JointPositionObservation(),
            "base_position": observation.BasePositionObservation(),
            "feet_contact": observation.FeetContactObservation.create(
                physics_model=model, foot_left_geom_names=["foot_a"], foot_right_geom_names=["foot_b"], floor_geom_names=["floor"]),
        }
    if cmds is None:
        cmds = {"vec": commands.FloatVectorCommand(ranges=((0.0, 1.0), (0.0, 1.0)))}
    return TaskSpec(model=model, config=config, actuators=TorqueActuators(), resets=[resets.RandomJointPositionReset.create(model)],
                    events=events or {}, observations=obs, commands=cmds, rewards=rewards or {},
                    terminations=terminations or {})


def synthetic_records(program: TaskProgram, actions: np.ndarray, done: np.ndarray, dtype=np.float32) -> np.ndarray:  # noqa: ANN001
    """rec[T][1][R] with only the action and done fields filled (what the reference's reward tests build)."""
    t = actions.shape[0]
    rec = np.zeros((t, 1, program.rec_size), dtype=dtype)
    a0 = program["TI_R_ACTION"]
    rec[:, 0, a0 : a0 + actions.shape[1]] = actions
    rec[:, 0, program["TI_R_DONE"]] = done
    return rec
