"""Actuator models.  Mirrors ksim/actuators.py:28-293 (class names, constructor arguments, error behaviour).

The torque computation itself runs fused inside the CUDA control-step kernel (ksim_b200/csrc/step.cu);
these classes validate the configuration and encode it into the task program.
"""

from __future__ import annotations

__all__ = ["Actuators", "StatefulActuators", "TorqueActuators", "PositionActuators", "PositionVelocityActuator"]

import logging
import math
from abc import ABC

from ksim_b200 import codes as C
from ksim_b200.mjcf import Model
from ksim_b200.noise import Noise, NoNoise, RandomVariable, UniformRandomVariable, encode_noise, encode_rv
from ksim_b200.types import Metadata

logger = logging.getLogger(__name__)


class Actuators(ABC):
    """Collection of actuators."""

    act_type: int = -1

    def action_size(self, model: Model) -> int:
        return model.nu

    def state_size(self, model: Model) -> int:
        return 0

    def encode(self, model: Model) -> tuple[list[int], list[float]]:
        raise NotImplementedError


class StatefulActuators(Actuators):
    pass


class TorqueActuators(Actuators):
    """Direct torque control (ksim/actuators.py:70-87)."""

    act_type = C.ACT_TORQUE

    def __init__(self, noise: Noise | None = None) -> None:
        super().__init__()
        self.noise = NoNoise() if noise is None else noise

    def encode(self, model: Model) -> tuple[list[int], list[float]]:
        return encode_noise(self.noise)


class PositionActuators(StatefulActuators):
    """MIT Cheetah-style PD controller operating on position (ksim/actuators.py:98-238)."""

    act_type = C.ACT_POSITION

    def __init__(
        self,
        physics_model: Model,
        metadata: Metadata,
        action_noise: Noise | None = None,
        torque_noise: Noise | None = None,
        action_scale: float = 1.0,
        action_bias: RandomVariable | None = None,
        torque_bias: RandomVariable | None = None,
        kp_scale: float | RandomVariable | None = None,
        kd_scale: float | RandomVariable | None = None,
        torque_limit_scale: float | RandomVariable | None = None,
    ) -> None:
        ctrl_name_to_idx = {name: i for i, name in enumerate(physics_model.names["actuator"])}
        n = len(ctrl_name_to_idx)
        kps, kds, clips = [-1.0] * n, [-1.0] * n, [math.inf] * n
        if metadata.joint_name_to_metadata is None:
            raise ValueError("Joint metadata is required for MITPositionActuators")
        for joint_name, params in metadata.joint_name_to_metadata.items():
            actuator_name = self.get_actuator_name(joint_name)
            if actuator_name not in ctrl_name_to_idx:
                if actuator_name != "root":
                    logger.warning("Joint %s has no actuator name. Skipping.", joint_name)
                continue
            idx = ctrl_name_to_idx[actuator_name]
            assert params.kp is not None and params.kd is not None, f"Missing kp or kd for joint {joint_name}"
            if params.soft_torque_limit is not None:
                if params.soft_torque_limit < 0:
                    raise ValueError(f"Soft torque limit for joint {joint_name} is negative: {params.soft_torque_limit}")
                clips[idx] = params.soft_torque_limit
            kps[idx], kds[idx] = params.kp, params.kd
        if any(kp == -1 for kp in kps):
            raise ValueError("Some KPs are not set. Check the provided metadata.")
        if any(kd == -1 for kd in kds):
            raise ValueError("Some KDs are not set. Check the provided metadata.")
        if any(k < 0 for k in kps) or any(k < 0 for k in kds):
            raise ValueError("Some KPs or KDs are negative. Check the provided metadata.")
        if any(k == 0 for k in kps) or any(k == 0 for k in kds):
            logger.warning("Some KPs or KDs are 0. Check the provided metadata.")
        self.kps, self.kds, self.ctrl_clip = kps, kds, clips
        self.action_noise = NoNoise() if action_noise is None else action_noise
        self.torque_noise = NoNoise() if torque_noise is None else torque_noise
        self.action_scale = action_scale
        self.action_bias, self.torque_bias = action_bias, torque_bias
        if isinstance(kp_scale, float):
            kp_scale = UniformRandomVariable(mean=kp_scale, mag=0.0)
        if isinstance(kd_scale, float):
            kd_scale = UniformRandomVariable(mean=kd_scale, mag=0.0)
        if isinstance(torque_limit_scale, float):
            torque_limit_scale = UniformRandomVariable(mean=torque_limit_scale, mag=0.0)
        self.kp_scale, self.kd_scale, self.torque_limit_scale = kp_scale, kd_scale, torque_limit_scale

    def get_actuator_name(self, joint_name: str) -> str:
        return f"{joint_name}_ctrl"

    def state_size(self, model: Model) -> int:
        return 2 * model.nu + 3

    def encode(self, model: Model) -> tuple[list[int], list[float]]:
        ani, anf = encode_noise(self.action_noise)
        tni, tnf = encode_noise(self.torque_noise)
        ints = ani + tni
        floats = [float(self.action_scale)] + anf + tnf
        for rv in (self.action_bias, self.torque_bias, self.kp_scale, self.kd_scale, self.torque_limit_scale):
            kind, f = encode_rv(rv)
            ints.append(kind)
            floats += f
        floats += [float(v) for v in self.kps] + [float(v) for v in self.kds] + [float(v) for v in self.ctrl_clip]
        return ints, floats


class PositionVelocityActuator(PositionActuators):
    """Position + velocity targets (ksim/actuators.py:241-293); action has 2*nu entries."""

    act_type = C.ACT_POSITION_VELOCITY

    def __init__(
        self,
        physics_model: Model,
        metadata: Metadata,
        pos_action_noise: Noise | None = None,
        vel_action_noise: Noise | None = None,
        torque_noise: Noise | None = None,
    ) -> None:
        super().__init__(physics_model=physics_model, metadata=metadata, action_noise=pos_action_noise, torque_noise=torque_noise)
        self.vel_action_noise = NoNoise() if vel_action_noise is None else vel_action_noise

    def action_size(self, model: Model) -> int:
        return 2 * model.nu

    def state_size(self, model: Model) -> int:
        return 0

    def encode(self, model: Model) -> tuple[list[int], list[float]]:
        """``get_ctrl`` (actuators.py:261-293): ``split(rng, 3)`` -> position, velocity and torque noise; ``ctrl = kps (noisy
        position target - q) + kds (noisy velocity target - qdot)``, torque noise, clip to ``ctrl_clip`` -- no action scale, no
        biases, no gain / limit scales (those belong to ``get_stateful_ctrl``).  Note on the reference: the class derives from
        the *stateful* ``PositionActuators``, so ``MjxEngine.step`` (engine.py:293-301) dispatches to the inherited
        ``get_stateful_ctrl`` with the 2 nu-long action and fails on the shape check; no reference example uses the class.  The
        engine here runs the class's own ``get_ctrl``.  Layout: the ``PositionActuators`` block (unused slots neutral) with the
        velocity noise appended."""
        ani, anf = encode_noise(self.action_noise)
        tni, tnf = encode_noise(self.torque_noise)
        vni, vnf = encode_noise(self.vel_action_noise)
        ints = ani + tni + [-1] * 5 + vni
        floats = [1.0] + anf + tnf + [0.0] * 15
        floats += [float(v) for v in self.kps] + [float(v) for v in self.kds] + [float(v) for v in self.ctrl_clip] + vnf
        return ints, floats
