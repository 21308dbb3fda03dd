"""Config-time metadata types.  Mirrors the parts of ksim/types.py:129-266 the hot path needs."""

from __future__ import annotations

__all__ = ["JointMetadata", "ActuatorMetadata", "Metadata"]

from dataclasses import dataclass, field

from ksim_b200.mjcf import Model


@dataclass
class JointMetadata:
    kp: float | None = None
    kd: float | None = None
    armature: float | None = None
    friction: float | None = None
    actuator_type: str | None = None
    soft_torque_limit: float | None = None

    @classmethod
    def from_model(cls, model: Model, **kwargs: float | str | None) -> dict[str, "JointMetadata"]:
        return {name: cls(**kwargs) for name in model.names["joint"]}  # type: ignore[arg-type]


@dataclass
class ActuatorMetadata:
    actuator_type: str | None = None

    @classmethod
    def from_model(cls, model: Model) -> dict[str, "ActuatorMetadata"]:
        return {"motor": cls(actuator_type="motor")}


@dataclass
class Metadata:
    joint_name_to_metadata: dict[str, JointMetadata] = field(default_factory=dict)
    actuator_type_to_metadata: dict[str, ActuatorMetadata] = field(default_factory=dict)
    control_frequency: float | None = None

    @classmethod
    def from_model(
        cls,
        model: Model,
        kp: float | None = None,
        kd: float | None = None,
        armature: float | None = None,
        friction: float | None = None,
        soft_torque_limit: float | None = None,
    ) -> "Metadata":
        return cls(
            joint_name_to_metadata=JointMetadata.from_model(
                model, kp=kp, kd=kd, armature=armature, friction=friction, soft_torque_limit=soft_torque_limit
            ),
            actuator_type_to_metadata=ActuatorMetadata.from_model(model),
        )

    @classmethod
    def from_json(cls, obj: dict) -> "Metadata":
        """Reads the K-Scale ``metadata.json`` layout (examples/kbot/robot/kbot-headless/metadata.json)."""
        def _f(v: object) -> float | None:
            return None if v is None else float(v)  # type: ignore[arg-type]

        joints = {}
        for name, jm in (obj.get("joint_name_to_metadata") or {}).items():
            joints[name] = JointMetadata(
                kp=_f(jm.get("kp")), kd=_f(jm.get("kd")), armature=_f(jm.get("armature")),
                friction=_f(jm.get("friction")), actuator_type=jm.get("actuator_type"),
                soft_torque_limit=_f(jm.get("soft_torque_limit")),
            )
        acts = {k: ActuatorMetadata(actuator_type=v.get("actuator_type")) for k, v in (obj.get("actuator_type_to_metadata") or {}).items()}
        return cls(joint_name_to_metadata=joints, actuator_type_to_metadata=acts, control_frequency=_f(obj.get("control_frequency")))
