"""The default humanoid walking task -- plugin lists of the reference's examples/walking.py:372-554, expressed
with this package's mirror classes.  BASELINE.json configs 1, 2 and 4 run this spec.
"""

from __future__ import annotations

__all__ = ["WalkingConfig", "load_humanoid", "make_spec", "RANDOMIZERS", "ZEROS", "ACTOR_INPUTS", "CRITIC_INPUTS", "network_columns", "make_model"]

import math
from dataclasses import dataclass, field
from pathlib import Path

from ksim_b200 import commands, events, observation, randomization, resets, rewards, terminations
from ksim_b200.actuators import PositionActuators
from ksim_b200.mjcf import Model, load_model
from ksim_b200.noise import AdditiveGaussianNoise, AdditiveUniformNoise, MultiplicativeUniformNoise
from ksim_b200.program import EngineConfig, TaskSpec
from ksim_b200.types import Metadata

_ASSET = Path(__file__).resolve().parent.parent / "assets" / "humanoid.npz"


@dataclass
class WalkingConfig:
    """examples/walking.py:255-340 + the ``__main__`` overrides at :772-788 (rollout-relevant fields only)."""

    num_envs: int = 4096
    rollout_length_frames: int = 24
    dt: float = 0.004
    ctrl_dt: float = 0.02
    zero_offset_std: float | None = math.radians(3.0)
    zero_offset_mag: float | None = math.radians(3.0)
    iterations: int = 8
    ls_iterations: int = 8
    tolerance: float = 1e-10
    action_latency_range: tuple[float, float] = (0.0, 0.0)
    drop_action_prob: float = 0.0
    action_clip_max: float = 1e3
    linear_velocity_range: tuple[float, float] = (1.0, 3.0)
    linear_velocity_max_yaw: float = math.pi / 4.0
    linear_velocity_accel: float = 0.5
    linear_velocity_zero_prob: float = 0.2
    linear_velocity_switch_prob: float = 0.001
    angular_velocity_accel: float = 0.5
    angular_velocity_range: tuple[float, float] = (0.0, 0.2)
    angular_velocity_zero_prob: float = 0.2
    angular_velocity_switch_prob: float = 0.001
    gait_period: float = 0.6
    max_foot_height: float = 0.2
    # PPO (ksim/task/ppo.py:251-306 defaults)
    gamma: float = 0.97
    lam: float = 0.95
    extra: dict = field(default_factory=dict)


def load_humanoid(config: WalkingConfig) -> Model:
    """``get_mujoco_model`` + ``set_mujoco_model_opts`` (examples/walking.py:389-391, ksim/task/rl.py:717-748)."""
    model = load_model(_ASSET)
    o = model.opt
    o.timestep, o.iterations, o.ls_iterations, o.tolerance = config.dt, config.iterations, config.ls_iterations, config.tolerance
    o.integrator, o.solver, o.disable_eulerdamp = "implicitfast", "newton", True
    return model


RANDOMIZERS = {
    "static_friction": lambda m: randomization.StaticFrictionRandomizer(),
    "armature": lambda m: randomization.ArmatureRandomizer(),
    "mass_multiplication": lambda m: randomization.MassMultiplicationRandomizer.from_body_name(m, "torso"),
    "joint_damping": lambda m: randomization.JointDampingRandomizer(),
}


def make_spec(config: WalkingConfig | None = None) -> TaskSpec:
    config = WalkingConfig() if config is None else config
    model = load_humanoid(config)
    metadata = Metadata.from_model(model, kp=50.0, kd=1.0)
    eng = EngineConfig(
        dt=config.dt, ctrl_dt=config.ctrl_dt, action_latency_range=config.action_latency_range,
        drop_action_prob=config.drop_action_prob, zero_offset_std=config.zero_offset_std,
        zero_offset_mag=config.zero_offset_mag, action_clip_max=config.action_clip_max,
    )
    obs = {
        "joint_position": observation.JointPositionObservation(noise=AdditiveUniformNoise(mag=math.radians(2.0))),
        "joint_velocity": observation.JointVelocityObservation(noise=MultiplicativeUniformNoise(mag=math.radians(10.0))),
        "actuator_force": observation.ActuatorForceObservation(),
        "center_of_mass_inertia": observation.CenterOfMassInertiaObservation(),
        "center_of_mass_velocity": observation.CenterOfMassVelocityObservation(),
        "base_position": observation.BasePositionObservation(),
        "base_orientation": observation.BaseOrientationObservation(),
        "base_linear_velocity": observation.BaseLinearVelocityObservation(),
        "base_angular_velocity": observation.BaseAngularVelocityObservation(),
        "imu_projected_gravity": observation.ProjectedGravityObservation.create(
            physics_model=model, framequat_name="orientation", noise=AdditiveGaussianNoise(std=0.01),
            min_lag=0.001, max_lag=0.005, imu_bias=math.radians(2.0)),
        "projected_gravity": observation.ProjectedGravityObservation.create(physics_model=model, framequat_name="orientation"),
        "imu_acc": observation.SensorObservation.create(physics_model=model, sensor_name="imu_acc", noise=AdditiveGaussianNoise(std=0.01)),
        "imu_gyro": observation.SensorObservation.create(physics_model=model, sensor_name="imu_gyro", noise=AdditiveGaussianNoise(std=0.01)),
        "feet_contact": observation.FeetContactObservation.create(
            physics_model=model, foot_left_geom_names=["foot_left"], foot_right_geom_names=["foot_right"], floor_geom_names=["floor"]),
        "feet_position": observation.BodyPositionObservation.create(physics_model=model, body_names=("foot_left", "foot_right")),
        "feet_force": observation.FeetForceObservation.create(physics_model=model, foot_left_body_name="foot_left", foot_right_body_name="foot_right"),
    }
    cmds = {
        "linvel": commands.LinearVelocityCommand(
            min_vel=config.linear_velocity_range[0], max_vel=config.linear_velocity_range[1],
            max_yaw=config.linear_velocity_max_yaw, zero_prob=config.linear_velocity_zero_prob,
            switch_prob=config.linear_velocity_switch_prob, ctrl_dt=config.ctrl_dt,
            linear_accel=config.linear_velocity_accel, angular_accel=config.angular_velocity_accel),
        "angvel": commands.AngularVelocityCommand(
            min_vel=config.angular_velocity_range[0], max_vel=config.angular_velocity_range[1],
            zero_prob=config.angular_velocity_zero_prob, switch_prob=config.angular_velocity_switch_prob,
            ctrl_dt=config.ctrl_dt, angular_accel=config.angular_velocity_accel),
    }
    rews = {
        "stay_alive": rewards.StayAliveReward(scale=100.0),
        "upright": rewards.UprightReward(scale=5.0),
        "linvel": rewards.LinearVelocityReward(cmd="linvel", scale=0.1),
        "angvel": rewards.AngularVelocityReward(cmd="angvel", scale=0.01),
        "foot_airtime": rewards.FeetAirTimeReward(ctrl_dt=config.ctrl_dt, gait_period=config.gait_period, air_time_percent=0.4,
                                                  contact_obs="feet_contact", scale=1.0),
        "foot_height": rewards.SparseTargetHeightReward(contact_obs="feet_contact", position_obs="feet_position",
                                                        height=config.max_foot_height, scale=1.0),
        "foot_force": rewards.ForcePenalty(force_obs="feet_force", ctrl_dt=config.ctrl_dt, expected_weight=10.0, scale=1e-1),
        "motionless_at_rest": rewards.MotionlessAtRestPenalty(scale=1e-2),
    }
    terms = {
        "bad_z": terminations.BadZTermination(min_z=0.5, max_z=3.0),
        "bad_velocity": terminations.BadVelocityTermination(max_vel=100.0),
        "far_from_origin": terminations.FarFromOriginTermination(max_dist=10.0),
    }
    return TaskSpec(
        model=model, config=eng,
        actuators=PositionActuators(physics_model=model, metadata=metadata),
        resets=[
            resets.RandomJointPositionReset.create(model, zeros={"abdomen_z": 0.0}),
            resets.RandomJointVelocityReset(),
            resets.RandomHeadingReset(),
        ],
        events={
            "push": events.LinearPushEvent(linvel=1.0, interval_range=(2.0, 5.0)),
            "jump": events.JumpEvent(jump_height_range=(0.1, 0.5), interval_range=(2.0, 5.0)),
        },
        observations=obs, commands=cmds, rewards=rews, terminations=terms,
        randomized_fields=("dof_frictionloss", "dof_armature", "body_mass", "dof_damping"),
    )


# examples/walking.py:19-37: per-joint bias added to the mixture means (all zero for the humanoid), in actuator order
ZEROS = ("abdomen_z", "abdomen_y", "abdomen_x", "hip_x_right", "hip_z_right", "hip_y_right", "knee_right", "hip_x_left", "hip_z_left",
         "hip_y_left", "knee_left", "shoulder1_right", "shoulder2_right", "elbow_right", "shoulder1_left", "shoulder2_left", "elbow_left")

# run_actor (walking.py:579-614) and run_critic (:616-672): (observation name, scale) in concatenation order, then the command
# columns (LinearVelocityCommandValue fields vel, yaw, xvel, yvel = columns 0..3 of "linvel"; AngularVelocityCommandValue.vel = 0)
ACTOR_INPUTS = ((("noisy_joint_position", 1.0), ("noisy_joint_velocity", 0.1), ("noisy_imu_projected_gravity", 1.0), ("noisy_imu_gyro", 1.0)),
                (("linvel", 0), ("linvel", 1), ("angvel", 0)))
CRITIC_INPUTS = ((("joint_position", 1.0), ("joint_velocity", 0.1), ("center_of_mass_inertia", 1.0), ("center_of_mass_velocity", 1.0),
                  ("imu_acc", 1.0), ("imu_gyro", 1.0), ("projected_gravity", 1.0), ("actuator_force", 0.01), ("base_position", 1.0),
                  ("base_orientation", 1.0), ("base_linear_velocity", 1.0), ("base_angular_velocity", 1.0), ("feet_force", 1.0)),
                 (("linvel", 0), ("linvel", 1), ("linvel", 2), ("linvel", 3), ("angvel", 0)))


def network_columns(program, inputs) -> tuple[list[int], list[float]]:  # noqa: ANN001
    """Record columns and scales that assemble a network input row from a pre-step record block: the gather ``run_actor`` /
    ``run_critic`` do by name on the observation / command dicts."""
    ob, cm = program["TI_R_OBS"], program["TI_R_CMD"]
    idx: list[int] = []
    scale: list[float] = []
    for name, sc in inputs[0]:
        off, dim, _ = program.obs_slices[name]
        idx += list(range(ob + off, ob + off + dim))
        scale += [sc] * dim
    for name, col in inputs[1]:
        off, dim = program.cmd_slices[name]
        if not 0 <= col < dim:
            raise ValueError(f"command {name} has {dim} columns")
        idx.append(cm + off + col)
        scale.append(1.0)
    return idx, scale


def make_model(spec: TaskSpec, hidden_size: int = 512, depth: int = 2, num_hidden_layers: int = 2, num_mixtures: int = 10):  # noqa: ANN201
    """``WalkingTask.get_model`` (walking.py:556-571)."""
    from ksim_b200.policy import Model

    rng = spec.model.jnt_range[1:]  # ksim.ClipPositions.from_physics_model (actions.py:196-200): every joint after the root
    return Model(num_actor_inputs=43, num_critic_inputs=340, num_joints=len(ZEROS), hidden_size=hidden_size, depth=depth,
                 num_hidden_layers=num_hidden_layers, num_mixtures=num_mixtures, min_std=0.01, max_std=1.0, var_scale=0.5,
                 mean_bias=[0.0] * len(ZEROS), clip_lo=rng[:, 0], clip_hi=rng[:, 1])
