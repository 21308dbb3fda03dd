"""Small hand-written MJCF models for tests (not derived from the reference's robots)."""

from __future__ import annotations

# A free-floating box with three hinged legs ending in sphere feet, three motors, a floor plane and an IMU site.
TRIPOD = """
<mujoco model="tripod">
  <option timestep="0.004" iterations="8" ls_iterations="8" tolerance="1e-10" integrator="implicitfast"/>
  <default>
    <joint damping="0.5" armature="0.01" limited="true" range="-60 60"/>
    <geom contype="0" conaffinity="0" condim="1" density="800"/>
    <motor ctrlrange="-1 1" ctrllimited="true"/>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="10 10 0.1"/>
    <body name="base" pos="0 0 0.5">
      <freejoint name="root"/>
      <geom name="base_geom" type="sphere" size="0.12"/>
      <site name="imu" pos="0 0 0.05"/>
      <body name="leg_a" pos="0.15 0 -0.05">
        <joint name="hip_a" axis="0 1 0"/>
        <geom name="leg_a_geom" type="capsule" fromto="0 0 0 0.1 0 -0.3" size="0.03"/>
        <body name="foot_a" pos="0.1 0 -0.3">
          <geom name="foot_a" type="sphere" size="0.04"/>
        </body>
      </body>
      <body name="leg_b" pos="-0.08 0.13 -0.05">
        <joint name="hip_b" axis="1 0 0"/>
        <geom name="leg_b_geom" type="capsule" fromto="0 0 0 -0.05 0.08 -0.3" size="0.03"/>
        <body name="foot_b" pos="-0.05 0.08 -0.3">
          <geom name="foot_b" type="sphere" size="0.04"/>
        </body>
      </body>
      <body name="leg_c" pos="-0.08 -0.13 -0.05">
        <joint name="hip_c" axis="1 0 0" stiffness="2"/>
        <geom name="leg_c_geom" type="capsule" fromto="0 0 0 -0.05 -0.08 -0.3" size="0.03"/>
        <body name="foot_c" pos="-0.05 -0.08 -0.3">
          <geom name="foot_c" type="sphere" size="0.04"/>
        </body>
      </body>
    </body>
  </worldbody>
  <contact>
    <pair geom1="foot_a" geom2="floor"/>
    <pair geom1="foot_b" geom2="floor"/>
    <pair geom1="foot_c" geom2="floor"/>
  </contact>
  <actuator>
    <motor name="hip_a_ctrl" joint="hip_a" gear="10"/>
    <motor name="hip_b_ctrl" joint="hip_b" gear="10"/>
    <motor name="hip_c_ctrl" joint="hip_c" gear="10"/>
  </actuator>
  <sensor>
    <gyro site="imu" name="imu_gyro"/>
    <accelerometer site="imu" name="imu_acc"/>
    <framequat objtype="site" objname="imu" name="orientation"/>
  </sensor>
</mujoco>
"""

# One body on a hinge about the world y axis: a physical pendulum.
PENDULUM = """
<mujoco model="pendulum">
  <option timestep="0.001" integrator="implicitfast"/>
  <worldbody>
    <body name="bob" pos="0 0 1">
      <joint name="pivot" type="hinge" axis="0 1 0" damping="0"/>
      <geom name="rod" type="sphere" size="0.05" pos="0 0 -0.5" density="1000" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
</mujoco>
"""

# A free sphere above a plane: free fall, then resting contact.
BALL = """
<mujoco model="ball">
  <option timestep="0.002" integrator="implicitfast"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" contype="0" conaffinity="0"/>
    <body name="ball" pos="0 0 1">
      <freejoint name="root"/>
      <geom name="ball" type="sphere" size="0.1" density="1000" contype="0" conaffinity="0" condim="1"/>
    </body>
  </worldbody>
  <contact>
    <pair geom1="ball" geom2="floor" condim="1"/>
  </contact>
</mujoco>
"""
