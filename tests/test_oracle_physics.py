"""Analytic pins for the oracle's physics (no reference test pins any physics result -- SURVEY.md section 4):
closed forms and invariants that any faithful MuJoCo-style step must satisfy."""

import math

import numpy as np
import pytest

from ksim_b200.examples.walking import make_spec
from ksim_b200.mjcf import load_mjcf_string, reference_kinematics
from tests.models import BALL, PENDULUM, TRIPOD


@pytest.fixture(scope="module")
def oracle_mod(oracle_built: None):  # noqa: ANN001, ANN201
    import oracle

    return oracle


def test_humanoid_compile_sizes_and_mass() -> None:
    m = make_spec().model
    # nq 24, nv 23, nu 17, nbody 17, njnt 18, ngeom 18, nsite 3, 19 sensors / 54 floats
    assert (m.nq, m.nv, m.nu, m.nbody, m.njnt, m.ngeom, m.nsite, m.nsensor, m.nsensordata, m.npair) == (24, 23, 17, 17, 18, 18, 3, 19, 54, 2)
    # total mass = sum over geoms of density * volume (default density 1000)
    assert 35.0 < m.body_mass.sum() < 45.0
    # actuator order differs from joint order (abdomen_y before abdomen_z): SURVEY.md 9.1
    assert m.names["actuator"][:2] == ["abdomen_y_ctrl", "abdomen_z_ctrl"] and m.names["joint"][1:3] == ["abdomen_z", "abdomen_y"]


def test_free_fall_matches_closed_form(oracle_mod) -> None:  # noqa: ANN001
    model = load_mjcf_string(BALL)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    dt, g, n = model.opt.timestep, 9.81, 100
    for _ in range(n):
        d.step()
    # semi-implicit Euler: v_k = -g k dt, z_k = z0 - g dt^2 k (k + 1) / 2
    assert d.get("qvel", 6)[2] == pytest.approx(-g * n * dt, rel=1e-12)
    assert d.get("qpos", 7)[2] == pytest.approx(1.0 - g * dt * dt * n * (n + 1) / 2, rel=1e-12)
    assert int(d.get("nefc", 1)[0]) == 0


def test_ball_rests_on_plane(oracle_mod) -> None:  # noqa: ANN001
    model = load_mjcf_string(BALL)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    for _ in range(3000):
        d.step()
    z, vz = d.get("qpos", 7)[2], d.get("qvel", 6)[2]
    assert abs(vz) < 1e-4 and 0.09 < z < 0.1001  # small penetration allowed by the soft constraint
    # at rest the contact force balances gravity
    force = d.get("efc_force", 1)[0]
    assert force == pytest.approx(model.body_mass[1] * 9.81, rel=2e-3)


def test_pendulum_period(oracle_mod) -> None:  # noqa: ANN001
    model = load_mjcf_string(PENDULUM)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    theta0 = 0.05
    d.set("qpos", [theta0])
    mass = model.body_mass[1]
    inertia_about_pivot = model.body_inertia[1][1] + mass * 0.25  # point at 0.5 m
    period = 2 * math.pi * math.sqrt(inertia_about_pivot / (mass * 9.81 * 0.5))
    dt, crossings, prev, t = model.opt.timestep, [], theta0, 0.0
    for _ in range(int(3 * period / dt)):
        d.step()
        t += dt
        q = d.get("qpos", 1)[0]
        if prev > 0 >= q:
            crossings.append(t)
        prev = q
    assert len(crossings) >= 2
    assert crossings[1] - crossings[0] == pytest.approx(period, rel=2e-3)


def test_energy_is_conserved_without_damping(oracle_mod) -> None:  # noqa: ANN001
    model = load_mjcf_string(PENDULUM)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    d.set("qpos", [1.0])
    mass = model.body_mass[1]
    inertia = model.body_inertia[1][1] + mass * 0.25

    def energy() -> float:
        q, v = d.get("qpos", 1)[0], d.get("qvel", 1)[0]
        return 0.5 * inertia * v * v - mass * 9.81 * 0.5 * math.cos(q)

    e0 = energy()
    for _ in range(2000):
        d.step()
    assert abs(energy() - e0) < 2e-3 * abs(e0)


@pytest.mark.parametrize("xml", [TRIPOD, None])
def test_kinematics_and_mass_matrix_match_numpy_reference(oracle_mod, xml) -> None:  # noqa: ANN001
    model = load_mjcf_string(xml) if xml else make_spec().model
    rs = np.random.RandomState(0)
    qpos = model.qpos0.copy()
    qpos[7:] += 0.3 * rs.randn(model.nq - 7)
    quat = rs.randn(4)
    qpos[3:7] = quat / np.linalg.norm(quat)
    ref = reference_kinematics(model, qpos)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    d.set("qpos", qpos)
    d.kinematics()
    np.testing.assert_allclose(d.mat("xpos", model.nbody, 3), ref["xpos"], atol=1e-12)
    np.testing.assert_allclose(d.mat("xquat", model.nbody, 4), ref["xquat"], atol=1e-12)
    np.testing.assert_allclose(d.mat("cinert", model.nbody, 10), ref["cinert"], atol=1e-10)
    qm = d.mat("qM", model.nv, model.nv, 32)
    np.testing.assert_allclose(qm, ref["qM"], atol=1e-10)
    assert np.all(np.linalg.eigvalsh(qm) > 0)


def test_joint_limit_pushes_back(oracle_mod) -> None:  # noqa: ANN001
    model = load_mjcf_string(TRIPOD)
    om = oracle_mod.OracleModel(model, "f64")
    d = oracle_mod.OracleData(om)
    qpos = model.qpos0.copy()
    qpos[2] = 5.0  # far above the floor: no contacts
    qpos[7] = math.radians(65.0)  # beyond the +60 degree limit
    d.set("qpos", qpos)
    d.forward()
    assert int(d.get("nefc", 1)[0]) == 1
    assert d.get("efc_force", 1)[0] > 0 and d.get("qfrc_constraint", model.nv)[6] < 0


def test_f32_oracle_tracks_f64_over_one_control_step(oracle_mod) -> None:  # noqa: ANN001
    model = make_spec().model
    out = {}
    for precision in ("f32", "f64"):
        om = oracle_mod.OracleModel(model, precision)
        d = oracle_mod.OracleData(om)
        d.set("ctrl", 0.2 * np.sin(np.arange(model.nu)))
        for _ in range(5):
            d.step()
        out[precision] = d.get("qpos", model.nq)
    np.testing.assert_allclose(out["f32"], out["f64"], atol=2e-5)


# ------------------------------------------------------------------------------- more analytic pins (round 2)
def _incline_ball(theta: float, mu: float) -> str:
    """A sphere on a plane with gravity tilted by theta about y (== a plane inclined by theta), pyramidal friction mu."""
    g = 9.81
    return f"""
<mujoco model="incline">
  <option timestep="0.001" integrator="implicitfast" gravity="{g * math.sin(theta)} 0 {-g * math.cos(theta)}"/>
  <worldbody>
    <geom name="floor" type="plane" size="50 5 0.1" contype="0" conaffinity="0"/>
    <body name="ball" pos="0 0 0.0995">
      <freejoint name="root"/>
      <geom name="ball" type="sphere" size="0.1" density="1000" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <contact>
    <pair geom1="ball" geom2="floor" condim="3" friction="{mu} {mu} 0.005 0.0001 0.0001"/>
  </contact>
</mujoco>
"""


def _accel_along_x(oracle_mod, xml: str, settle: int = 400, span: int = 400) -> float:  # noqa: ANN001
    model = load_mjcf_string(xml)
    d = oracle_mod.OracleData(oracle_mod.OracleModel(model, "f64"))
    for _ in range(settle):
        d.step()
    v0 = d.get("qvel", 6)[0]
    for _ in range(span):
        d.step()
    return float((d.get("qvel", 6)[0] - v0) / (span * model.opt.timestep))


def test_sphere_rolls_without_slipping_down_an_incline(oracle_mod) -> None:  # noqa: ANN001
    """Enough friction: a solid sphere (I = 2/5 m r^2) rolls with a = 5/7 g sin(theta); the friction force needed is
    2/7 m g sin(theta), far inside the cone at mu = 1."""
    theta = math.radians(15.0)
    a = _accel_along_x(oracle_mod, _incline_ball(theta, 1.0))
    assert a == pytest.approx(5.0 / 7.0 * 9.81 * math.sin(theta), rel=0.02)


def test_sphere_slides_when_friction_is_too_small(oracle_mod) -> None:  # noqa: ANN001
    """mu < 2/7 tan(theta): the contact slips and the centre accelerates with g (sin(theta) - mu cos(theta)); the slip is along
    one axis of the friction pyramid, where the pyramidal and the elliptic cone agree."""
    theta, mu = math.radians(20.0), 0.05
    assert mu < 2.0 / 7.0 * math.tan(theta)
    a = _accel_along_x(oracle_mod, _incline_ball(theta, mu))
    assert a == pytest.approx(9.81 * (math.sin(theta) - mu * math.cos(theta)), rel=0.02)


def _loss_pendulum(frictionloss: float) -> str:
    return f"""
<mujoco model="loss_pendulum">
  <option timestep="0.001" integrator="implicitfast"/>
  <worldbody>
    <body name="bob" pos="0 0 1">
      <joint name="pivot" type="hinge" axis="0 1 0" damping="0" frictionloss="{frictionloss}"/>
      <geom name="rod" type="sphere" size="0.05" pos="0 0 -0.5" density="1000" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
</mujoco>
"""


def test_dry_joint_friction_holds_below_its_limit_and_slips_above(oracle_mod) -> None:  # noqa: ANN001
    """frictionloss is a force-limited SOFT constraint (MuJoCo's friction rows: quadratic zone, then constant force): a pendulum
    held horizontal with m g l < frictionloss is carried by a constraint force equal to the gravity torque and only creeps at a
    small terminal speed; above the limit the force saturates at frictionloss and the pendulum starts with angular acceleration
    (m g l - frictionloss) / I."""
    model = load_mjcf_string(_loss_pendulum(0.0))
    m = float(model.body_mass[1])
    gravity_torque = m * 9.81 * 0.5
    inertia = m * 0.5**2 + 0.4 * m * 0.05**2     # point mass at 0.5 m + the sphere's own 2/5 m r^2
    free_speed_after_20 = gravity_torque / inertia * 0.02
    for floss, holds in ((1.5 * gravity_torque, True), (0.4 * gravity_torque, False)):
        model = load_mjcf_string(_loss_pendulum(floss))
        d = oracle_mod.OracleData(oracle_mod.OracleModel(model, "f64"))
        d.set("qpos", [math.pi / 2])
        for _ in range(20):
            d.step()
        w20 = float(d.get("qvel", 1)[0])
        assert int(d.get("nefc", 1)[0]) == 1
        if holds:
            for _ in range(180):
                d.step()
            w200 = float(d.get("qvel", 1)[0])
            for _ in range(200):
                d.step()
            w400 = float(d.get("qvel", 1)[0])
            assert abs(w20) < 0.06 * free_speed_after_20 and abs(w400) < 0.03          # a creep, not a fall
            assert abs(w400 - w200) < 1e-5                                              # terminal speed reached
            assert abs(d.get("efc_force", 1)[0]) == pytest.approx(gravity_torque, rel=1e-3)
        else:
            assert abs(d.get("efc_force", 1)[0]) == pytest.approx(floss, rel=1e-9)      # saturated
            assert abs(w20) == pytest.approx((gravity_torque - floss) / inertia * 20 * model.opt.timestep, rel=0.01)


CAPSULE = """
<mujoco model="capsule">
  <option timestep="0.002" integrator="implicitfast"/>
  <worldbody>
    <geom name="floor" type="plane" size="5 5 0.1" contype="0" conaffinity="0"/>
    <body name="log" pos="0 0 0.3">
      <freejoint name="root"/>
      <geom name="log" type="capsule" fromto="-0.2 0 0 0.2 0 0" size="0.05" density="1000" contype="0" conaffinity="0"/>
    </body>
  </worldbody>
  <contact>
    <pair geom1="log" geom2="floor" condim="3"/>
  </contact>
</mujoco>
"""


def test_capsule_rests_on_its_two_end_contacts(oracle_mod) -> None:  # noqa: ANN001
    """A capsule lying on a plane: two contacts (one under each end of the segment), the axis stays horizontal at one radius
    (minus the soft penetration) and the normal forces carry the weight."""
    model = load_mjcf_string(CAPSULE)
    d = oracle_mod.OracleData(oracle_mod.OracleModel(model, "f64"))
    for _ in range(2500):
        d.step()
    q, v = d.get("qpos", 7), d.get("qvel", 6)
    assert np.abs(v).max() < 1e-3
    assert 0.049 < q[2] < 0.0501
    np.testing.assert_allclose(q[3:7] / np.linalg.norm(q[3:7]), [1, 0, 0, 0], atol=1e-4)
    assert int(d.get("ncon", 1)[0]) == 2
    ne = int(d.get("nefc", 1)[0])
    assert ne == 8   # two contacts x four pyramid edges
    f = d.get("efc_force", ne)
    # pyramid edge forces: the normal force of a contact is the sum of its four edge forces
    assert f.sum() == pytest.approx(float(model.body_mass[1]) * 9.81, rel=5e-3)
    assert f[:4].sum() == pytest.approx(f[4:].sum(), rel=1e-3)


def _momenta(model, d) -> tuple[np.ndarray, np.ndarray, np.ndarray]:  # noqa: ANN001
    """(total mass centre C, linear momentum P, angular momentum L about C) of the whole tree from the per-body quantities the
    oracle exposes: cvel = [omega, v at subtree_com[root]] (MuJoCo's convention), xipos / ximat, body mass / inertia."""
    nb = model.nbody
    mass = np.asarray(model.body_mass, np.float64)
    inertia = np.asarray(model.body_inertia, np.float64)
    r = d.get("xipos", 3 * nb).reshape(nb, 3)
    rot = d.get("ximat", 9 * nb).reshape(nb, 3, 3)
    cvel = d.get("cvel", 6 * nb).reshape(nb, 6)
    c0 = d.get("subtree_com", 3 * nb).reshape(nb, 3)[1]
    om, v = cvel[:, :3], cvel[:, 3:] + np.cross(cvel[:, :3], r - c0)
    com = (mass[:, None] * r).sum(0) / mass.sum()
    p = (mass[:, None] * v).sum(0)
    spin = np.einsum("bij,bj,bkj,bk->bi", rot, inertia, rot, om)
    ang = (mass[:, None] * np.cross(r - com, v)).sum(0) + spin.sum(0)
    return com, p, ang


@pytest.mark.parametrize("robot", ["humanoid", "kbot"])
def test_robot_in_free_flight_conserves_momentum(oracle_mod, robot) -> None:  # noqa: ANN001
    """The articulated robot (the humanoid; the K-Bot with its dry joint friction and armatures) thrown into the air with spinning limbs and its motors fighting each other: joint torques,
    damping, armature and joint limits are internal, so the linear momentum changes by M g t only and the angular momentum about
    the mass centre stays constant.  The semi-implicit Euler step is first-order accurate, so the test checks CONVERGENCE: the
    deviation from the conservation laws halves with the time step, and the Richardson extrapolation of the dt / 2 and dt / 4
    runs meets them to 1e-3.  Pins the composite-inertia mass matrix, the bias (Coriolis / gravity) forces, the damping and the
    actuator transmission of the real model against each other -- a wrong term would leave an error that does not vanish with dt."""
    import copy

    flight = 0.24
    res = {}
    for dt_scale in (1.0, 0.5, 0.25):
        if robot == "kbot":
            from ksim_b200.examples.kbot import make_spec as make_kbot

            model = copy.deepcopy(make_kbot().model)
        else:
            model = copy.deepcopy(make_spec().model)
        model.opt.timestep = model.opt.timestep * dt_scale
        d = oracle_mod.OracleData(oracle_mod.OracleModel(model, "f64"))
        rs = np.random.RandomState(0)
        qpos = np.asarray(model.qpos0, np.float64).copy()
        qpos[2] = 5.0                                   # far above the floor: no contact during the flight
        qpos[7:] += rs.uniform(-0.2, 0.2, model.nq - 7)
        qvel = np.concatenate([[0.4, -0.3, 2.0], rs.uniform(-2, 2, 3), rs.uniform(-4, 4, model.nv - 6)])
        d.set("qpos", qpos); d.set("qvel", qvel); d.set("ctrl", rs.uniform(-0.4, 0.4, model.nu))
        d.forward()
        _, p0, l0 = _momenta(model, d)
        n = int(round(flight / model.opt.timestep))
        for _ in range(n):
            d.step()
        assert int(d.get("ncon", 1)[0]) == 0 or d.get("qpos", 3)[2] > 2.0   # still in the air
        d.forward()
        _, p1, l1 = _momenta(model, d)
        total_mass = float(np.sum(model.body_mass))
        res[dt_scale] = (p1 - (p0 + total_mass * np.asarray(model.opt.gravity) * n * model.opt.timestep), l1 - l0)
    scale_p, scale_l = total_mass * 2.0, float(np.linalg.norm(l0))
    assert scale_l > 2.0
    for k, scale in ((0, scale_p), (1, scale_l)):
        e1, e2, e4 = (np.linalg.norm(res[s_][k]) / scale for s_ in (1.0, 0.5, 0.25))
        assert e1 < 0.06 and 0.35 < e2 / e1 < 0.65 and 0.35 < e4 / e2 < 0.65, (k, e1, e2, e4)
        extrapolated = np.linalg.norm(2.0 * res[0.25][k] - res[0.5][k]) / scale
        assert extrapolated < 1e-3, (k, extrapolated)
