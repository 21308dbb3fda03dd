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
