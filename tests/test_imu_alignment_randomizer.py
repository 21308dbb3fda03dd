"""``IMUAlignmentRandomizer`` (ksim/randomization.py:245-283): per-environment site frame (``site_quat`` / ``site_pos`` overrides).

CPU: properties of the host class against the reference's expressions (unit quaternions, small-angle composition, reproducible
from the key); the per-env frame reaches the site-frame sensors of the oracle (framequat of a resting robot == the randomised
site quaternion composed with the base's); host emulation == oracle.  GPU: CUDA == oracle, and the task is NOT served by the
humanoid-sized instantiation, which has no per-env geometry."""

import math

import numpy as np
import pytest

from ksim_b200 import observation, randomization, rng as R
from ksim_b200.envinit import make_env_init
from ksim_b200.program import build_program
from tests.helpers import tripod_model, tripod_spec
from tests.test_emu_vs_oracle import _p, emu  # noqa: F401

RZ = {"imu": lambda m: randomization.IMUAlignmentRandomizer(site_name="imu", tilt_std_rad=math.radians(4), yaw_std_rad=math.radians(3),
                                                           translate_std_m=0.01)}


def _spec():  # noqa: ANN202
    model = tripod_model()
    obs = {"quat": observation.SensorObservation.create(physics_model=model, sensor_name="orientation"),
           "gyro": observation.SensorObservation.create(physics_model=model, sensor_name="imu_gyro"),
           "acc": observation.SensorObservation.create(physics_model=model, sensor_name="imu_acc"),
           "site": observation.SiteOrientationObservation.create(physics_model=model, site_names=("imu",)),
           "base_orientation": observation.BaseOrientationObservation()}
    spec = tripod_spec(observations=obs)
    spec.randomized_fields = ("site_quat", "site_pos")
    return spec


def _qmul(a, b):  # noqa: ANN001, ANN202
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], axis=-1)


def test_host_class_matches_the_reference_expressions() -> None:
    model = tripod_model()
    rz = RZ["imu"](model)
    keys = R.split(R.prng_key(3), 64)
    out = rz(model, keys)
    q, p = out["site_quat"][:, 0].astype(np.float64), out["site_pos"][:, 0].astype(np.float64)
    np.testing.assert_allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-6)
    # transcription of randomization.py:252-281 for one key, float64 arithmetic on the same float32 normal draws
    for e in (0, 17, 63):
        rng, sub = R.split(keys[e], 2)
        rx, ry = (R.normal(sub, 2) * np.float32(math.radians(4))).astype(np.float64)
        rng, sub = R.split(rng, 2)
        rzz = float(R.normal(sub, 1)[0] * np.float32(math.radians(3)))
        qo = _qmul(np.array([1, 0, 0, rzz * 0.5]), _qmul(np.array([1, 0, ry * 0.5, 0]), np.array([1, rx * 0.5, 0, 0])))
        qo /= np.linalg.norm(qo)
        np.testing.assert_allclose(q[e], _qmul(qo, model.site_quat[0]), atol=2e-7)
        rng, sub = R.split(rng, 2)
        np.testing.assert_allclose(p[e], model.site_pos[0] + (R.normal(sub, 3) * np.float32(0.01)).astype(np.float64), atol=1e-7)
    ang = 2 * np.arccos(np.clip(np.abs(q[:, 0]), 0, 1))
    assert 0.005 < ang.mean() < 0.2 and np.unique(q[:, 1]).size == 64
    assert set(randomization.IMUAlignmentRandomizer(site_name="imu")(model, keys)) == {"site_quat"}   # no translation by default


def _oracle(spec, n, t, seed=9):  # noqa: ANN001, ANN202
    from oracle import OracleEngine

    prog = build_program(spec)
    rz = {k: f(spec.model) for k, f in RZ.items()}
    init = make_env_init(prog, rz, n, seed=seed, level=0.0)
    ora = OracleEngine(prog, "f32")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    actions = (0.3 * np.random.RandomState(seed).randn(t, n, prog.action_size)).astype(np.float32)
    ora.rollout(actions, init.rand, st.copy(), keys.copy(), rec)
    return prog, init, st, keys, rec, actions, rz


def _obs(prog, rec, name):  # noqa: ANN001, ANN202
    o, d, _ = prog.obs_slices[name]
    return rec[..., prog["TI_R_OBS"] + o : prog["TI_R_OBS"] + o + d]


def test_site_frame_reaches_the_oracle_sensors(oracle_built) -> None:  # noqa: ANN001
    spec = _spec()
    n, t = 16, 6
    prog, init, _, _, rec, _, _ = _oracle(spec, n, t)
    o, d = prog.rand_slices["site_quat"]
    site_q = init.rand[:, o : o + d].reshape(n, spec.model.nsite, 4)[:, 0].astype(np.float64)
    assert np.unique(site_q[:, 1]).size == n
    # framequat(site) == base quaternion * per-env site quaternion; level 0: the reset leaves the base quaternion at its default
    base_q = _obs(prog, rec, "base_orientation")[0].astype(np.float64)
    base_q /= np.linalg.norm(base_q, axis=-1, keepdims=True)
    np.testing.assert_allclose(_obs(prog, rec, "quat")[0], _qmul(base_q, site_q), atol=2e-6)
    # and the site's rotation matrix rows (SiteOrientationObservation) are that quaternion's
    q = _qmul(base_q, site_q)
    w_, x, y, z = (q[..., i] for i in range(4))
    rows = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y),
                     2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)], axis=-1)
    np.testing.assert_allclose(_obs(prog, rec, "site")[0], rows, atol=3e-6)


def test_emu_matches_oracle(emu, oracle_built) -> None:  # noqa: ANN001, F811
    from ksim_b200.blob import pack_blob

    spec = _spec()
    n, t = 8, 6
    prog, init, st, keys, rec_o, actions, _ = _oracle(spec, n, t)
    blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
    st_e, k_e = np.zeros_like(st), np.zeros_like(keys)
    rec_e = np.zeros_like(rec_o)
    ak = np.zeros((n, 2), np.uint32)
    emu.emu_init_envs(_p(blob), n, _p(init.init_keys), _p(init.levels), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[0]), _p(ak))
    for s in range(t):
        emu.emu_step_ctrl(_p(blob), n, _p(actions[s]), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[s]), _p(rec_e[s + 1]), None)
    ob = prog["TI_R_OBS"]
    np.testing.assert_allclose(rec_e[:, :, ob : ob + prog.obs_size], rec_o[:, :, ob : ob + prog.obs_size], rtol=0, atol=2e-3)
    np.testing.assert_allclose(_obs(prog, rec_e, "quat"), _obs(prog, rec_o, "quat"), rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    spec = _spec()
    n, t = 40, 8
    prog, init, st, keys, rec_o, actions, rz = _oracle(spec, n, t)
    eng = B200Engine(spec, n, device="cuda:0", randomizers=rz, seed=9, curriculum_level=0.0)
    assert eng.info("VARIANT") >= 1
    np.testing.assert_array_equal(eng.rand.cpu().numpy(), init.rand)
    rec = eng.rollout(torch.from_numpy(actions).cuda()).cpu().numpy()
    np.testing.assert_allclose(_obs(prog, rec, "quat")[:3], _obs(prog, rec_o, "quat")[:3], rtol=0, atol=2e-5)
    np.testing.assert_allclose(_obs(prog, rec, "site")[:3], _obs(prog, rec_o, "site")[:3], rtol=0, atol=2e-5)
    ob = prog["TI_R_OBS"]
    ok = np.abs(rec[:, :, ob : ob + prog.obs_size] - rec_o[:, :, ob : ob + prog.obs_size]).max(axis=(0, 2)) < 5e-2
    assert ok.mean() >= 0.9, f"{(~ok).sum()} of {n} envs drift"


@pytest.mark.gpu
def test_humanoid_with_per_env_geometry_leaves_the_small_instantiation(walking_spec) -> None:  # noqa: ANN001
    """A humanoid task that randomises site frames must be served by an instantiation that carries them (not variant 0)."""
    import copy

    from ksim_b200.engine import B200Engine

    spec = copy.copy(walking_spec)
    spec.randomized_fields = tuple(spec.randomized_fields) + ("site_quat",)
    rz = {k: f(spec.model) for k, f in __import__("ksim_b200.examples.walking", fromlist=["RANDOMIZERS"]).RANDOMIZERS.items()}
    rz["imu"] = randomization.IMUAlignmentRandomizer(site_name=spec.model.names["site"][0])
    eng = B200Engine(spec, 8, device="cuda:0", randomizers=rz, seed=0)
    assert eng.info("VARIANT") == 2
