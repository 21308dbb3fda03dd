"""The observation classes added in round 2 -- ``DelayedJointPositionObservation`` / ``DelayedJointVelocityObservation``
(ksim/observation.py:229-283), ``ContactObservation`` (452-478), ``BodyOrientationObservation`` / ``FeetOrientationObservation``
(619-703), ``SiteOrientationObservation`` (645-669), ``ActPosObservation`` (717-747).

CPU: the oracle's readings against a numpy restatement of the reference's expressions evaluated on the trajectory's own recorded
fields (the delayed readings are earlier undelayed readings; after a reset the ring holds the reset state); the host emulation of
the kernel code against the oracle.  GPU: CUDA == oracle through the fused control step (generic instantiation).
"""

import numpy as np
import pytest

from ksim_b200 import observation, terminations
from ksim_b200.envinit import make_env_init
from ksim_b200.program import build_program
from tests.helpers import tripod_model, tripod_spec
from tests.test_emu_vs_oracle import _p, emu  # noqa: F401  (the host-emulation library fixture)

DELAY_P, DELAY_V = 3, 2


def _spec():  # noqa: ANN202
    model = tripod_model()
    obs = {
        "joint_position": observation.JointPositionObservation(),
        "joint_velocity": observation.JointVelocityObservation(),
        "delayed_joint_position": observation.DelayedJointPositionObservation(delay_steps=DELAY_P),
        "delayed_joint_velocity": observation.DelayedJointVelocityObservation(delay_steps=DELAY_V),
        "contact": observation.ContactObservation.create(physics_model=model, geom_names=["foot_a", "foot_b", "floor"]),
        "body_orientation": observation.BodyOrientationObservation.create(physics_model=model, body_names=["base", "leg_b"]),
        "feet_orientation": observation.FeetOrientationObservation.create_from_feet(
            physics_model=model, foot_left_body_name="foot_a", foot_right_body_name="foot_b"),
        "site_orientation": observation.SiteOrientationObservation.create(physics_model=model, site_names=("imu",)),
        "act_pos": observation.ActPosObservation.create(physics_model=model, joint_name="hip_b"),
        "base_orientation": observation.BaseOrientationObservation(),
        "feet_contact": observation.FeetContactObservation.create(
            physics_model=model, foot_left_geom_names=["foot_a"], foot_right_geom_names=["foot_b"], floor_geom_names=["floor"]),
    }
    # short episodes so that the rings are re-initialised by reset-on-done inside the rollout
    term = {"episode_length": terminations.EpisodeLengthTermination(max_length_sec=0.17)}
    return tripod_spec(observations=obs, terminations=term, zero_offset_std=0.05)


def _oracle_rollout(spec, n: int, t: int, seed: int = 4):  # noqa: ANN001, ANN202
    from oracle import OracleEngine

    prog = build_program(spec)
    init = make_env_init(prog, {}, n, seed=seed, level=1.0)
    ora = OracleEngine(prog, "f32")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    actions = (0.5 * np.random.RandomState(seed + 1).randn(t, n, prog.action_size)).astype(np.float32)
    ora.rollout(actions, init.rand, st.copy(), keys.copy(), rec)
    return prog, init, ora, st, keys, rec, actions


def _obs(prog, rec: np.ndarray, name: str) -> np.ndarray:  # noqa: ANN001
    off, dim, _ = prog.obs_slices[name]
    o = prog["TI_R_OBS"]
    return rec[..., o + off : o + off + dim]


def test_shapes_and_names() -> None:
    spec = _spec()
    prog = build_program(spec)
    m = spec.model
    shape = {k: v[2] for k, v in prog.obs_slices.items()}
    assert shape["delayed_joint_position"] == (m.nq - 7,)
    assert shape["body_orientation"] == (2, 4) and shape["site_orientation"] == (1, 6)
    assert shape["contact"] == (3,) and shape["act_pos"] == (2,)
    assert prog["TI_OBS_CARRY_SIZE"] == (DELAY_P + DELAY_V) * (m.nq - 7)
    assert spec.observations["feet_orientation"].name == "feet_orientation_observation_foot_a_foot_b"
    with pytest.raises(ValueError):
        observation.FeetOrientationObservation.create(physics_model=m, body_names=["base"])


def test_oracle_matches_numpy_restatement(oracle_built) -> None:  # noqa: ANN001
    spec = _spec()
    n, t = 6, 30
    prog, _, _, _, _, rec, actions = _oracle_rollout(spec, n, t)
    m = spec.model
    done = rec[:t, :, prog["TI_R_DONE"]] != 0
    assert done.any() and not done.all()
    # slot s holds the observation the policy sees BEFORE step s; slot s + 1 was written after step s (after its reset, if any)
    jp, jv = _obs(prog, rec, "joint_position"), _obs(prog, rec, "joint_velocity")
    dp, dv = _obs(prog, rec, "delayed_joint_position"), _obs(prog, rec, "delayed_joint_velocity")
    for e in range(n):
        ring_p = ring_v = None
        for s in range(t + 1):
            fresh = s == 0 or done[s - 1, e]
            if fresh:   # initial_carry: every row = the state right after the reset (its undelayed reading in this slot)
                ring_p, ring_v = [jp[s, e].copy()] * DELAY_P, [jv[s, e].copy()] * DELAY_V
            np.testing.assert_allclose(dp[s, e], ring_p[0], rtol=0, atol=1e-6)
            np.testing.assert_allclose(dv[s, e], ring_v[0], rtol=0, atol=1e-6)
            ring_p, ring_v = ring_p[1:] + [jp[s, e].copy()], ring_v[1:] + [jv[s, e].copy()]
    # orientation observations against the recorded post-step fields of the previous slot (no reset in between)
    xq = rec[:t, :, prog["TI_R_XQUAT"] : prog["TI_R_XQUAT"] + 4 * m.nbody].reshape(t, n, m.nbody, 4)
    keep = ~done
    bo = _obs(prog, rec, "body_orientation")[1:].reshape(t, n, 2, 4)
    ids = [m.id("body", "base"), m.id("body", "leg_b")]
    np.testing.assert_allclose(bo[keep], xq[:, :, ids][keep], rtol=0, atol=1e-6)
    fo = _obs(prog, rec, "feet_orientation")[1:].reshape(t, n, 2, 4)
    ids = [m.id("body", "foot_a"), m.id("body", "foot_b")]
    np.testing.assert_allclose(fo[keep], xq[:, :, ids][keep], rtol=0, atol=1e-6)
    # the imu site is fixed to the base with the identity orientation: its matrix is the base quaternion's, rows 0 and 1
    q = xq[:, :, m.id("body", "base")]
    w_, x, y, z = (q[..., i] for i in range(4))
    rows = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w_ * z), 2 * (x * z + w_ * y),
                     2 * (x * y + w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w_ * x)], axis=-1)
    np.testing.assert_allclose(_obs(prog, rec, "site_orientation")[1:][keep], rows[keep], rtol=0, atol=2e-6)
    # ContactObservation: foot i touches something listed <=> it touches the floor (the feet never touch each other here);
    # the floor's own entry is the `any` over both feet
    fc = _obs(prog, rec, "feet_contact")   # (1, 2): [foot_a on floor, foot_b on floor]
    co = _obs(prog, rec, "contact")
    np.testing.assert_array_equal(co[..., 0], fc[..., 0])
    np.testing.assert_array_equal(co[..., 1], fc[..., 1])
    np.testing.assert_array_equal(co[..., 2], np.maximum(fc[..., 0], fc[..., 1]))
    assert 0 < co[..., :2].mean() < 1
    # ActPosObservation: (last_ctrl[ctrl_idx], qpos[qpos_idx]); last_ctrl is the previous step's action (0 after a reset)
    ap = _obs(prog, rec, "act_pos")
    ci, qi = m.id("actuator", "hip_b_ctrl"), int(m.jnt_qposadr[m.id("joint", "hip_b")])
    want = np.where(done, 0.0, actions[:, :, ci])
    np.testing.assert_allclose(ap[1:, :, 0], want, rtol=0, atol=1e-6)
    qp = rec[:t, :, prog["TI_R_QPOS"] + qi]
    np.testing.assert_allclose(ap[1:, :, 1][keep], qp[keep], rtol=0, atol=1e-6)


def test_emu_matches_oracle(emu, oracle_built) -> None:  # noqa: ANN001, F811
    from ksim_b200.blob import pack_blob

    lib = emu
    spec = _spec()
    n, t = 4, 16
    prog, init, ora, st, keys, rec_o, actions = _oracle_rollout(spec, n, t)
    blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
    st_e, k_e = np.zeros_like(st), np.zeros_like(keys)
    rec_e = np.zeros_like(rec_o)
    ak = np.zeros((n, 2), np.uint32)
    lib.emu_init_envs(_p(blob), n, _p(init.init_keys), _p(init.levels), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[0]), _p(ak))
    np.testing.assert_allclose(st_e, st, rtol=0, atol=2e-5)
    for s in range(t):
        lib.emu_step_ctrl(_p(blob), n, _p(actions[s]), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[s]), _p(rec_e[s + 1]), None)
    o = prog["TI_R_OBS"]
    np.testing.assert_array_equal(rec_e[:t, :, prog["TI_R_DONE"]], rec_o[:t, :, prog["TI_R_DONE"]])
    np.testing.assert_allclose(rec_e[:, :, o : o + prog.obs_size], rec_o[:, :, o : o + prog.obs_size], rtol=0, atol=2e-3)


@pytest.mark.gpu
def test_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    spec = _spec()
    n, t = 40, 24
    prog, init, ora, st, keys, rec_o, actions = _oracle_rollout(spec, n, t, seed=4)
    eng = B200Engine(spec, n, device="cuda:0", seed=4, curriculum_level=1.0)
    assert eng.info("VARIANT") == 2   # the ring needs the generic instantiation's observation-carry budget
    np.testing.assert_array_equal(eng.keys.cpu().numpy().view(np.uint32), keys)
    np.testing.assert_allclose(eng.state.cpu().numpy(), st, rtol=0, atol=2e-5)
    rec = eng.rollout(torch.from_numpy(actions).cuda()).cpu().numpy()
    np.testing.assert_array_equal(rec[:t, :, prog["TI_R_DONE"]], rec_o[:t, :, prog["TI_R_DONE"]])
    o = prog["TI_R_OBS"]
    got, want = rec[:, :, o : o + prog.obs_size], rec_o[:, :, o : o + prog.obs_size]
    # chained rollout: an env whose foot contact flips on an ulp drops out from then on (counted, must stay rare)
    ok = np.ones(n, bool)
    for s in range(t + 1):
        ok &= np.abs(got[s] - want[s]).max(axis=1) < 5e-3
    assert ok.mean() >= 0.9, f"{(~ok).sum()} of {n} envs drift"
    for name in ("delayed_joint_position", "delayed_joint_velocity", "body_orientation", "site_orientation", "act_pos", "contact"):
        np.testing.assert_allclose(_obs(prog, rec, name)[:, ok], _obs(prog, rec_o, name)[:, ok], rtol=0, atol=5e-3, err_msg=name)
