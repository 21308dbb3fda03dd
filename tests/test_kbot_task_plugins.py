"""The K-Bot task's own plugins (examples/kbot/train.py:128-834): the oracle's C restatement against an independent
vectorised numpy restatement written from the reference's JAX code (padding, scans, lexsort and all), and the warp code
(host emulation) against the oracle.  The -m gpu counterpart is tests/test_gpu_parity.py::test_kbot_*."""

import ctypes
import math

import numpy as np
import pytest

from ksim_b200 import rng as R
from ksim_b200.program import build_program

F = np.float64


# ----------------------------------------------------------------------------------------- numpy restatements


def quat_to_euler(q: np.ndarray) -> np.ndarray:
    q = q / (np.linalg.norm(q, axis=-1, keepdims=True) + 1e-6)
    w, x, y, z = np.moveaxis(q, -1, 0)
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sinp = 2 * (w * y - z * x)
    pitch = np.where(np.abs(sinp) >= 1, np.sign(sinp) * np.pi / 2, np.arcsin(np.clip(sinp, -1, 1)))
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return np.stack([roll, pitch, yaw], axis=-1)


def euler_to_quat(e: np.ndarray) -> np.ndarray:
    r, p, y = np.moveaxis(e * 0.5, -1, 0)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy], axis=-1)


def rotate(v: np.ndarray, q: np.ndarray, inverse: bool = False) -> np.ndarray:
    q = q / (np.linalg.norm(q, axis=-1, keepdims=True) + 1e-6)
    w, xyz = q[..., :1], q[..., 1:] * (-1 if inverse else 1)
    t = 2 * np.cross(xyz, v)
    return v + w * t + np.cross(xyz, t)


class Traj:
    """The fields of ksim's Trajectory the K-Bot rewards read, sliced out of one env's records [T][R]."""

    def __init__(self, prog, rec: np.ndarray) -> None:  # noqa: ANN001
        m = prog.spec.model
        g = lambda k, n: rec[:, prog[k] : prog[k] + n]  # noqa: E731
        self.qpos, self.qvel = g("TI_R_QPOS", m.nq), g("TI_R_QVEL", m.nv)
        self.xpos = g("TI_R_XPOS", 3 * m.nbody).reshape(-1, m.nbody, 3)
        self.xquat = g("TI_R_XQUAT", 4 * m.nbody).reshape(-1, m.nbody, 4)
        self.done = rec[:, prog["TI_R_DONE"]] != 0
        ob = g("TI_R_OBS", prog.obs_size)
        self.obs = {k: ob[:, o : o + d] for k, (o, d, _) in prog.obs_slices.items()}
        co, cd = prog.cmd_slices["unified_command"]
        self.cmd = g("TI_R_CMD", prog["TI_CMD_SIZE"])[:, co : co + cd]


def reference_rewards(prog, tr: Traj, carries: dict) -> dict[str, np.ndarray]:  # noqa: ANN001
    """train.py:128-496, unscaled, in the order of get_rewards (train.py:1223-1253)."""
    rw = prog.spec.rewards
    m = prog.spec.model
    zero = np.linalg.norm(tr.cmd[:, :3], axis=-1) < 1e-3
    left = tr.obs["left_foot_touch"][:, 0] > 0.1
    right = tr.obs["right_foot_touch"][:, 0] > 0.1
    out: dict[str, np.ndarray] = {}
    # LinearVelocityTrackingReward
    e = quat_to_euler(tr.xquat[:, 1, :]); e[:, :2] = 0
    g = rotate(np.pad(tr.cmd[:, :2], ((0, 0), (0, 1))), euler_to_quat(e))
    ve = np.linalg.norm(tr.qvel[:, :2] - g[:, :2], axis=-1)
    out["linvel"] = np.exp(-np.where(zero, ve, ve**2) / rw["linvel"].error_scale)
    out["angvel"] = np.exp(-np.abs(tr.qvel[:, 5] - tr.cmd[:, 2]) / rw["angvel"].error_scale)
    # XYOrientationReward
    e = quat_to_euler(tr.xquat[:, 1, :]); e[:, 2] = 0
    qc = euler_to_quat(np.stack([tr.cmd[:, 4], tr.cmd[:, 5], np.zeros(len(zero))], axis=-1))
    qe = 1 - np.sum(qc * euler_to_quat(e), axis=-1) ** 2
    out["roll_pitch"] = np.exp(-qe / np.where(zero, rw["roll_pitch"].error_scale_zero_cmd, rw["roll_pitch"].error_scale))
    # TerrainBaseHeightReward
    bh = rw["base_height"]
    low = np.minimum(tr.xpos[:, bh.foot_left_idx, 2], tr.xpos[:, bh.foot_right_idx, 2]) - bh.foot_origin_height
    out["base_height"] = np.exp(-np.abs(tr.xpos[:, bh.base_idx, 2] - low - (tr.cmd[:, 3] + bh.standard_height)) / bh.error_scale)
    # ArmPositionReward
    ap = rw["arm_pos"]
    err = ((tr.qpos[:, np.array(ap.joint_indices) + 7] - (tr.cmd[:, 6:16] + np.array(ap.joint_biases))) ** 2).sum(-1)
    out["arm_pos"] = np.exp(-err / ap.error_scale)
    # SingleFootContactReward (scan)
    sc, t = rw["single_contact"], carries["single_contact"].copy()
    single, ts = np.logical_xor(left, right), []
    for s in range(len(zero)):
        t = np.where(single[s], 0.0, t + sc.ctrl_dt)
        t = np.where(zero[s], sc.grace_period, t)
        ts.append(t)
    carries["single_contact"] = t
    out["single_contact"] = np.where(zero, 1.0, (np.array(ts)[:, 0] < sc.grace_period).astype(F))
    out["no_contact_p"] = np.where(zero, 0.0, np.where(left | right, 0.0, 1.0))
    # FeetAirtimeReward (scan + shift)
    fa = rw["feet_airtime"]
    air_c, con_c = carries["feet_airtime"]
    contact = np.stack([left, right], axis=-1)
    cod, t, airs = contact | tr.done[:, None], air_c.copy(), []
    for s in range(len(zero)):
        t = np.where(cod[s], 0.0, t + fa.ctrl_dt)
        airs.append(t)
    airtime = np.array(airs)
    prev = np.concatenate([con_c[None, :], contact[:-1]], axis=0)
    first = (contact & ~prev) * ~tr.done[:, None]
    shifted = np.concatenate([air_c[None, :], airtime], axis=0)[:-1]
    out["feet_airtime"] = np.where(zero, 0.0, np.sum((shifted - fa.touchdown_penalty) * first.astype(F), axis=-1))
    carries["feet_airtime"] = (airtime[-1], contact[-1])
    # FeetOrientationReward (is_rotating: norm over the TIME axis, train.py:464)
    fo = rw["feet_orient"]
    yaw = quat_to_euler(tr.xquat[:, 1, :])[:, 2]
    se = np.stack([np.stack([np.full_like(yaw, -np.pi / 2), np.zeros_like(yaw), yaw - np.pi], -1),
                   np.stack([np.full_like(yaw, np.pi / 2), np.zeros_like(yaw), yaw - np.pi], -1)], axis=1)
    fq = tr.xquat[:, [fo.foot_left_idx, fo.foot_right_idx], :]
    rpy = np.sum(1 - np.sum(euler_to_quat(se) * fq, axis=-1) ** 2, axis=-1)
    fe = quat_to_euler(fq); fe[:, :, 2] = 0
    se0 = se.copy(); se0[:, :, 2] = 0
    rp = (1 - np.sum(euler_to_quat(se0) * euler_to_quat(fe), axis=-1) ** 2).sum(axis=1)
    rotating = np.linalg.norm(tr.cmd[:, 2], axis=-1) > 1e-3
    out["feet_orient"] = np.exp(-np.where(rotating, rp, rpy) / fo.error_scale)
    # COMDistanceReward
    cd = tr.obs["com_distance"][:, 0]
    out["com_distance"] = np.where(cd >= 0, np.where(zero, np.exp(-cd / rw["com_distance"].error_scale), 0.0), 0.0)
    # BaseAccelerationReward
    bv = np.pad(tr.qvel[:, :6], ((1, 0), (0, 0)), mode="edge")
    dp = np.pad(tr.done, ((1, 0),), mode="edge")
    acc = np.where(dp[:-1, None], 0.0, bv[1:] - bv[:-1])
    out["base_accel"] = np.exp(-np.abs(acc).sum(-1) / rw["base_accel"].error_scale)
    return out


def reference_unified_command(cmd, key: np.ndarray) -> np.ndarray:  # noqa: ANN001
    """UnifiedCommand.initial_command (train.py:714-756) on the numpy restatement of jax.random (ksim_b200/rng.py)."""
    ks = R.split(key, 9)
    u = lambda k, rng: R.uniform(k, 1, rng[0], rng[1])  # noqa: E731
    vx, vy, wz = u(ks[1], cmd.vx_range), u(ks[2], cmd.vy_range), u(ks[3], cmd.wz_range)
    bh, rx, ry = u(ks[4], cmd.bh_range), u(ks[5], cmd.rx_range), u(ks[6], cmd.ry_range)
    arms = R.uniform(ks[7], 10, np.array(cmd.arms_range[0]), np.array(cmd.arms_range[1])) * R.bernoulli(ks[7], 10, 0.5)
    z1, z10 = np.zeros(1, np.float32), np.zeros(10, np.float32)
    modes = [np.concatenate(v) for v in ([vx, z1, z1, z1, z1, z1, z10], [z1, vy, z1, z1, z1, z1, z10], [z1, z1, wz, z1, z1, z1, z10],
                                         [vx, vy, wz, z1, z1, z1, arms], [z1, z1, z1, bh, rx, ry, arms], [z1] * 6 + [z10])]
    # jax.random.randint(key, (), 0, 6): two draws from split(key), ((hi % 6) * (2^32 % 6) + lo % 6) % 6
    k1, k2 = R.split(ks[0], 2)
    hi, lo = int(R.bits(k1, 1)[0]), int(R.bits(k2, 1)[0])
    mode = ((hi % 6) * ((1 << 32) % 6) + lo % 6) % 6
    return modes[mode]


def reference_com_distance(points: np.ndarray, com: np.ndarray) -> float:
    """COMDistanceObservation (train.py:508-649): lexsort, monotone chain with `cross <= 0` pops, both chains minus their
    last vertex, shoelace centroid with the vertex-mean fallback."""
    order = np.lexsort((points[:, 1], points[:, 0]))
    sp = points[order]
    cross = lambda a, b, c: (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])  # noqa: E731

    def build(idx):  # noqa: ANN001, ANN202
        st: list[int] = []
        for i in idx:
            while len(st) >= 2 and cross(sp[st[-2]], sp[st[-1]], sp[i]) <= 0:
                st.pop()
            st.append(i)
        return st

    n = len(sp)
    lower, upper = build(range(n)), build(range(n - 1, -1, -1))
    poly = sp[lower[:-1] + upper[:-1]] if len(lower) + len(upper) > 2 else sp[:0]
    if len(poly) == 0:
        c = np.zeros(2)
    else:
        nxt = np.roll(poly, -1, axis=0)
        cr = poly[:, 0] * nxt[:, 1] - nxt[:, 0] * poly[:, 1]
        area = 0.5 * cr.sum()
        c = poly.mean(axis=0) if abs(area) < 1e-12 else np.array([((poly[:, 0] + nxt[:, 0]) * cr).sum(), ((poly[:, 1] + nxt[:, 1]) * cr).sum()]) / (6 * area)
    return float(np.linalg.norm(c - com[:2]))


# ----------------------------------------------------------------------------------------------------- tests


@pytest.fixture(scope="module")
def kbot_program(kbot_spec):  # noqa: ANN001, ANN201
    return build_program(kbot_spec)


def _oracle_rollout(prog, rz, n, t, precision, seed=0, actions_scale=0.3):  # noqa: ANN001, ANN202
    from ksim_b200.envinit import make_env_init
    from oracle import OracleEngine

    init = make_env_init(prog, rz, n, seed=seed, level=1.0)
    ora = OracleEngine(prog, precision)
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), ora.real)
    rec[0] = rec0
    actions = (actions_scale * np.random.RandomState(seed + 1).randn(t, n, prog.action_size)).astype(ora.real)
    ora.rollout(actions, init.rand.astype(ora.real), st, keys, rec)
    return ora, init, rec


def test_unified_command_matches_the_jax_restatement(oracle_built, kbot_program) -> None:  # noqa: ANN001
    import oracle

    lib = oracle.lib("f32")
    cmd = kbot_program.spec.commands["unified_command"]
    _, _, cf, dim = cmd.encode()
    cf = np.asarray(cf, np.float32)
    modes = set()
    for seed in range(200):
        key = R.split(R.prng_key(seed), 2)[1]
        out = np.zeros(16, np.float32)
        lib.o_test_cmd_unified(ctypes.c_void_p(cf.ctypes.data), ctypes.c_void_p(key.ctypes.data), ctypes.c_void_p(out.ctypes.data))
        ref = reference_unified_command(cmd, key)
        np.testing.assert_array_equal(out, ref)
        modes.add(tuple(out != 0))
    assert dim == 16 and len(modes) >= 6  # all six modes (and several arm masks) were drawn


def test_com_distance_matches_the_hull_restatement(oracle_built) -> None:  # noqa: ANN001
    import oracle

    lib = oracle.lib("f64")
    lib.o_test_com_distance.restype = ctypes.c_double
    rs = np.random.RandomState(0)
    cases = [rs.randn(8, 2) for _ in range(50)]
    cases += [np.zeros((8, 2)), np.tile(rs.randn(1, 2), (8, 1)), np.linspace(0, 1, 8)[:, None] * np.array([[1.0, 2.0]])]  # degenerate
    cases += [np.concatenate([rs.randn(4, 2), np.zeros((4, 2))]), np.repeat(rs.randn(4, 2), 2, axis=0)]  # masked / duplicated points
    foot = np.array([[0.1, 0.05], [0.1, -0.05], [-0.1, 0.05], [-0.1, -0.05]])
    cases += [np.concatenate([foot + [0.0, 0.1], foot - [0.0, 0.1]])]  # two rectangular feet
    for pts in cases:
        pts = np.ascontiguousarray(pts, F)
        com = rs.randn(3)
        got = lib.o_test_com_distance(ctypes.c_void_p(pts.ctypes.data), ctypes.c_int(len(pts)), ctypes.c_void_p(com.ctypes.data))
        assert got == pytest.approx(reference_com_distance(pts, com), rel=1e-9, abs=1e-12)


def test_rewards_match_the_jax_restatement(oracle_built, kbot_program, kbot_randomizers) -> None:  # noqa: ANN001
    """Two consecutive rollouts (the stateful rewards' carries cross the boundary), f64 oracle vs vectorised numpy."""
    prog, n, t = kbot_program, 6, 40
    ora, init, rec = _oracle_rollout(prog, kbot_randomizers, n, 2 * t, "f64", seed=2, actions_scale=0.6)
    names = prog.reward_components
    scales = {k: (-1.0 if r.is_penalty else 1.0) * r.scale.scale for k, r in prog.spec.rewards.items()}
    carry = np.zeros((n, prog["TI_REW_CARRY_SIZE"]), F)
    carries = [{"single_contact": np.array([0.0]), "feet_airtime": (np.zeros(2), np.array([True, True]))} for _ in range(n)]
    assert rec[: 2 * t, :, prog["TI_R_DONE"]].sum() > 0, "the run must contain terminations"
    seen = {k: 0.0 for k in names}
    for chunk in range(2):
        r = np.ascontiguousarray(rec[chunk * t : (chunk + 1) * t])
        total, comps = ora.rewards(r, init.levels.astype(F), carry)
        for e in range(n):
            if chunk == 1 and rec[t - 1, e, prog["TI_R_DONE"]] != 0:   # get_rewards resets the carry where done[-1] (rl.py:206-214)
                carries[e] = {"single_contact": np.array([0.0]), "feet_airtime": (np.zeros(2), np.array([True, True]))}
            ref = reference_rewards(prog, Traj(prog, r[:, e]), carries[e])
            for k, name in enumerate(names):
                np.testing.assert_allclose(comps[k, e], ref[name] * scales[name], rtol=2e-5, atol=1e-9, err_msg=name)  # 1e-6 eps of the quaternion normalisation enters the two rotation formulas differently
                seen[name] += float(np.abs(comps[k, e]).sum())
            np.testing.assert_allclose(total[e], sum(ref[k] * scales[k] for k in names), rtol=2e-5, atol=1e-8)
    assert all(v > 0 for v in seen.values()), seen  # every reward was exercised with non-zero values


def test_observations_terminations_commands_properties(oracle_built, kbot_program, kbot_randomizers) -> None:  # noqa: ANN001
    prog, n, t = kbot_program, 64, 30
    m = prog.spec.model
    _, _, rec = _oracle_rollout(prog, kbot_randomizers, n, t, "f64", seed=4)
    ob = rec[1 : t + 1, :, prog["TI_R_OBS"] : prog["TI_R_OBS"] + prog.obs_size]   # pre-step obs of record s+1 = post-step state of s
    post = rec[:t]
    xpos = post[:, :, prog["TI_R_XPOS"] : prog["TI_R_XPOS"] + 3 * m.nbody].reshape(t, n, m.nbody, 3)
    xquat = post[:, :, prog["TI_R_XQUAT"] : prog["TI_R_XQUAT"] + 4 * m.nbody].reshape(t, n, m.nbody, 4)
    done = post[:, :, prog["TI_R_DONE"]] != 0
    live = ~done  # after a done the next observation belongs to the reset state
    fp = prog.spec.observations["feet_position"]
    o, d, _ = prog.obs_slices["feet_position"]
    yaw = quat_to_euler(xquat[:, :, fp.base_idx])[..., 2]
    yq = euler_to_quat(np.stack([np.zeros_like(yaw), np.zeros_like(yaw), yaw], -1))
    want = np.concatenate([rotate(xpos[:, :, i] - xpos[:, :, fp.base_idx], yq, inverse=True) for i in (fp.foot_left_idx, fp.foot_right_idx)], -1)
    np.testing.assert_allclose(ob[..., o : o + d][live], want[live], atol=5e-6)  # eps of the quaternion normalisation, as above
    o, _, _ = prog.obs_slices["base_height"]
    np.testing.assert_allclose(ob[..., o][live], xpos[:, :, 1, 2][live], atol=1e-12)
    o, _, _ = prog.obs_slices["com_distance"]
    assert (ob[..., o] >= 0).all() and (ob[..., o][live] < 1.0).all()
    # TerrainBadZTermination: first termination component
    tz = prog.spec.terminations["bad_z"]
    height = xpos[:, :, tz.base_idx, 2] - np.minimum(xpos[:, :, tz.foot_left_idx, 2], xpos[:, :, tz.foot_right_idx, 2])
    np.testing.assert_array_equal(post[:, :, prog["TI_R_TERM"]], np.where(height < tz.unhealthy_z, -1.0, 0.0))
    # commands: every value is one of the six mode patterns
    co, cd = prog.cmd_slices["unified_command"]
    cmd = rec[: t + 1, :, prog["TI_R_CMD"] + co : prog["TI_R_CMD"] + co + cd]
    nzp = {tuple(row) for row in (cmd[..., :6] != 0).reshape(-1, 6).tolist()}
    assert nzp <= {(True, False, False, False, False, False), (False, True, False, False, False, False), (False, False, True, False, False, False),
                   (True, True, True, False, False, False), (False, False, False, True, True, True), (False,) * 6}
    assert len(nzp) == 6


def test_warp_code_matches_oracle_on_the_task_plugins(oracle_built, kbot_program, kbot_randomizers, monkeypatch) -> None:  # noqa: ANN001
    """Observation block, command block and rewards of the full K-Bot task: host emulation of the warp code vs f32 oracle."""
    from tests.test_emu_vs_oracle import EMU, _p, _run
    import subprocess

    monkeypatch.setenv("B2S_EMU_SYNC_MASK", "0x7f")
    out = EMU / "libb2s_emu.so"
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-x", "c++", "-o", str(out), str(EMU / "emu.cpp")], check=True)
    emu = ctypes.CDLL(str(out))
    prog, n, t = kbot_program, 8, 30
    _, rec_e, rec_o, _, _, ora, blob, init = _run(emu, prog, kbot_randomizers, n=n, t=t, level=1.0, seed=7, state_rtol=1e-3)
    ob, cm = prog["TI_R_OBS"], prog["TI_R_CMD"]
    for name in ("feet_position", "base_height", "com_distance"):
        o, d, _ = prog.obs_slices[name]
        np.testing.assert_allclose(rec_e[1 : t + 1, :, ob + o : ob + o + d], rec_o[1 : t + 1, :, ob + o : ob + o + d], rtol=2e-3, atol=2e-3, err_msg=name)
    np.testing.assert_array_equal(rec_e[: t + 1, :, cm : cm + 16], rec_o[: t + 1, :, cm : cm + 16])
    k, c = prog["TI_N_REW_COMP"], prog["TI_REW_CARRY_SIZE"]
    carry_o, carry_e = np.zeros((n, c), np.float32), np.zeros((n, c), np.float32)
    rec_in = np.ascontiguousarray(rec_o[:t])
    total_o, comps_o = ora.rewards(rec_in, init.levels, carry_o)
    total_e, comps_e = np.zeros((n, t), np.float32), np.zeros((k, n, t), np.float32)
    emu.emu_rewards(_p(blob), n, t, _p(rec_in), _p(init.levels), _p(carry_e), _p(total_e), _p(comps_e))
    np.testing.assert_allclose(comps_e, comps_o, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(total_e, total_o, rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(carry_e, carry_o, atol=1e-6)


def test_collision_body_and_floor_friction_randomizers(kbot_spec, kbot_randomizers, kbot_program) -> None:  # noqa: ANN001
    """randomization.py:75-104, 380-442: value ranges, untouched geoms, key discipline (three successive splits), and the
    floor-friction override being accepted only as the no-op it is for this model."""
    from ksim_b200 import randomization
    from ksim_b200.envinit import make_env_init

    m = kbot_spec.model
    cb = kbot_randomizers["collision_body"]
    keys = R.split(R.prng_key(3), 50)
    out = cb(m, keys)
    size0, pos0 = m.arrays["geom_size"].astype(np.float32), m.arrays["geom_pos"].astype(np.float32)
    others = [g for g in range(m.ngeom) if g not in cb.geom_ids]
    np.testing.assert_array_equal(out["geom_size"][:, others], np.broadcast_to(size0[others], (50, len(others), 3)))
    np.testing.assert_array_equal(out["geom_pos"][:, others], np.broadcast_to(pos0[others], (50, len(others), 3)))
    g = list(cb.geom_ids)
    rs, ls = out["geom_size"][:, g, 0] / size0[g, 0], out["geom_size"][:, g, 1] / size0[g, 1]
    assert (np.abs(rs - 1) <= 0.01 + 1e-6).all() and (np.abs(ls - 1) <= 0.03 + 1e-6).all() and rs.std() > 1e-3 and ls.std() > 1e-3
    np.testing.assert_array_equal(out["geom_size"][:, g, 2], np.broadcast_to(size0[g, 2], (50, 4)))
    d = out["geom_pos"][:, g] - pos0[g]
    assert (np.abs(d) <= np.array([0.020, 0.005, 0.005]) + 1e-6).all() and (d.std(axis=(0, 1)) > 1e-4).all()
    # key discipline of the first draw: rng, radius_rng = split(rng); uniform(radius_rng, (4,), .99, 1.01)
    want = R.uniform(R.split(keys[7], 2)[1], 4, 0.99, 1.01) * size0[g, 0]
    np.testing.assert_array_equal(out["geom_size"][7, g, 0], want.astype(np.float32))
    # the program carries both fields; the floor-friction override is dropped as a checked no-op
    assert {"geom_size", "geom_pos"} <= set(kbot_program.rand_slices) and "geom_friction" not in kbot_program.rand_slices
    init = make_env_init(kbot_program, kbot_randomizers, 4, seed=0, level=1.0)
    o, c = kbot_program.rand_slices["geom_size"]
    assert (init.rand[:, o : o + c].reshape(4, m.ngeom, 3)[:, g, 0] != size0[g, 0]).any()
    ff = randomization.FloorFrictionRandomizer(floor_geom_id=int(cb.geom_ids[0]), scale_lower=0.5, scale_upper=1.5)  # a foot capsule
    assert not randomization.pair_friction_unaffected(m, ff(m, keys)["geom_friction"])
    with pytest.raises(NotImplementedError):
        make_env_init(kbot_program, {**kbot_randomizers, "floor_friction": ff}, 4, seed=0, level=1.0)
