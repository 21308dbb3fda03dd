"""``CartesianCoordinateCommand`` / ``SinusoidalGaitCommand`` / ``JointPositionCommand`` (ksim/commands.py:537-900) and the rewards
that read them -- ``PositionTrackingReward`` (rewards.py:704-744), ``SinusoidalGaitReward`` (1320-1348), ``JointPositionReward``
(1352-1414) -- plus ``IntersectionPenalty`` (1418-1430).

CPU: the oracle against a numpy transcription of the reference's expressions (commands: on this repo's threefry restatement,
driven by the environment keys the oracle exposes before every step; rewards: on the trajectory's own recorded fields); the host
emulation of the kernel code against the oracle.  GPU: CUDA == oracle (commands 1e-6, rewards 1e-5)."""

import numpy as np
import pytest

from ksim_b200 import commands, observation, rewards, rng as R, terminations
from ksim_b200.program import build_program
from tests.helpers import tripod_model, tripod_spec
from tests.test_emu_vs_oracle import _p, emu  # noqa: F401
from tests.test_more_commands import _rollout_with_keys

BOX_MIN, BOX_MAX = (-0.3, -0.2, 0.1), (0.3, 0.2, 0.5)
CART = dict(dt=0.02, min_speed=0.5, max_speed=3.0, switch_prob=0.7, jump_prob=0.1, curriculum_scale=0.5)
GAIT = dict(gait_period=0.25, ctrl_dt=0.02, max_height=0.15, height_offset=0.04, num_feet=2, stance_ratio=0.6, moving_threshold=0.02)
JP_RANGES = ((-0.5, 0.5), (-0.3, 0.8))


def _spec():  # noqa: ANN202
    model = tripod_model()
    cmds = {"cart": commands.CartesianCoordinateCommand(box_min=BOX_MIN, box_max=BOX_MAX, **CART),
            "gait": commands.SinusoidalGaitCommand(**GAIT),
            "jpos": commands.JointPositionCommand(indices=(0, 2), ranges=JP_RANGES, ctrl_dt=0.02, min_time=0.03, max_time=0.12)}
    obs = {"joint_position": observation.JointPositionObservation(),
           "feet_pos": observation.BodyPositionObservation.create(physics_model=model, body_names=["foot_a", "foot_b"]),
           "all_feet": observation.BodyPositionObservation.create(physics_model=model, body_names=["foot_a", "foot_b", "foot_c"])}
    rews = {"track": rewards.PositionTrackingReward.create(model, "cart", "foot_a", "base", norm="l2", kernel_scale=0.4, scale=1.0),
            "gait": rewards.SinusoidalGaitReward(ctrl_dt=0.02, max_height=0.15, pos_obs="feet_pos", pos_cmd="gait", scale=0.5),
            "jpos": rewards.JointPositionReward(command_name="jpos", indices=(0, 2), scale=2.0),
            "intersect": rewards.IntersectionPenalty(position_obs="all_feet", min_distance=0.27, scale=1.0)}
    return tripod_spec(cmds=cmds, observations=obs, rewards=rews,
                       terminations={"episode_length": terminations.EpisodeLengthTermination(max_length_sec=0.23)})


f32 = np.float32


def _sample_box(key, level):  # noqa: ANN001, ANN202
    lo, hi = np.array(BOX_MIN, f32), np.array(BOX_MAX, f32)
    center, half = (lo + hi) / f32(2), (hi - lo) / f32(2) * (f32(CART["curriculum_scale"]) * f32(level) + f32(1))
    return R.uniform(key, 3, center - half, center + half)


def _cart_initial(key, level):  # noqa: ANN001, ANN202
    kt, ks = R.split(key, 2)
    tg = _sample_box(kt, level)
    return np.concatenate([tg, tg, R.uniform(ks, 1, CART["min_speed"], CART["max_speed"])]).astype(f32)


def _cart_update(prev, key, level):  # noqa: ANN001, ANN202
    cur, tg, speed = prev[:3], prev[3:6], prev[6]
    dist = np.sqrt(((tg - cur) ** 2).sum(dtype=f32), dtype=f32)
    ka, kb, kc = R.split(key, 3)
    reached = dist < f32(CART["dt"]) * speed * f32(0.5)
    change = (reached and R.bernoulli(kc, 1, CART["switch_prob"])[0]) or R.bernoulli(kc, 1, CART["jump_prob"])[0]
    if change:
        tg, speed = _sample_box(ka, level), R.uniform(kb, 1, CART["min_speed"], CART["max_speed"])[0]
    d = tg - cur
    dn = np.sqrt((d * d).sum(dtype=f32), dtype=f32)
    d = d / dn if dn > 0 else d
    return np.concatenate([cur + d * min(speed * f32(CART["dt"]), dist), tg, [speed]]).astype(f32)


def _gait_heights(phase):  # noqa: ANN001, ANN202
    pf = (f32(phase) + np.arange(2, dtype=f32) / f32(2)) % f32(1)
    prog = np.clip((pf - f32(0.6)) / (f32(1) - f32(0.6)), 0, 1)
    return (np.where(pf >= f32(0.6), f32(0.15) * np.sin(f32(np.pi) * prog), 0) + f32(0.04)).astype(f32)


def test_oracle_commands_match_numpy_restatement(oracle_built) -> None:  # noqa: ANN001
    spec = _spec()
    n, t = 12, 26
    prog, init, rec, _, env_keys, states = _rollout_with_keys(spec, n, t)
    cm = prog["TI_R_CMD"]
    sl = {k: slice(cm + o, cm + o + d) for k, (o, d) in prog.cmd_slices.items()}
    names = list(spec.commands)
    done = rec[:t, :, prog["TI_R_DONE"]] != 0
    qs, vs = prog["TI_S_QPOS"], prog["TI_S_QVEL"]
    n_change = n_fresh_jp = n_stopped = 0
    for e in range(n):
        prev = {k: rec[0, e, v].copy() for k, v in sl.items()}
        key = init.init_keys[e, 2:4]
        for name in names:
            key, k = R.split(key, 2)
            if name == "cart":
                np.testing.assert_allclose(prev["cart"], _cart_initial(k, 1.0), rtol=0, atol=1e-6)
            if name == "gait":
                np.testing.assert_allclose(prev["gait"], np.concatenate([[1, 0], _gait_heights(0)]), rtol=0, atol=1e-6)
        for s in range(t):
            key = R.split(R.split(env_keys[s, e], 3)[2], 6)[0]
            for name in names:
                key, k = R.split(key, 2)
                got = rec[s + 1, e, sl[name]]
                qpos, qvel = states[s, e, qs : qs + 10], states[s, e, vs : vs + 9]
                if name == "cart":
                    want = _cart_initial(k, 1.0) if done[s, e] else _cart_update(prev[name], k, 1.0)
                    n_change += int(not done[s, e] and not np.allclose(want[3:6], prev[name][3:6]))
                    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6, err_msg=f"cart env {e} step {s}")
                elif name == "gait":
                    if done[s, e]:
                        want = np.concatenate([[1, 0], _gait_heights(0)])
                    else:
                        moving = prev[name][0] != 0
                        ph = (prev[name][1] + f32(0.02) / f32(0.25)) % f32(1) if moving else f32(0)
                        is_moving = np.sqrt(qvel[0] ** 2 + qvel[1] ** 2) > 0.02
                        if not is_moving:
                            moving, ph = False, f32(0)
                            n_stopped += 1
                        want = np.concatenate([[float(moving), ph], _gait_heights(ph)])
                    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6, err_msg=f"gait env {e} step {s}")
                else:
                    step_id = prev[name][6] + 1
                    if done[s, e] or step_id > prev[name][5]:
                        ka, kb = R.split(k, 2)
                        tt = R.uniform(kb, 1, 0.03, 0.12)[0]
                        tg = R.uniform(ka, 2, np.array([r[0] for r in JP_RANGES], f32), np.array([r[1] for r in JP_RANGES], f32))
                        want = np.concatenate([qpos[[7, 9]], tg, [tt, np.ceil(tt / f32(0.02)), 0]])
                        n_fresh_jp += 1
                    else:
                        want = np.concatenate([prev[name][:5], [prev[name][5], step_id]])
                    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6, err_msg=f"jpos env {e} step {s}")
                prev[name] = got.copy()
    assert n_change > 5 and n_fresh_jp > 20


def _want_rewards(spec, prog, rec, t):  # noqa: ANN001, ANN202
    m = spec.model
    cm, ob = prog["TI_R_CMD"], prog["TI_R_OBS"]
    co = {k: cm + o for k, (o, d) in prog.cmd_slices.items()}
    xpos = rec[:t, :, prog["TI_R_XPOS"] : prog["TI_R_XPOS"] + 3 * m.nbody].reshape(t, -1, m.nbody, 3).astype(np.float64)
    r = rec[:t].astype(np.float64)
    out = {}
    err = (((xpos[:, :, m.id("body", "foot_a")] - xpos[:, :, m.id("body", "base")]) - r[:, :, co["cart"] : co["cart"] + 3]) ** 2).sum(-1)
    out["track"] = np.exp(-(err**2) / (2 * 0.4**2))
    o, d, _ = prog.obs_slices["feet_pos"]
    feet = r[:, :, ob + o : ob + o + d].reshape(t, -1, 2, 3)
    out["gait"] = 0.5 * (1.0 - np.abs(feet[..., 2] - r[:, :, co["gait"] + 2 : co["gait"] + 4]).sum(-1) / 0.15)
    c = co["jpos"]
    start, target = r[:, :, c : c + 2], r[:, :, c + 2 : c + 4]
    total_time, num_steps, step_id = r[:, :, c + 4 : c + 5], r[:, :, c + 5 : c + 6], r[:, :, c + 6 : c + 7]
    qp, qv = prog["TI_R_QPOS"], prog["TI_R_QVEL"]
    pos_diff = start + (target - start) * step_id / num_steps - r[:, :, [qp + 7, qp + 9]]
    vel_diff = (target - start) / total_time - r[:, :, [qv + 6, qv + 8]]
    pen = lambda x: np.exp(-(x**2) / (2 * 0.25**2)) - 0.1 * x**2 - 0.1 * np.abs(x)  # noqa: E731
    out["jpos/pos"], out["jpos/vel"] = 2.0 * pen(pos_diff).mean(-1), 2.0 * pen(vel_diff).mean(-1)
    o, d, _ = prog.obs_slices["all_feet"]
    pts = r[:, :, ob + o : ob + o + d].reshape(t, -1, 3, 3)
    dist = np.linalg.norm(pts[:, :, :, None, :] - pts[:, :, None, :, :], axis=-1) + np.triu(np.ones((3, 3))) * 0.27 * 2
    out["intersect"] = -np.where((dist < 0.27).any(axis=(-2, -1)), 1.0, 0.0)
    return out


def test_oracle_rewards_match_numpy_restatement(oracle_built) -> None:  # noqa: ANN001
    from oracle import OracleEngine

    spec = _spec()
    n, t = 16, 26
    prog, init, rec, *_ = _rollout_with_keys(spec, n, t)
    total, comps = OracleEngine(prog, "f32").rewards(rec[:t], init.levels, np.zeros((n, max(prog["TI_REW_CARRY_SIZE"], 1)), np.float32))
    want = _want_rewards(spec, prog, rec, t)
    names = prog.reward_components
    for k, w in want.items():
        np.testing.assert_allclose(comps[names.index(k)].T, w, rtol=2e-5, atol=2e-6, err_msg=k)
    assert 0 < (want["intersect"] != 0).mean() < 1, "both branches of the intersection penalty should occur"
    np.testing.assert_allclose(total.T, sum(want.values()), rtol=1e-4, atol=1e-5)


def test_emu_matches_oracle(emu, oracle_built) -> None:  # noqa: ANN001, F811
    from ksim_b200.blob import pack_blob
    from oracle import OracleEngine

    spec = _spec()
    n, t = 8, 20
    prog, init, rec_o, actions, _, _ = _rollout_with_keys(spec, n, t)
    blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
    st_e = np.zeros((n, prog.state_size), np.float32)
    k_e = np.zeros((n, prog.key_size), np.uint32)
    rec_e = np.zeros_like(rec_o)
    ak = np.zeros((n, 2), np.uint32)
    emu.emu_init_envs(_p(blob), n, _p(init.init_keys), _p(init.levels), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[0]), _p(ak))
    for s in range(t):
        emu.emu_step_ctrl(_p(blob), n, _p(actions[s]), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[s]), _p(rec_e[s + 1]), None)
    cm, cs = prog["TI_R_CMD"], prog["TI_CMD_SIZE"]
    np.testing.assert_allclose(rec_e[:, :, cm : cm + cs], rec_o[:, :, cm : cm + cs], rtol=0, atol=5e-5)
    # the reward code alone, on the ORACLE's records
    k = len(prog.reward_components)
    tot_o, comps_o = OracleEngine(prog, "f32").rewards(rec_o[:t], init.levels, np.zeros((n, 1), np.float32))
    tot_e, comps_e = np.zeros((n, t), np.float32), np.zeros((k, n, t), np.float32)
    carry = np.zeros((n, 1), np.float32)
    emu.emu_rewards(_p(blob), n, t, _p(np.ascontiguousarray(rec_o[:t])), _p(init.levels), _p(carry), _p(tot_e), _p(comps_e))
    np.testing.assert_allclose(comps_e, comps_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tot_e, tot_o, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine
    from oracle import OracleEngine

    spec = _spec()
    n, t = 48, 26
    prog, init, rec_o, actions, _, _ = _rollout_with_keys(spec, n, t)
    eng = B200Engine(spec, n, device="cuda:0", seed=7, curriculum_level=1.0)
    rec = eng.rollout(torch.from_numpy(actions).cuda())
    got = rec.cpu().numpy()
    np.testing.assert_array_equal(got[:t, :, prog["TI_R_DONE"]], rec_o[:t, :, prog["TI_R_DONE"]])
    cm, cs = prog["TI_R_CMD"], prog["TI_CMD_SIZE"]
    ok = np.abs(got[:, :, cm : cm + cs] - rec_o[:, :, cm : cm + cs]).max(axis=(0, 2)) < 1e-3   # physics-dependent commands: drop drifting envs
    assert ok.mean() >= 0.9
    o, d = prog.cmd_slices["cart"]
    np.testing.assert_allclose(got[:, :, cm + o : cm + o + d], rec_o[:, :, cm + o : cm + o + d], rtol=0, atol=2e-6)   # RNG only
    # rewards of the ORACLE's records through the CUDA scoring kernels
    tot_o, comps_o = OracleEngine(prog, "f32").rewards(rec_o[:t], init.levels, np.zeros((n, 1), np.float32))
    dev_rec = torch.from_numpy(rec_o).cuda()
    total, comps = eng.rewards(dev_rec, t)
    np.testing.assert_allclose(comps.cpu().numpy(), comps_o, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(total.cpu().numpy(), tot_o, rtol=1e-5, atol=2e-6)
    scored = eng.score(dev_rec, t)
    np.testing.assert_allclose(scored["comps"].cpu().numpy(), comps_o, rtol=1e-5, atol=2e-6)
