"""The ksim/rewards.py plugins beyond the two example tasks' lists (SmallJointAcceleration / Jerk, AvoidLimits, Torque, Energy,
JointDeviation, FlatBody, NoRoll, LinkAcceleration / Jerk): the C oracle against an independent vectorised numpy restatement
written from the reference's jnp code (edge padding, done masking, get_norm, exp kernels), on records of a noisy humanoid
rollout with resets; then the host emulation of the CUDA reward code against the oracle.  The GPU leg is
tests/test_gpu_parity.py::test_more_rewards_on_gpu."""

import ctypes
import dataclasses

import numpy as np
import pytest

from tests.test_emu_vs_oracle import _p, _run, emu  # noqa: F401  (emu: fixture)


def extra_rewards(model):  # noqa: ANN001, ANN201
    from ksim_b200 import rewards as R
    return {
        "joint_acc": R.SmallJointAccelerationReward(scale=0.5, kernel_scale=0.1),
        "joint_jerk": R.SmallJointJerkReward(scale=0.25, kernel_scale=0.2),
        "limits": R.AvoidLimitsPenalty.create(model, factor=-0.3, scale=2.0),   # negative factor: limits inside the range, so they bite
        "torque": R.TorquePenalty.create(model, scale=0.01),
        "energy": R.EnergyPenalty.create(model, scale=0.02),
        "deviation": R.JointDeviationPenalty.create(model, ("abdomen_y", "knee_right", "elbow_left"), (0.1, -0.2, 0.3),
                                                    joint_weights=(1.0, 2.0, 0.5), scale=0.3),
        "deviation_l1": R.JointDeviationPenalty(norm="l1", joint_indices=(3, 4), joint_targets=(0.0, 0.1), scale=0.1),
        "flat": R.FlatBodyReward.create(model, ("torso", "foot_left", "foot_right"), scale=1.5),
        "no_roll": R.NoRollReward(scale=0.7, roll_scale=0.3),
        "link_acc": R.LinkAccelerationPenalty(scale=3.0),
        "link_jerk": R.LinkJerkPenalty(norm="l1", scale=4.0),
        "slip": R.FeetSlipPenalty(contact_obs="feet_contact", velocity_obs="feet_velocity", scale=0.2),
        "reach": R.ReachabilityPenalty(delta_max_j=tuple(0.05 + 0.01 * k for k in range(model.nu)), scale=0.6),
        "reach_l1": R.ReachabilityPenalty(delta_max_j=tuple(0.1 for _ in range(model.nu)), squared=False, scale=0.4),
        "symmetry": R.SymmetryReward.create(model, ("hip_y_left", "hip_y_right", "knee_left"), (0.1, -0.1, 0.25), alpha=0.3, scale=0.9),
        "pair": R.PairwiseSymmetryReward.create(model, "shoulder1_left", "shoulder1_right", left_zero=0.05, right_zero=-0.1,
                                                flipped=True, alpha=0.2, scale=1.1),
    }


def extra_spec(walking_spec):  # noqa: ANN001, ANN201
    """The walking task with the extra rewards; one unused observation makes room for the feet velocities (16 observation
    plugins: the humanoid-sized kernel instantiation still serves it)."""
    from ksim_b200 import observation

    model = walking_spec.model
    obs = {k: v for k, v in walking_spec.observations.items() if k != "center_of_mass_velocity"}
    obs["feet_velocity"] = observation.BodyVelocityObservation.create(physics_model=model, body_names=("foot_left", "foot_right"))
    return dataclasses.replace(walking_spec, observations=obs, rewards=extra_rewards(model))


@pytest.fixture(scope="module")
def extra_program(walking_spec):  # noqa: ANN001, ANN201
    from ksim_b200.program import build_program

    return build_program(extra_spec(walking_spec))


def _exp_kernel(x, scale):  # noqa: ANN001, ANN202
    return np.exp(-np.square(x) / (2 * scale**2))


def _exp_kernel_pen(x, scale, sq, ab):  # noqa: ANN001, ANN202
    return np.abs(x) * -ab + np.square(x) * -sq + np.exp(-np.square(x) / (2 * scale**2))


def _rot_inv(v, q):  # noqa: ANN001, ANN202
    from tests.test_kbot_task_plugins import rotate

    return rotate(v, q, inverse=True)


def _diffs(x, done, order):  # noqa: ANN001, ANN202
    """The reference's pattern (rewards.py:479-508, 809-841) for one env: x (T, ...) padded `order` at the front (edge),
    done padded the same way and cut by one; each difference level is zeroed where the (shifted) done flag is set."""
    xz = np.concatenate([np.repeat(x[:1], order, axis=0), x], axis=0)
    dz = np.concatenate([np.repeat(done[:1], order, axis=0), done], axis=0)[:-1]
    dz = dz.reshape((-1,) + (1,) * (x.ndim - 1))
    return xz, dz


def numpy_rewards(prog, rec, model, carry):  # noqa: ANN001, ANN201
    """{component name: (N, T)} from records rec[T][N][R], reference expression by expression (unscaled, unsigned).
    carry: {reward name: (N, dim)} of the stateful rewards, updated in place (reset to the initial carry where done[-1],
    rl.py:205-213)."""
    q0, v0, xp, xq, ct, dn = (prog[k] for k in ("TI_R_QPOS", "TI_R_QVEL", "TI_R_XPOS", "TI_R_XQUAT", "TI_R_CTRL", "TI_R_DONE"))
    t, n, _ = rec.shape
    nb = model.nbody
    out = {}
    rw = prog.spec.rewards
    for e in range(n):
        r = rec[:, e].astype(np.float64)
        qpos, qvel, ctrl, done = r[:, q0 : q0 + model.nq], r[:, v0 : v0 + model.nv], r[:, ct : ct + model.nu], r[:, dn] != 0
        xpos, xquat = r[:, xp : xp + 3 * nb].reshape(t, nb, 3), r[:, xq : xq + 4 * nb].reshape(t, nb, 4)
        res = {}
        for name, order in (("joint_acc", 2), ("joint_jerk", 3)):
            xz, dz = _diffs(qpos[:, 7:], done, order)
            vel = np.where(dz, 0.0, xz[1:] - xz[:-1])
            acc = np.where(dz[1:], 0.0, vel[1:] - vel[:-1])
            val = acc if order == 2 else np.where(dz[2:], 0.0, acc[1:] - acc[:-1])
            res[name] = _exp_kernel(val, rw[name].kernel_scale).mean(axis=-1)
        lim = np.array(rw["limits"].joint_limits)
        jp = qpos[:, 7:]
        res["limits"] = (-np.clip(jp - lim[:, 0], None, 0.0) + np.clip(jp - lim[:, 1], 0.0, None)).mean(axis=-1)
        cn = ctrl / np.array(rw["torque"].ctrl_scales)
        res["torque"] = np.abs(cn).mean(axis=-1)
        res["energy"] = np.abs(cn * qvel[:, 6:]).mean(axis=-1)
        for name in ("deviation", "deviation_l1"):
            d = rw[name]
            diff = qpos[:, np.array(d.joint_indices) + 7] - np.array(d.joint_targets)
            if d.joint_weights is not None:
                diff = diff * np.array(d.joint_weights)
            res[name] = (np.square(diff) if d.norm == "l2" else np.abs(diff)).sum(axis=-1)
        fl = rw["flat"]
        plane = np.array(fl.plane)
        res["flat"] = np.einsum("...i,...i->...", _rot_inv(np.broadcast_to(plane, (t, len(fl.body_indices), 3)),
                                                          xquat[:, list(fl.body_indices)]), plane).mean(axis=-1)
        nr = rw["no_roll"]
        zb = _rot_inv(np.broadcast_to(np.array([0.0, 0.0, 1.0]), (t, 3)), qpos[:, 3:7])
        res["no_roll/angvel"] = _exp_kernel_pen(qvel[:, 3], nr.angvel_scale, nr.angvel_sq_scale, nr.angvel_abs_scale)
        res["no_roll/pose"] = _exp_kernel_pen(-np.arctan2(zb[:, 1], zb[:, 2]), nr.roll_scale, nr.roll_sq_scale, nr.roll_abs_scale)
        for name, order in (("link_acc", 2), ("link_jerk", 3)):
            xz, dz = _diffs(xpos[:, 1:], done, order)
            vel = np.where(dz[..., 0], 0.0, np.linalg.norm(xz[1:] - xz[:-1], axis=-1))
            acc = np.where(dz[1:, :, 0], 0.0, vel[1:] - vel[:-1])
            val = acc if order == 2 else np.where(dz[2:, :, 0], 0.0, acc[1:] - acc[:-1])
            res[name] = (np.square(val) if rw[name].norm == "l2" else np.abs(val)).mean(axis=-1)
        ob, ac = prog["TI_R_OBS"], prog["TI_R_ACTION"]
        co, _, cshape = prog.obs_slices["feet_contact"]
        vo, _, vshape = prog.obs_slices["feet_velocity"]
        contact = (r[:, ob + co : ob + co + int(np.prod(cshape))].reshape((t,) + cshape) > 0.5).any(axis=-2)
        vel6 = r[:, ob + vo : ob + vo + int(np.prod(vshape))].reshape((t,) + vshape)
        res["slip"] = np.where(contact, np.linalg.norm(vel6, axis=-1), 0.0).sum(axis=-1)
        action = r[:, ac : ac + model.nu]
        for name in ("reach", "reach_l1"):
            excess = np.maximum(0.0, np.abs(action - qpos[:, 7:]) - np.array(rw[name].delta_max_j))
            res[name] = (excess**2 if rw[name].squared else excess).sum(axis=-1)
        sy = rw["symmetry"]
        c = np.array(sy.joint_targets) if done[-1] else carry["symmetry"][e]
        qs = qpos[:, np.array(sy.joint_indices) + 7] - np.array(sy.joint_targets)
        c = c * (1.0 - sy.alpha) + qs.mean(axis=-2) * sy.alpha
        res["symmetry"] = (qs * -c[None, :]).sum(axis=-1)
        carry["symmetry"][e] = c
        pw = rw["pair"]
        c = np.array([pw.left_zero, pw.right_zero]) if done[-1] else carry["pair"][e]
        ql, qr = qpos[:, pw.left_joint_index + 7] - pw.left_zero, qpos[:, pw.right_joint_index + 7] - pw.right_zero
        if pw.flipped:
            qr = -qr
        res["pair/lr"] = _exp_kernel_pen(ql - qr, pw.kernel_scale, pw.sq_scale, pw.abs_scale)
        qp = np.stack([ql, qr], axis=-1)
        c = c * (1.0 - pw.alpha) + qp.mean(axis=-2) * pw.alpha
        res["pair/self"] = (qp * -c[None, :]).sum(axis=-1)
        carry["pair"][e] = c
        for k, v in res.items():
            out.setdefault(k, np.zeros((n, t)))[e] = v
    return out


def test_oracle_matches_numpy_restatement(oracle_built, extra_program, walking_randomizers) -> None:  # noqa: ANN001
    from ksim_b200.envinit import make_env_init
    from oracle import OracleEngine

    prog = extra_program
    n, t = 12, 70
    init = make_env_init(prog, walking_randomizers, n, seed=7, level=1.0)
    ora = OracleEngine(prog, "f64")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size))
    rec[0] = rec0
    actions = 0.4 * np.random.RandomState(8).randn(t, n, prog.action_size)
    ora.rollout(actions, init.rand, st, keys, rec)
    done = rec[:t, :, prog["TI_R_DONE"]] != 0
    assert done.sum() >= 3 and done[0].sum() == 0, "the case must exercise the done masking away from step 0"
    # two consecutive reward passes over the two halves of the rollout: the stateful carries cross a boundary, and the
    # environments whose first half ends on a done step start the second half from the initial carry
    rw_all = prog.spec.rewards
    np_carry = {"symmetry": np.tile(np.array(rw_all["symmetry"].joint_targets), (n, 1)),
                "pair": np.tile(np.array([rw_all["pair"].left_zero, rw_all["pair"].right_zero]), (n, 1))}
    o_carry = np.zeros((n, prog["TI_REW_CARRY_SIZE"]))
    assert prog["TI_REW_CARRY_SIZE"] == 3 + 2
    h = t // 2
    done[h - 1, :3] = True
    rec[h - 1, :3, prog["TI_R_DONE"]] = 1.0   # force the reset-of-carry branch for three environments
    names = prog.reward_components
    for seg in (slice(0, h), slice(h, t)):
        total, comps = ora.rewards(np.ascontiguousarray(rec[seg]), init.levels, o_carry)
        want = numpy_rewards(prog, rec[seg], prog.spec.model, np_carry)
        assert set(names) == set(want), (names, sorted(want))
        signed_sum = np.zeros((n, h))
        for i, name in enumerate(names):
            rw = rw_all[name.split("/")[0]]
            sign = -1.0 if rw.is_penalty else 1.0
            scale = float(rw.scale.scale)
            # parameters are stored as fp32 (1e-6); the 1e-6 eps of xax's quaternion normalisation enters the two rotation
            # formulas (polynomial in the oracle, cross products here) differently (2e-5, as in tests/test_kbot_task_plugins.py)
            rot = name in ("flat", "no_roll/pose")
            np.testing.assert_allclose(comps[i], want[name] * scale * sign, rtol=2e-5 if rot else 1e-6, atol=2e-5 if rot else 1e-8, err_msg=name)
            assert np.abs(comps[i]).max() > 0, name
            signed_sum += comps[i]
        np.testing.assert_allclose(total, signed_sum, rtol=1e-9, atol=1e-9)


def test_emu_matches_oracle(emu, oracle_built, extra_program, walking_randomizers) -> None:  # noqa: ANN001, F811
    prog, rec_e, rec_o, _, _, ora, blob, init = _run(emu, extra_program, walking_randomizers, n=6, t=40, level=1.0, seed=5)
    n, t = 6, 40
    k, c = prog["TI_N_REW_COMP"], max(prog["TI_REW_CARRY_SIZE"], 1)
    carry_o, carry_e = np.zeros((n, c), np.float32), np.zeros((n, c), np.float32)
    for seg in (slice(0, t // 2), slice(t // 2, t)):   # two passes: the stateful carries cross a boundary
        h = t // 2
        rec_in = np.ascontiguousarray(rec_o[seg])
        total_o, comps_o = ora.rewards(rec_in, init.levels, carry_o)
        total_e, comps_e = np.zeros((n, h), np.float32), np.zeros((k, n, h), np.float32)
        emu.emu_rewards(_p(blob), n, h, _p(rec_in), _p(init.levels), _p(carry_e), _p(total_e), _p(comps_e))
        np.testing.assert_allclose(comps_e, comps_o, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(total_e, total_o, rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(carry_e, carry_o, rtol=1e-5, atol=1e-6)


def test_encode_rejects_mismatched_sizes(walking_spec) -> None:  # noqa: ANN001
    from ksim_b200 import rewards as R
    from ksim_b200.program import build_program

    for bad in (R.TorquePenalty(ctrl_scales=(1.0, 2.0), scale=1.0), R.AvoidLimitsPenalty(joint_limits=((0.0, 1.0),), scale=1.0),
                R.JointDeviationPenalty(joint_indices=(99,), joint_targets=(0.0,), scale=1.0),
                R.FlatBodyReward(body_indices=(), scale=1.0)):
        with pytest.raises(ValueError):
            build_program(dataclasses.replace(walking_spec, rewards={"x": bad}))
    with pytest.raises(ValueError):
        R.AvoidLimitsPenalty(joint_limits=((1.0, 0.0),), scale=1.0)
