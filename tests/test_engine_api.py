"""The engine-only boundary (``b200sim_engine_reset`` / ``b200sim_engine_step`` / ``b200sim_terminations``: ksim's
``PhysicsEngine.reset`` / ``.step`` and ``get_terminations``) and the reference's own known-answer vectors pushed through the
CUDA path (not through the oracle): the 9 action-penalty reward vectors of ``tests/test_rewards.py:60-271`` through
``b200sim_rewards`` and the 7 termination cases of ``tests/test_terminations.py:66-124`` through ``b200sim_terminations``."""

import json
from pathlib import Path

import numpy as np
import pytest

from ksim_b200 import codes as C
from ksim_b200 import rewards, rng, terminations
from tests.helpers import synthetic_records, tripod_spec

GOLDEN = Path(__file__).resolve().parent / "golden"

# four geoms whose ids reproduce the reference fixture's contact list (geom1 = [0, 2], geom2 = [1, 3]) as explicit pairs
FOUR_GEOMS = """
<mujoco model="four_geoms">
  <option timestep="0.004" iterations="8" ls_iterations="8" tolerance="1e-10" integrator="implicitfast"/>
  <default>
    <joint damping="0.5" armature="0.01" limited="true" range="-60 60"/>
    <geom contype="0" conaffinity="0" condim="1" density="800"/>
    <motor ctrlrange="-1 1" ctrllimited="true"/>
  </default>
  <worldbody>
    <geom name="g0" type="plane" size="10 10 0.1"/>
    <body name="base" pos="0 0 0.5">
      <freejoint name="root"/>
      <geom name="g1" type="sphere" size="0.1"/>
      <body name="arm_a" pos="0.2 0 0">
        <joint name="ja" axis="0 1 0"/>
        <geom name="g2" type="sphere" size="0.05"/>
      </body>
      <body name="arm_b" pos="-0.2 0 0">
        <joint name="jb" axis="0 1 0"/>
        <geom name="g3" type="sphere" size="0.05"/>
      </body>
    </body>
  </worldbody>
  <contact>
    <pair geom1="g0" geom2="g1"/>
    <pair geom1="g2" geom2="g3"/>
  </contact>
  <actuator>
    <motor name="ma" joint="ja" gear="10"/>
    <motor name="mb" joint="jb" gear="10"/>
  </actuator>
</mujoco>
"""


def _four_geom_spec(term):  # noqa: ANN001, ANN202
    from ksim_b200 import commands, observation, resets
    from ksim_b200.actuators import TorqueActuators
    from ksim_b200.mjcf import load_mjcf_string
    from ksim_b200.program import EngineConfig, TaskSpec

    model = load_mjcf_string(FOUR_GEOMS)
    return TaskSpec(model=model, config=EngineConfig(dt=0.004, ctrl_dt=0.02), actuators=TorqueActuators(),
                    resets=[resets.RandomJointPositionReset.create(model)], events={},
                    observations={"joint_position": observation.JointPositionObservation()},
                    commands={"vec": commands.FloatVectorCommand(ranges=((0.0, 1.0),))}, rewards={}, terminations={"t": term})


@pytest.mark.gpu
def test_reference_reward_vectors_through_cuda() -> None:
    import torch

    from ksim_b200.engine import B200Engine

    cases = json.loads((GOLDEN / "reward_vectors.json").read_text())["cases"]
    assert len(cases) == 9
    for case in cases:
        cls = getattr(rewards, case["reward"])
        eng = B200Engine(tripod_spec(rewards={"pen": cls(scale=1.0, norm=case["norm"])}), 1, device="cuda:0")
        actions = np.array(case["actions"], dtype=np.float64)
        t = actions.shape[0]
        rec = torch.zeros((t + 1, 1, eng.R), device="cuda:0")
        rec[:t] = torch.from_numpy(synthetic_records(eng.program, actions, np.array(case["done"], dtype=np.float64), np.float32)).cuda()
        total, comps = eng.rewards(rec)        # default: the T transitions of a [T+1] record buffer
        assert total.shape == (1, t)
        # the engine applies the Penalty sign flip (rewards.py:118-124, rl.py:219-221); the reference's vectors are pre-flip
        np.testing.assert_allclose(-comps[0, 0].cpu().numpy(), case["expected"], rtol=0, atol=1e-5, err_msg=case["ref"])
        np.testing.assert_array_equal(total[0].cpu().numpy(), comps[0, 0].cpu().numpy())


@pytest.mark.gpu
def test_reference_termination_truth_tables_through_cuda() -> None:
    import torch

    from ksim_b200.engine import B200Engine

    data = json.loads((GOLDEN / "termination_vectors.json").read_text())
    assert len(data["cases"]) == 7
    for case in data["cases"]:
        eng = B200Engine(_four_geom_spec(getattr(terminations, case["termination"])(**case["args"])), 2, device="cuda:0")
        d = eng.new_data()
        off = d.offsets
        assert d.qpos.shape == (2, 9) and off[C.B2S_DATA_N_FIELDS] == d.block.shape[1]
        d.block[:, off[C.B2S_DATA_QPOS] : off[C.B2S_DATA_QPOS] + 7] = torch.tensor(data["qpos"], dtype=torch.float32, device="cuda:0")
        con = data["contacts"]
        ncon = off[C.B2S_DATA_CONTACT_GEOM2] - off[C.B2S_DATA_CONTACT_GEOM1]
        assert ncon == 2
        d.block[:, off[C.B2S_DATA_CONTACT_DIST] : off[C.B2S_DATA_CONTACT_DIST] + ncon] = 1.0
        if case["contacts"]:
            for name, field in (("geom1", C.B2S_DATA_CONTACT_GEOM1), ("geom2", C.B2S_DATA_CONTACT_GEOM2), ("dist", C.B2S_DATA_CONTACT_DIST)):
                d.block[:, off[field] : off[field] + ncon] = torch.tensor(con[name], dtype=torch.float32, device="cuda:0")
        # env 1: the same state with no contact at all and a raised, upright base -> every one of these terminations is silent
        d.block[1, off[C.B2S_DATA_CONTACT_DIST] : off[C.B2S_DATA_CONTACT_DIST] + ncon] = 1.0
        d.block[1, off[C.B2S_DATA_QPOS] : off[C.B2S_DATA_QPOS] + 7] = torch.tensor([0, 0, 5, 1, 0, 0, 0], dtype=torch.float32, device="cuda:0")
        codes, done, success = eng.get_terminations(d)
        got = int(codes["t"][0])
        assert got == case["expected"], case["ref"]
        assert bool(done[0]) == (case["expected"] != 0) and bool(success[0]) == (case["expected"] == 1)
        assert int(codes["t"][1]) == 0 and not bool(done[1])


@pytest.mark.gpu
def test_engine_step_is_the_fused_step_minus_the_task_layer(walking_spec, walking_randomizers) -> None:
    """``b200sim_engine_step`` on a copy of the environments, keyed with the physics key the fused step derives
    (``split(env.rng, 3)[1]``, rl.py:1180), lands bit-exactly on the post-step fields ``b200sim_step_ctrl`` records."""
    import torch

    from ksim_b200.engine import B200Engine

    n = 96
    engs = [B200Engine(walking_spec, n, device="cuda:0", randomizers=walking_randomizers, seed=3, curriculum_level=0.5) for _ in range(2)]
    fused, alone = engs
    torch.testing.assert_close(fused.state, alone.state, rtol=0, atol=0)
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    p = fused.program
    for step in range(3):
        action = 0.4 * torch.randn((n, fused.NA), generator=gen, device="cuda:0")
        env_keys = fused.keys[:, :2].cpu().numpy().view(np.uint32)
        physics_rng = torch.from_numpy(np.ascontiguousarray(rng.split(env_keys, 3)[:, 1, :]).view(np.int32)).cuda()
        alone.state.copy_(fused.state); alone.keys.copy_(fused.keys)      # same PhysicsState going in
        rec = torch.zeros((2, n, fused.R), device="cuda:0")
        rec[0].copy_(fused.first_rec)
        fused.step(action, rec[0], rec[1])
        fused.first_rec.copy_(rec[1])
        d = alone.engine_step(action, physics_rng)
        view = fused.view(rec[:1])
        torch.testing.assert_close(d.qpos, view.qpos[0], rtol=0, atol=0)
        torch.testing.assert_close(d.qvel, view.qvel[0], rtol=0, atol=0)
        torch.testing.assert_close(d.xpos, view.xpos[0], rtol=0, atol=0)
        torch.testing.assert_close(d.xquat, view.xquat[0], rtol=0, atol=0)
        torch.testing.assert_close(d.ctrl, view.ctrl[0], rtol=0, atol=0)
        torch.testing.assert_close(d.time, view.timestep[0], rtol=0, atol=0)
        # the plugins' other inputs are there and sane: unit quaternions, positive masses in cinert, contact slots of the two feet
        assert torch.allclose(d.xquat[:, 1:].norm(dim=-1), torch.ones(n, walking_spec.model.nbody - 1, device="cuda:0"), atol=1e-4)
        assert bool((d.cinert[:, 1:, 9] > 0).all()) and bool(torch.isfinite(d.block).all())
        assert d.contact.geom1.shape == (n, 2) and int(d.ncon[0]) == 2
        # get_terminations on the data block = the codes the fused step recorded
        codes, done, success = alone.get_terminations(d, curriculum_level=fused.levels)
        for name, rec_codes in view.termination_components.items():
            torch.testing.assert_close(codes[name], rec_codes[0], rtol=0, atol=0)
        torch.testing.assert_close(done, view.done[0])
        torch.testing.assert_close(success, view.success[0])
        # the observation plugins the reference writes against data.* can be restated on the views, e.g. CenterOfMassVelocity
        off, dim, _ = p.obs_slices["center_of_mass_velocity"]
        not_reset = ~view.done[0]
        nxt_obs = rec[1][:, p["TI_R_OBS"] + off : p["TI_R_OBS"] + off + dim]
        torch.testing.assert_close(d.cvel[:, 1:].reshape(n, -1)[not_reset], nxt_obs[not_reset], rtol=0, atol=0)


@pytest.mark.gpu
def test_engine_reset_matches_the_fused_initialisation(walking_spec, walking_randomizers) -> None:
    """``b200sim_engine_reset`` with the reset key of ``b200sim_init_envs`` reproduces the physics part of the initial state."""
    import torch

    from ksim_b200.engine import B200Engine

    n = 64
    eng = B200Engine(walking_spec, n, device="cuda:0", randomizers=walking_randomizers, seed=9, curriculum_level=1.0)
    p = eng.program
    want = eng.state.clone()
    reset_keys = torch.from_numpy(np.ascontiguousarray(eng.env_init.init_keys[:, 0:2]).view(np.int32)).cuda()
    d = eng.engine_reset(reset_keys, torch.from_numpy(eng.env_init.levels).cuda())
    nq, nv = walking_spec.model.nq, walking_spec.model.nv
    for slot, width in (("TI_S_QPOS", nq), ("TI_S_QVEL", nv), ("TI_S_ZERO_OFF", walking_spec.model.nu), ("TI_S_EVENT", p["TI_EVENT_SIZE"]),
                        ("TI_S_ACT_STATE", p["TI_ACT_STATE_SIZE"]), ("TI_S_LATENCY", 1), ("TI_S_TIME", 1)):
        torch.testing.assert_close(eng.state[:, p[slot] : p[slot] + width], want[:, p[slot] : p[slot] + width], rtol=0, atol=0, msg=slot)
    torch.testing.assert_close(d.qpos, eng.state[:, p["TI_S_QPOS"] : p["TI_S_QPOS"] + nq], rtol=0, atol=0)
    assert float(d.qvel.abs().max()) == 0.0 and float(d.time.abs().max()) == 0.0


def _oracle_setup(spec, randomizers, n, seed, level):  # noqa: ANN001, ANN202
    from ksim_b200.envinit import make_env_init
    from ksim_b200.program import build_program
    from oracle import OracleEngine

    prog = build_program(spec)
    init = make_env_init(prog, randomizers, n, seed=seed, level=level)
    ora = OracleEngine(prog, "f32")
    return prog, init, ora


def test_oracle_engine_step_is_its_fused_step_minus_the_task_layer(oracle_built, walking_spec, walking_randomizers) -> None:  # noqa: ANN001
    """CPU: the oracle's engine-only entry points (the checker of the CUDA ones, and the entry MJX fixtures are compared
    through) against its own fused control step."""
    n = 8
    prog, init, ora = _oracle_setup(walking_spec, walking_randomizers, n, 3, 0.5)
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    st2, _, data0 = ora.engine_reset(init.init_keys[:, 0:2], init.levels, init.rand, 2)
    nq = walking_spec.model.nq
    q = prog["TI_S_QPOS"]
    np.testing.assert_array_equal(st2[:, q : q + nq], st[:, q : q + nq])
    np.testing.assert_array_equal(data0[:, :nq], st[:, q : q + nq])
    act = (0.3 * np.random.RandomState(0).randn(n, prog.action_size)).astype(np.float32)
    st_a, keys_a = st.copy(), keys.copy()
    d = ora.engine_step(act, rng.split(keys[:, :2], 3)[:, 1, :], init.rand, st_a, keys_a, 2)
    rec = np.zeros((2, n, prog.rec_size), np.float32)
    rec[0] = rec0
    ora.step_ctrl(act, init.rand, st, keys, rec[0], rec[1])
    o = prog["TI_R_QPOS"]
    np.testing.assert_array_equal(d[:, :nq], rec[0][:, o : o + nq])
    assert d.shape[1] == ora.data_size(2) and np.isfinite(d).all()


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["walking", "kbot"])
def test_engine_entry_points_match_the_oracle(oracle_built, task, request) -> None:  # noqa: ANN001
    """CUDA ``b200sim_engine_reset`` / ``_step`` against the oracle's, whole data block (every ``physics_state.data`` field a
    plugin reads): fp32 tolerance 2e-4 absolute + 2e-4 relative over two control steps from the same keys and actions."""
    import torch

    from ksim_b200.engine import B200Engine

    spec = request.getfixturevalue(f"{task}_spec")
    rz = request.getfixturevalue(f"{task}_randomizers")
    n, level = 48, 1.0 if task == "kbot" else 0.5
    prog, init, ora = _oracle_setup(spec, rz, n, 4, level)
    eng = B200Engine(spec, n, device="cuda:0", randomizers=rz, seed=4, curriculum_level=level)
    keys_r = np.ascontiguousarray(init.init_keys[:, 0:2])
    d = eng.engine_reset(torch.from_numpy(keys_r.view(np.int32)).cuda(), torch.from_numpy(init.levels).cuda())
    ncon = d.contact.geom1.shape[1]
    st, keys, d_o = ora.engine_reset(keys_r, init.levels, init.rand, ncon)
    off = d.offsets

    def compare(got: np.ndarray, want: np.ndarray, what: str) -> None:
        for name in ("QPOS", "QVEL", "QACC", "CTRL", "TIME", "XPOS", "XQUAT", "CINERT", "CVEL", "ACTUATOR_FORCE", "SENSORDATA", "SITE_XPOS",
                     "SITE_XMAT", "SUBTREE_COM", "XFRC_APPLIED", "CFRC_EXT"):
            f = getattr(C, f"B2S_DATA_{name}")
            a, b = got[:, off[f] : off[f + 1]], want[:, off[f] : off[f + 1]]
            scale = max(float(np.abs(b).max()), 1.0)
            bad = np.abs(a - b) > 2e-4 * scale + 2e-4 * np.abs(b)
            # contact forces flip with ulp-level changes of a penetration depth: allow isolated environments to differ there
            # (K-Bot: capsule contacts + friction-loss rows make such isolated flips visible in every field; same clause and hard
            # bound as tests/test_gpu_parity.py)
            limit = 0.05 if name in ("CFRC_EXT", "QACC", "SENSORDATA", "ACTUATOR_FORCE") or task == "kbot" else 0.0
            if name == "QPOS":
                assert np.abs(a - b).max() <= 2e-2, f"{what}: qpos off by {np.abs(a - b).max():.3e}"
            assert bad.any(axis=1).mean() <= limit, f"{what}: {name} differs in {bad.any(axis=1).sum()} of {n} envs (max {np.abs(a - b).max():.3e})"

    compare(d.block.cpu().numpy(), d_o, "reset")
    np.testing.assert_allclose(eng.state.cpu().numpy(), st, rtol=2e-4, atol=2e-4)
    r = np.random.RandomState(1)
    for step in range(2):
        action = (0.3 * r.randn(n, prog.action_size)).astype(np.float32)
        key = r.randint(0, 2**32, size=(n, 2), dtype=np.uint64).astype(np.uint32)
        eng.state.copy_(torch.from_numpy(st).cuda())       # resynchronised: the comparison is per step, not accumulated
        d = eng.engine_step(torch.from_numpy(action).cuda(), torch.from_numpy(key.view(np.int32)).cuda())
        d_o = ora.engine_step(action, key, init.rand, st, keys, ncon)
        compare(d.block.cpu().numpy(), d_o, f"step {step}")
