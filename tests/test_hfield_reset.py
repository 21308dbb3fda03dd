"""``HFieldXYPositionReset`` (ksim/resets.py:45-89) and the ``get_xy_position_reset`` factory (210-280).

CPU: the oracle's reset positions against a numpy transcription of the reference's expressions on this repo's threefry
restatement (same keys), and the host emulation of the kernel code against the oracle.  GPU: CUDA == oracle."""

import types

import numpy as np
import pytest

from ksim_b200 import resets, terminations
from ksim_b200.envinit import make_env_init
from ksim_b200.program import build_program
from tests.helpers import tripod_model, tripod_spec
from tests.test_emu_vs_oracle import _p, emu  # noqa: F401

NX, NY = 7, 5
BOUNDS = (3.0, 2.0, 0.25, 0.1)   # x_bound, y_bound, z_top, z_bottom


def _heights() -> np.ndarray:
    i, j = np.meshgrid(np.arange(NX), np.arange(NY), indexing="ij")
    return (0.05 * np.sin(1.3 * i) + 0.03 * j).astype(np.float32)


def _spec():  # noqa: ANN202
    model = tripod_model()
    rs = [resets.RandomJointPositionReset.create(model),
          resets.HFieldXYPositionReset(bounds=BOUNDS, padded_bounds=(-2.5, 2.5, -1.5, 1.5), x_range=4.0, y_range=1.0,
                                       hfield_data=_heights(), robot_base_height=0.4)]
    spec = tripod_spec(terminations={"episode_length": terminations.EpisodeLengthTermination(max_length_sec=0.09)})
    spec.resets = rs
    return spec


def _want_z(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """resets.py:72-87 in float32."""
    xb, yb, z_top, _ = (np.float32(v) for v in BOUNDS)
    xi = np.clip((((x + xb) / (np.float32(2) * xb)) * np.float32(NX - 1)).astype(np.int32), 0, NX - 1)
    yi = np.clip((((y + yb) / (np.float32(2) * yb)) * np.float32(NY - 1)).astype(np.int32), 0, NY - 1)
    return _heights()[xi, yi] + z_top + np.float32(0.4)


def _oracle(spec, n, t, seed=5):  # noqa: ANN001, ANN202
    from oracle import OracleEngine

    prog = build_program(spec)
    init = make_env_init(prog, {}, n, seed=seed, level=1.0)
    ora = OracleEngine(prog, "f32")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    actions = np.zeros((t, n, prog.action_size), np.float32)
    st0, k0 = st.copy(), keys.copy()
    ora.rollout(actions, init.rand, st, keys, rec)
    return prog, init, st0, k0, st, rec, actions


def test_reset_lands_on_the_height_field(oracle_built) -> None:  # noqa: ANN001
    spec = _spec()
    n, t = 64, 10
    prog, init, st0, _, st, rec, _ = _oracle(spec, n, t)
    q = prog["TI_S_QPOS"]
    x, y, z = st0[:, q], st0[:, q + 1], st0[:, q + 2]
    assert (np.abs(x) <= 2.5).all() and (np.abs(y) <= 1.0).all() and np.unique(x).size > n // 2
    assert (x == 2.5).any() or (x == -2.5).any()   # x_range 4 > padded bound 2.5: the clip is exercised
    np.testing.assert_array_equal(z, _want_z(x, y))
    assert np.unique(z).size > 5
    # reset-on-done inside the rollout lands on the field as well (state after the last step's reset)
    done = rec[:t, :, prog["TI_R_DONE"]] != 0
    assert done[-1].any()
    e = np.nonzero(done[-1])[0]
    np.testing.assert_array_equal(st[e, q + 2], _want_z(st[e, q], st[e, q + 1]))


def test_factory_picks_the_height_field_when_the_model_has_one() -> None:
    model = tripod_model()
    plane = resets.get_xy_position_reset(model, x_range=1.0, y_range=1.0, x_edge_padding=1.0, y_edge_padding=1.0)
    assert isinstance(plane, resets.PlaneXYPositionReset)
    fake = types.SimpleNamespace(hfield_size=np.array([[3.0, 2.0, 0.25, 0.1]]), hfield_nrow=np.array([NX]), hfield_ncol=np.array([NY]),
                                 hfield_data=_heights().reshape(-1), geom_type=model.geom_type, geom_pos=model.geom_pos)
    hf = resets.get_xy_position_reset(fake, x_range=4.0, y_range=1.0, x_edge_padding=0.5, y_edge_padding=0.5, robot_base_height=0.4)
    assert isinstance(hf, resets.HFieldXYPositionReset)
    assert hf.bounds == (3.0, 2.0, 0.25, pytest.approx(0.1)) and hf.padded_bounds == (-2.5, 2.5, -1.5, 1.5)
    code, ints, floats = hf.encode(model)
    assert ints == [NX, NY] and len(floats) == 10 + NX * NY


def test_emu_matches_oracle(emu, oracle_built) -> None:  # noqa: ANN001, F811
    from ksim_b200.blob import pack_blob

    spec = _spec()
    n, t = 8, 8
    prog, init, st0, k0, _, rec_o, actions = _oracle(spec, n, t)
    blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
    st_e, k_e = np.zeros_like(st0), np.zeros_like(k0)
    rec_e = np.zeros_like(rec_o)
    ak = np.zeros((n, 2), np.uint32)
    emu.emu_init_envs(_p(blob), n, _p(init.init_keys), _p(init.levels), _p(init.rand), _p(st_e), _p(k_e), _p(rec_e[0]), _p(ak))
    np.testing.assert_array_equal(k_e, k0)
    np.testing.assert_allclose(st_e, st0, rtol=0, atol=2e-5)
    q = prog["TI_S_QPOS"]
    np.testing.assert_array_equal(st_e[:, q : q + 3], st0[:, q : q + 3])


@pytest.mark.gpu
def test_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    spec = _spec()
    n, t = 64, 10
    prog, init, st0, k0, st, rec_o, actions = _oracle(spec, n, t)
    eng = B200Engine(spec, n, device="cuda:0", seed=5, curriculum_level=1.0)
    q = prog["TI_S_QPOS"]
    np.testing.assert_array_equal(eng.state.cpu().numpy()[:, q : q + 3], st0[:, q : q + 3])
    rec = eng.rollout(torch.from_numpy(actions).cuda()).cpu().numpy()
    np.testing.assert_array_equal(rec[:t, :, prog["TI_R_DONE"]], rec_o[:t, :, prog["TI_R_DONE"]])
    got = eng.state.cpu().numpy()
    e = np.nonzero(rec_o[t - 1, :, prog["TI_R_DONE"]] != 0)[0]
    assert e.size > 0
    np.testing.assert_array_equal(got[e, q : q + 3], st[e, q : q + 3])


# ------------------------------------------------------------------------------------------ InitialMotionStateReset
def _motion_spec():  # noqa: ANN202
    model = tripod_model()
    frames = 37
    rs = np.random.RandomState(0)
    bank = resets.MotionReferenceData(qpos=rs.uniform(-0.4, 0.4, (frames, model.nq)).astype(np.float32),
                                      qvel=rs.uniform(-1.0, 1.0, (frames, model.nq)).astype(np.float32), ctrl_dt=0.02)
    spec = tripod_spec(terminations={"episode_length": terminations.EpisodeLengthTermination(max_length_sec=100.0)})
    spec.resets = [resets.InitialMotionStateReset(reference_motion=bank)]
    return spec, bank


def test_initial_motion_state_reset(emu, oracle_built) -> None:  # noqa: ANN001, F811
    """resets.py:284-304 against a numpy transcription on the repo's threefry: frame = randint(rng, (1,), 0, T)[0]; joints from the
    bank, base untouched, time = frame * ctrl_dt (the engine zeroes qvel after its forward pass, engine.py:216-219)."""
    from ksim_b200 import rng as R
    from ksim_b200.blob import pack_blob

    spec, bank = _motion_spec()
    n = 48
    prog, init, st0, k0, *_ = _oracle(spec, n, 2)
    q, tm = prog["TI_S_QPOS"], prog["TI_S_TIME"]
    frames = np.zeros(n, np.int64)
    for e in range(n):
        reset_rng = R.split(init.init_keys[e, 0:2], 2)[1]    # rng, reset_rng = split(rng) for the first (only) reset
        k1, k2 = R.split(reset_rng, 2)
        span = np.uint32(bank.num_frames)
        mult = ((np.uint32(65536) % span) * (np.uint32(65536) % span)) % span
        frames[e] = int(((R.bits(k1, 1)[0] % span) * mult + R.bits(k2, 1)[0] % span) % span)
    assert np.unique(frames).size > 15
    np.testing.assert_array_equal(st0[:, q + 7 : q + 10], np.asarray(bank.qpos)[frames, 7:10])
    np.testing.assert_array_equal(st0[:, tm], frames.astype(np.float32) * np.float32(0.02))
    model = spec.model
    np.testing.assert_allclose(st0[:, q : q + 7], np.broadcast_to(model.qpos0[:7].astype(np.float32), (n, 7)), atol=1e-7)
    blob = np.frombuffer(pack_blob(spec.model, prog.ti, prog.tf), dtype=np.uint8).copy()
    st_e, k_e = np.zeros_like(st0), np.zeros_like(k0)
    rec0 = np.zeros((n, prog.rec_size), np.float32)
    ak = np.zeros((n, 2), np.uint32)
    emu.emu_init_envs(_p(blob), n, _p(init.init_keys), _p(init.levels), _p(init.rand), _p(st_e), _p(k_e), _p(rec0), _p(ak))
    np.testing.assert_array_equal(st_e[:, q + 7 : q + 10], st0[:, q + 7 : q + 10])
    np.testing.assert_array_equal(st_e[:, tm], st0[:, tm])
    with pytest.raises(NotImplementedError):
        resets.InitialMotionStateReset(reference_motion=bank, freejoint=True).encode(model)


@pytest.mark.gpu
def test_initial_motion_state_reset_cuda(oracle_built) -> None:  # noqa: ANN001
    from ksim_b200.engine import B200Engine

    spec, _ = _motion_spec()
    n = 48
    prog, init, st0, *_ = _oracle(spec, n, 2)
    eng = B200Engine(spec, n, device="cuda:0", seed=5, curriculum_level=1.0)
    q, tm = prog["TI_S_QPOS"], prog["TI_S_TIME"]
    got = eng.state.cpu().numpy()
    np.testing.assert_array_equal(got[:, q : q + 10], st0[:, q : q + 10])
    np.testing.assert_array_equal(got[:, tm], st0[:, tm])
