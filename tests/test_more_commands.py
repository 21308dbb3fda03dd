"""``IntVectorCommand``, ``StartPositionCommand``, ``StartQuaternionCommand``, ``BaseHeightCommand`` (ksim/commands.py:137-214,
773-796).  CPU: the oracle's command blocks against a numpy transcription of the reference's expressions on this repo's
threefry restatement, driven by the environment keys the oracle exposes before every step (key discipline of
``RLTask.step_engine``, rl.py:1034, 1080-1098); host emulation of the kernel code against the oracle.  GPU: CUDA == oracle,
bit-exact (commands are RNG and copies only)."""

import numpy as np
import pytest

from ksim_b200 import commands, rng as R, terminations
from ksim_b200.envinit import make_env_init
from ksim_b200.program import build_program
from tests.helpers import tripod_spec
from tests.test_emu_vs_oracle import _p, emu  # noqa: F401

RANGES = ((-2, 3), (0, 0), (5, 9))


def _spec():  # noqa: ANN202
    cmds = {"ints": commands.IntVectorCommand(ranges=RANGES, switch_prob=0.3, zero_prob=0.2),
            "start_pos": commands.StartPositionCommand(), "start_quat": commands.StartQuaternionCommand(),
            "height": commands.BaseHeightCommand(min_height=0.4, max_height=0.9, switch_prob=0.25),
            "vec": commands.FloatVectorCommand(ranges=((0.0, 1.0),), switch_prob=0.5)}
    return tripod_spec(cmds=cmds, terminations={"episode_length": terminations.EpisodeLengthTermination(max_length_sec=0.11)})


def _randint(key: np.ndarray, lo: np.ndarray, hi_incl: np.ndarray) -> np.ndarray:
    """jax.random.randint(key, (n,), lo, hi_incl + 1) for int32 (jax/_src/random.py: two 32-bit draws + modular arithmetic)."""
    k1, k2 = R.split(key, 2)
    n = len(lo)
    span = (hi_incl + 1 - lo).astype(np.uint32)
    mult = ((np.uint32(65536) % span) * (np.uint32(65536) % span)) % span
    off = ((R.bits(k1, n) % span) * mult + R.bits(k2, n) % span) % span
    return lo + off.astype(np.int32)


def _int_initial(key: np.ndarray) -> np.ndarray:
    lo, hi = np.array([r[0] for r in RANGES], np.int32), np.array([r[1] for r in RANGES], np.int32)
    zero = R.bernoulli(key, 1, 0.2)[0]
    return np.where(zero, 0.0, _randint(key, lo, hi)).astype(np.float32)


def _rollout_with_keys(spec, n, t, seed=7):  # noqa: ANN001, ANN202
    from oracle import OracleEngine

    prog = build_program(spec)
    init = make_env_init(prog, {}, n, seed=seed, level=1.0)
    ora = OracleEngine(prog, "f32")
    st, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    rec = np.zeros((t + 1, n, prog.rec_size), np.float32)
    rec[0] = rec0
    actions = (0.3 * np.random.RandomState(seed).randn(t, n, prog.action_size)).astype(np.float32)
    env_keys, states = [], []
    for s in range(t):
        env_keys.append(keys[:, :2].copy())
        ora.step_ctrl(actions[s], init.rand, st, keys, rec[s], rec[s + 1])
        states.append(st.copy())
    return prog, init, rec, actions, np.stack(env_keys), np.stack(states)


def test_oracle_matches_numpy_restatement(oracle_built) -> None:  # noqa: ANN001
    spec = _spec()
    n, t = 24, 14
    prog, init, rec, _, env_keys, states = _rollout_with_keys(spec, n, t)
    cm = prog["TI_R_CMD"]
    sl = {k: slice(cm + o, cm + o + d) for k, (o, d) in prog.cmd_slices.items()}
    names = list(spec.commands)
    done = rec[:t, :, prog["TI_R_DONE"]] != 0
    assert done.any()
    qs = prog["TI_S_QPOS"]
    n_switch = 0
    for e in range(n):
        # initial commands: chain over the commands from the env's command key (rl.py:2231-2313 / get_initial_commands)
        key = init.init_keys[e, 2:4]
        cur = {}
        for name in names:
            key, k = R.split(key, 2)
            cur[name] = k
        np.testing.assert_array_equal(rec[0, e, sl["ints"]], _int_initial(cur["ints"]))
        np.testing.assert_array_equal(rec[0, e, sl["height"]], R.uniform(cur["height"], 1, 0.4, 0.9))
        prev = {k: rec[0, e, v].copy() for k, v in sl.items()}
        for s in range(t):
            state_rng = R.split(env_keys[s, e], 3)[2]
            key = R.split(state_rng, 6)[0]
            for name in names:
                key, k = R.split(key, 2)
                got = rec[s + 1, e, sl[name]]
                if done[s, e]:   # get_initial_commands on the reset state
                    want = {"ints": lambda: _int_initial(k), "height": lambda: R.uniform(k, 1, 0.4, 0.9),
                            "start_pos": lambda: states[s, e, qs : qs + 3], "start_quat": lambda: states[s, e, qs + 3 : qs + 7]}.get(name)
                    if want is not None:
                        np.testing.assert_array_equal(got, want(), err_msg=f"{name} after reset, env {e} step {s}")
                elif name == "ints":
                    ra, rb = R.split(k, 2)
                    sw = R.bernoulli(ra, 1, 0.3)[0]
                    n_switch += int(sw)
                    np.testing.assert_array_equal(got, _int_initial(rb) if sw else prev[name])
                elif name == "height":
                    fresh = R.uniform(k, 1, 0.4, 0.9)
                    np.testing.assert_array_equal(got, fresh if R.bernoulli(k, 1, 0.25)[0] else prev[name])
                elif name in ("start_pos", "start_quat"):
                    np.testing.assert_array_equal(got, prev[name])
                prev[name] = got.copy()
    assert n_switch > 10
    ints = rec[:, :, sl["ints"]]
    assert set(np.unique(ints[..., 1]).tolist()) == {0.0} and ints[..., 0].min() >= -2 and ints[..., 0].max() <= 3
    assert (ints[..., 2] == 0).any() and ints[..., 2].max() == 9 and (np.round(ints) == ints).all()


def test_emu_matches_oracle(emu, oracle_built) -> None:  # noqa: ANN001, F811
    from ksim_b200.blob import pack_blob

    spec = _spec()
    n, t = 8, 12
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
    o, d = prog.cmd_slices["start_pos"]
    keep = np.ones(cs, bool)
    keep[o : o + d] = False
    o2, d2 = prog.cmd_slices["start_quat"]
    keep[o2 : o2 + d2] = False
    np.testing.assert_array_equal(rec_e[:, :, cm : cm + cs][..., keep], rec_o[:, :, cm : cm + cs][..., keep])
    np.testing.assert_allclose(rec_e[:, :, cm : cm + cs], rec_o[:, :, cm : cm + cs], rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_cuda_matches_oracle(oracle_built) -> None:  # noqa: ANN001
    import torch

    from ksim_b200.engine import B200Engine

    spec = _spec()
    n, t = 48, 14
    prog, init, rec_o, actions, _, _ = _rollout_with_keys(spec, n, t)
    eng = B200Engine(spec, n, device="cuda:0", seed=7, curriculum_level=1.0)
    rec = eng.rollout(torch.from_numpy(actions).cuda()).cpu().numpy()
    cm, cs = prog["TI_R_CMD"], prog["TI_CMD_SIZE"]
    np.testing.assert_array_equal(rec[:t, :, prog["TI_R_DONE"]], rec_o[:t, :, prog["TI_R_DONE"]])
    keep = np.ones(cs, bool)
    for name in ("start_pos", "start_quat"):
        o, d = prog.cmd_slices[name]
        keep[o : o + d] = False
    np.testing.assert_array_equal(rec[:, :, cm : cm + cs][..., keep], rec_o[:, :, cm : cm + cs][..., keep])
    np.testing.assert_allclose(rec[:, :, cm : cm + cs], rec_o[:, :, cm : cm + cs], rtol=0, atol=1e-5)
