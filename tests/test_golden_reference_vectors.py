"""The oracle against every known-answer vector the reference's own tests hold for this path."""

import ctypes
import json
from pathlib import Path

import numpy as np
import pytest

from ksim_b200 import rewards, terminations
from ksim_b200.program import build_program
from tests.helpers import synthetic_records, tripod_spec

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_action_penalties_reference_vectors(oracle_built: None, precision: str) -> None:
    from oracle import OracleEngine

    cases = json.loads((GOLDEN / "reward_vectors.json").read_text())["cases"]
    assert len(cases) == 9
    for case in cases:
        cls = getattr(rewards, case["reward"])
        prog = build_program(tripod_spec(rewards={"pen": cls(scale=1.0, norm=case["norm"])}))
        ora = OracleEngine(prog, precision)
        actions = np.array(case["actions"], dtype=np.float64)
        assert actions.shape[1] == prog.action_size == 3
        rec = synthetic_records(prog, actions, np.array(case["done"], dtype=np.float64), ora.real)
        carry = np.zeros((1, max(prog["TI_REW_CARRY_SIZE"], 1)), dtype=ora.real)
        total, comps = ora.rewards(rec, np.zeros(1), carry)
        # the engine applies the Penalty sign flip (rewards.py:118-124, rl.py:219-221); the vectors are pre-flip
        np.testing.assert_allclose(-comps[0, 0], case["expected"], rtol=0, atol=1e-5, err_msg=case["ref"])
        np.testing.assert_allclose(total[0], comps[0, 0], rtol=0, atol=0)


def _term_from_case(case: dict) -> terminations.Termination:
    cls = getattr(terminations, case["termination"])
    return cls(**case["args"])


def test_termination_truth_tables(oracle_built: None) -> None:
    import oracle

    data = json.loads((GOLDEN / "termination_vectors.json").read_text())
    qpos = np.array(data["qpos"], dtype=np.float64)
    for precision in ("f32", "f64"):
        lib = oracle.lib(precision)
        real = np.float32 if precision == "f32" else np.float64
        creal = ctypes.c_float if precision == "f32" else ctypes.c_double
        for case in data["cases"]:
            prog = build_program(tripod_spec(terminations={"t": _term_from_case(case)}))
            ti = np.ascontiguousarray(prog.ti, dtype=np.int32)
            tf = np.ascontiguousarray(prog.tf.astype(real))
            q, v = qpos.astype(real), np.zeros(6, dtype=real)
            con = data["contacts"] if case["contacts"] else {"geom1": [], "geom2": [], "dist": []}
            g1, g2 = np.array(con["geom1"], dtype=np.int32), np.array(con["geom2"], dtype=np.int32)
            dist = np.array(con["dist"], dtype=real)
            got = lib.o_test_termination(ctypes.c_void_p(ti.ctypes.data), ctypes.c_void_p(tf.ctypes.data), 0, ctypes.c_void_p(q.ctypes.data),
                                         q.size, ctypes.c_void_p(v.ctypes.data), v.size, creal(0.0), creal(0.0), g1.size,
                                         ctypes.c_void_p(g1.ctypes.data), ctypes.c_void_p(g2.ctypes.data), ctypes.c_void_p(dist.ctypes.data))
            assert got == case["expected"], case["ref"]


def test_termination_name_rule() -> None:
    # tests/test_terminations.py:55-59 and :126-131
    assert terminations.BadZTermination(min_z=0.1, max_z=1.0).termination_name == "bad_z_termination"
    model = tripod_spec().model
    term = terminations.IllegalContactTermination.create(model, geom_names=["foot_a", "foot_b"])
    assert term.termination_name == "illegal_contact_termination" and len(term.illegal_geom_idxs) == 2
    with pytest.raises(ValueError):
        terminations.IllegalContactTermination.create(model, geom_names=["nope"])


def test_float_vector_command_properties(oracle_built: None) -> None:
    """tests/test_commands.py:66-96: shape, bounds, and identity when switch_prob = 0 -- through a full env step."""
    from ksim_b200.envinit import make_env_init
    from oracle import OracleEngine

    prog = build_program(tripod_spec())
    ora = OracleEngine(prog, "f32")
    n = 100
    init = make_env_init(prog, {}, n, seed=3)
    state, keys, rec0, _ = ora.init_envs(init.init_keys, init.levels, init.rand)
    c0 = prog["TI_S_CMD"]
    cmd0 = state[:, c0 : c0 + 2].copy()
    assert cmd0.shape == (n, 2) and (cmd0 >= 0).all() and (cmd0 <= 1).all() and cmd0.std() > 0.1
    rec = np.zeros((2, n, prog.rec_size), dtype=np.float32)
    rec[0] = rec0
    ora.step_ctrl(np.zeros((n, prog.action_size), dtype=np.float32), init.rand, state, keys, rec[0], rec[1])
    np.testing.assert_array_equal(state[:, c0 : c0 + 2], cmd0)  # switch_prob = 0 -> unchanged
