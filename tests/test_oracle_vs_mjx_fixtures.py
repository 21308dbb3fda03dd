"""The oracle against trajectories of the UNMODIFIED reference ``MjxEngine`` (``scripts/gen_mjx_fixtures.py``).

The fixtures cannot be minted in this repository's image (no jax / mujoco-mjx offline), so until ``tests/golden/mjx_*_c1.npz``
is committed from a machine that has them this test prints ``PARITY UNPINNED`` and skips: physics parity then rests on the
analytic pins of ``tests/test_oracle_physics.py`` and on CUDA == oracle, which is what DESIGN.md section 2 states.

With a fixture present: ``engine.reset`` from the recorded keys must reproduce the recorded ``PhysicsState`` (qpos / qvel / time
and the engine-side draws: zero offsets, latency, event timers -- this pins the threefry / jax.random restatement and the Reset
plugins), and chained ``engine.step`` calls from the recorded actions and keys must stay within the short-horizon tolerance of
``BASELINE.json``'s north_star (fp32: 1e-4 absolute + 1e-4 relative on qpos / qvel over the first 3 control steps, then per-step
comparison after resynchronising the oracle's qpos / qvel to the fixture -- contact dynamics are chaotic over longer horizons).
"""

from pathlib import Path

import numpy as np
import pytest

from ksim_b200 import codes as C

GOLDEN = Path(__file__).resolve().parent / "golden"
TASKS = {"humanoid": ("walking", 0.0), "kbot": ("kbot", 1.0)}


@pytest.mark.parametrize("task", sorted(TASKS))
def test_oracle_matches_mjx_fixture(oracle_built, task, request) -> None:  # noqa: ANN001
    path = GOLDEN / f"mjx_{task}_c1.npz"
    if not path.exists():
        print(f"PARITY UNPINNED: {path.name} has not been generated (scripts/gen_mjx_fixtures.py needs jax + mujoco-mjx)")
        pytest.skip(f"PARITY UNPINNED: no MJX fixture for {task}")
    from ksim_b200.program import build_program
    from oracle import OracleEngine

    fx = np.load(path)
    spec = request.getfixturevalue(f"{TASKS[task][0]}_spec")
    prog = build_program(spec)
    ora = OracleEngine(prog, "f32")
    n, t = int(fx["n_envs"]), int(fx["n_steps"])
    m = spec.model
    rand = np.zeros((n, max(prog.rand_size, 1)), np.float32)
    for name, sl in prog.rand_slices.items():               # identity overrides: the fixture uses the shared base model
        rand[:, sl[0] : sl[0] + sl[1]] = np.asarray(m.arrays[name], dtype=np.float32).reshape(-1)[: sl[1]]
    ncon = m.ncon if hasattr(m, "ncon") else 2
    state, keys, data = ora.engine_reset(fx["reset_keys"], fx["level"].astype(np.float32), rand, ncon)
    off = np.cumsum([0, m.nq, m.nv, m.nv, m.nu, 1])

    def close(got, want, what, atol=1e-4, rtol=1e-4):  # noqa: ANN001, ANN202
        np.testing.assert_allclose(got, np.asarray(want, dtype=np.float64).reshape(got.shape), atol=atol, rtol=rtol, err_msg=what)

    close(data[:, off[0] : off[1]], fx["reset/data.qpos"], "reset qpos")
    close(data[:, off[1] : off[2]], fx["reset/data.qvel"], "reset qvel")
    if fx["reset/zero_offset"].size:
        z = prog["TI_S_ZERO_OFF"]
        close(state[:, z : z + m.nu], fx["reset/zero_offset"], "zero offsets (jax.random restatement + reset discipline)")
    close(state[:, prog["TI_S_LATENCY"]], fx["reset/action_latency"], "action latency")
    qp, qv = prog["TI_S_QPOS"], prog["TI_S_QVEL"]
    for s in range(t):
        data = ora.engine_step(fx["actions"][s], fx["step_keys"][s], rand, state, keys, ncon)
        tol = 1e-4 if s < 3 else 5e-4
        close(data[:, off[0] : off[1]], fx["step/data.qpos"][s], f"qpos after step {s}", atol=tol, rtol=tol)
        close(data[:, off[1] : off[2]], fx["step/data.qvel"][s], f"qvel after step {s}", atol=10 * tol, rtol=10 * tol)
        close(data[:, off[3] : off[4]], fx["step/data.ctrl"][s], f"ctrl after step {s}", atol=10 * tol, rtol=10 * tol)
        if s >= 2:   # resynchronise: from here on the comparison is per step
            state[:, qp : qp + m.nq] = fx["step/data.qpos"][s]
            state[:, qv : qv + m.nv] = fx["step/data.qvel"][s]
    assert C.B2S_DATA_QPOS == 0 and C.B2S_DATA_CTRL == 3   # the block prefix this test slices
