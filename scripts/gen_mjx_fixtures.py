"""Golden MJX trajectories for ``tests/test_oracle_vs_mjx_fixtures.py``: the UNMODIFIED reference ``MjxEngine`` (ksim/engine.py:
202-361) on the JAX CPU backend, run from recorded keys and actions, dumped to ``tests/golden/mjx_<task>_c1.npz``.

    python scripts/gen_mjx_fixtures.py [--reference /root/reference] [--task humanoid|kbot] [--envs 32] [--steps 24]

C1 of SURVEY.md 8d: humanoid, 32 environments x 24 control steps, curriculum level 0, no per-env randomisation.  What is stored:
the reset key and level of every environment and the ``PhysicsState`` ``engine.reset`` returns; the action and the key of every
``engine.step`` call; after every step the ``data`` fields the plugins read (qpos, qvel, qacc, ctrl, time, xpos, xquat, cvel,
cinert, subtree_com, actuator_force, sensordata, cfrc_ext) and the engine-side state (last_ctrl, actuator state, event states).
Keys are stored as raw uint32 pairs (``jax.random.key_data``), so that the threefry restatement of this repository is pinned by
the same file.

NOT RUN in this repository's image (no jax / mujoco wheels offline): until somebody runs it where ``pip install ksim`` works and
commits the ``.npz``, the fixture test prints ``PARITY UNPINNED`` and skips, and DESIGN.md says so.
"""

from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import mjx_common  # noqa: E402

DATA_FIELDS = ("qpos", "qvel", "qacc", "ctrl", "time", "xpos", "xquat", "cvel", "cinert", "subtree_com", "actuator_force", "sensordata", "cfrc_ext")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--task", default="humanoid", choices=["humanoid", "kbot"])
    ap.add_argument("--envs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--level", type=float, default=0.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    why = mjx_common.try_imports()
    if why is not None:
        print(f"MJX unavailable: {why}")
        sys.exit(0)
    import jax
    import jax.numpy as jnp
    import numpy as np

    task, mjx_model, engine = mjx_common.build(args.reference, args.task)
    reset, step = mjx_common.batched(engine, mjx_model)
    n, t = args.envs, args.steps
    level = jnp.full((n,), args.level)
    key = jax.random.PRNGKey(0)
    key, rk, ak, sk = jax.random.split(key, 4)
    reset_keys = jax.random.split(rk, n)
    step_keys = jax.random.split(sk, t * n).reshape(t, n, -1)
    actions = 0.3 * jax.random.normal(ak, (t, n, int(mjx_model.nu)))
    state = reset(level, reset_keys)

    def leaves(prefix: str, tree) -> dict[str, np.ndarray]:  # noqa: ANN001
        flat, _ = jax.tree_util.tree_flatten_with_path(tree)
        return {prefix + jax.tree_util.keystr(p): np.asarray(v) for p, v in flat}

    def snapshot(st) -> dict[str, np.ndarray]:  # noqa: ANN001
        out = {f"data.{f}": np.asarray(getattr(st.data, f)) for f in DATA_FIELDS}
        out["last_ctrl"] = np.asarray(st.most_recent_action if hasattr(st, "most_recent_action") else st.last_ctrl)
        out["action_latency"] = np.asarray(st.action_latency)
        out["zero_offset"] = np.asarray(st.zero_offset) if getattr(st, "zero_offset", None) is not None else np.zeros(0)
        out.update(leaves("actuator_state", st.actuator_state))
        out.update(leaves("event_states", st.event_states))
        return out

    raw = lambda k: np.asarray(jax.random.key_data(k) if jnp.issubdtype(k.dtype, jax.dtypes.prng_key) else k, dtype=np.uint32)  # noqa: E731
    store: dict[str, np.ndarray] = {"reset_keys": raw(reset_keys), "step_keys": raw(step_keys), "actions": np.asarray(actions),
                                    "level": np.asarray(level), "n_envs": np.array(n), "n_steps": np.array(t)}
    for k, v in snapshot(state).items():
        store[f"reset/{k}"] = v
    per_step: dict[str, list[np.ndarray]] = {}
    for s in range(t):
        state = step(actions[s], state, level, step_keys[s])
        for k, v in snapshot(state).items():
            per_step.setdefault(k, []).append(v)
    for k, v in per_step.items():
        store[f"step/{k}"] = np.stack(v)
    store["versions"] = np.array([f"jax {jax.__version__}", f"mujoco {__import__('mujoco').__version__}"])
    out = Path(args.out) if args.out else ROOT / "tests" / "golden" / f"mjx_{args.task}_c1.npz"
    np.savez_compressed(out, **store)
    print(f"wrote {out} ({out.stat().st_size / 1e6:.2f} MB): {n} envs x {t} steps, level {args.level}")


if __name__ == "__main__":
    main()
