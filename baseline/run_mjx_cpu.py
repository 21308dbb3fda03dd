"""Host-CPU MJX baseline (BASELINE.md section 3 / SURVEY.md 8d): the reference's own ``MjxEngine`` on ``JAX_PLATFORMS=cpu``, same
workload shape as ``bench.py`` (N humanoid environments x T control steps of 5 physics sub-steps, random-init-policy surrogate
actions ``0.3 N(0, 1)``), wall-clock after compilation, all host cores.  Prints ONE JSON line; when the reference's stack is not
importable it prints ``{"impl": "mjx", "unavailable": "<why>"}`` and exits 0.

    python baseline/run_mjx_cpu.py [--reference /root/reference] [--task humanoid|kbot] [--envs 1024] [--steps 64] [--reps 3]

NOT RUN in this repository's image (no jax / mujoco wheels offline); see baseline/mjx_common.py.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import mjx_common  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--task", default="humanoid", choices=["humanoid", "kbot"])
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    why = mjx_common.try_imports()
    if why is None and not Path(args.reference, "ksim").is_dir():
        why = f"no reference checkout at {args.reference}"
    if why is not None:
        print(json.dumps({"impl": "mjx", "unavailable": why}))
        return
    import jax
    import jax.numpy as jnp

    task, mjx_model, engine = mjx_common.build(args.reference, args.task)
    reset, step = mjx_common.batched(engine, mjx_model)
    n, t = args.envs, args.steps
    level = jnp.full((n,), 1.0 if args.task == "kbot" else 0.0)
    key = jax.random.PRNGKey(0)
    key, rk, ak = jax.random.split(key, 3)
    state = reset(level, jax.random.split(rk, n))
    nu = int(mjx_model.nu)
    actions = 0.3 * jax.random.normal(ak, (t, n, nu))

    @jax.jit
    def rollout(state, key):  # noqa: ANN001, ANN202
        def body(carry, a):  # noqa: ANN001, ANN202
            st, k = carry
            k, sk = jax.random.split(k)
            st = step(a, st, level, jax.random.split(sk, n))
            return (st, k), st.data.qpos[:, 2]
        (state, key), z = jax.lax.scan(body, (state, key), actions)
        return state, z

    state, z = rollout(state, key)          # compile + warm-up
    jax.block_until_ready(z)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        state, z = rollout(state, key)
        jax.block_until_ready(z)
    dt = (time.perf_counter() - t0) / args.reps
    cores = os.cpu_count() or 1
    print(json.dumps({"impl": "mjx", "metric": f"{args.task} env-steps/sec", "value": n * t / dt, "unit": "env-steps/s", "cores": cores,
                      "per_core": n * t / dt / cores, "envs": n, "steps": t, "reps": args.reps, "seconds_per_rollout": dt,
                      "backend": jax.default_backend(), "jax": jax.__version__,
                      "what": "MjxEngine.step vmapped over envs and scanned over time (engine only: no observations / rewards)"}))


if __name__ == "__main__":
    main()
