"""Shared set-up of the two scripts that run the UNMODIFIED reference (kscalelabs/ksim ``MjxEngine``, ksim/engine.py:202-361) on
the JAX CPU backend: ``baseline/run_mjx_cpu.py`` (the host-CPU MJX baseline of BASELINE.md section 3) and
``scripts/gen_mjx_fixtures.py`` (golden trajectories for ``tests/test_oracle_vs_mjx_fixtures.py``).

Neither jax, mujoco nor mujoco-mjx exists in this repository's build image or on its GPU boxes (``pip download --no-index
--find-links /opt/wheelhouse jax mujoco mujoco-mjx``: "No matching distribution found for jax", recorded in DESIGN.md), so these
scripts have NOT been executed here; they are written against the reference's public API and are meant to be run once on any
machine where ``pip install ksim`` works.  Nothing in ``tests/``, ``bench.py`` or ``__graft_entry__`` imports this module.
"""

from __future__ import annotations

import os
import sys
from pathlib import Path


def try_imports() -> str | None:
    """None when the reference's stack is importable, else a one-line reason."""
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    for mod in ("jax", "mujoco", "mujoco.mjx", "xax", "equinox"):
        try:
            __import__(mod)
        except Exception as e:  # noqa: BLE001
            return f"{mod} is not importable ({type(e).__name__}: {e})"
    return None


def build(reference_root: str, task: str = "humanoid"):  # noqa: ANN201
    """(task object, mjx model, engine) exactly as ``RLTask`` builds them (ksim/task/rl.py:715-785, 2520-2560): the example's own
    task class and config, ``get_mujoco_model`` -> ``set_mujoco_model_opts`` -> metadata -> ``update_mj_model`` ->
    ``get_mjx_model`` -> ``get_engine``."""
    root = Path(reference_root).resolve()
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    if task == "humanoid":
        import math

        from examples.walking import WalkingConfig, WalkingTask

        # the overrides of the example's __main__ (examples/walking.py:772-788)
        cfg = WalkingConfig(num_envs=4096, batch_size=512, num_passes=4, rollout_length_frames=24, dt=0.004, ctrl_dt=0.02,
                            zero_offset_std=math.radians(3.0), zero_offset_mag=math.radians(3.0), iterations=8, ls_iterations=8)
        t = WalkingTask(cfg)
    elif task == "kbot":
        from examples.kbot.train import HumanoidWalkingTask, HumanoidWalkingTaskConfig

        t = HumanoidWalkingTask(HumanoidWalkingTaskConfig())
    else:
        raise ValueError(task)
    mj_model = t.get_mujoco_model()
    mj_model = t.set_mujoco_model_opts(mj_model)
    metadata = t.get_mujoco_model_metadata(mj_model)
    t.update_mj_model(mj_model, metadata)
    mjx_model = t.get_mjx_model(mj_model)
    engine = t.get_engine(mjx_model, metadata)
    return t, mjx_model, engine


def batched(engine, mjx_model):  # noqa: ANN001, ANN201
    """``engine.reset`` / ``engine.step`` vmapped over environments the way ``_single_unroll`` does (rl.py:1615): the model is
    shared (no per-env randomisation in these scripts), everything else is batched."""
    import jax

    reset = jax.jit(jax.vmap(lambda level, key: engine.reset(mjx_model, level, key)))
    step = jax.jit(jax.vmap(lambda action, state, level, key: engine.step(action, mjx_model, state, level, key)))
    return reset, step
