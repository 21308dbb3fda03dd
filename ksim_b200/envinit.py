"""Host-side initialisation of the per-environment key tree and model randomizations.

Follows ``RLTask._get_env_state`` (ksim/task/rl.py:2175-2272): one root key is split nine ways, each purpose key
is split per environment, and ``apply_randomizations`` (rl.py:361-372) derives the randomizer and reset keys.
Runs once per training run, in numpy; the keys then live on the GPU.
"""

from __future__ import annotations

__all__ = ["EnvInit", "make_env_init"]

from dataclasses import dataclass
from typing import Mapping

import numpy as np

from ksim_b200 import rng as R
from ksim_b200.program import TaskProgram
from ksim_b200.randomization import PhysicsRandomizer, pair_friction_unaffected


@dataclass
class EnvInit:
    init_keys: np.ndarray      # (N, 10) uint32: reset, command, obs-carry, obs-bias, env rng
    rand: np.ndarray           # (N, RAND) float32 per-env model overrides
    levels: np.ndarray         # (N,) float32 curriculum level
    model_carry_keys: np.ndarray  # (N, 2) uint32 (for the policy's initial carry; host framework)
    reward_keys: np.ndarray    # (N, 2) uint32
    randomizations: dict[str, np.ndarray]


def make_env_init(program: TaskProgram, randomizers: Mapping[str, PhysicsRandomizer], num_envs: int, seed: int = 0,
                  level: float = 0.0, env_offset: int = 0, total_envs: int | None = None) -> EnvInit:
    """Keys for environments ``[env_offset, env_offset + num_envs)`` of a job with ``total_envs`` environments.

    Sharding a job across GPUs only changes ``env_offset``: environment ``e`` gets the same keys whatever the
    number of ranks, so results are invariant to the GPU count.
    """
    total = num_envs if total_envs is None else total_envs
    model = program.spec.model
    root = R.prng_key(seed)
    (_, carry_model, carry_obs, carry_obs_bias, command, rand_rng, rollout, _curriculum, reward) = R.split(root, 9)
    sl = slice(env_offset, env_offset + num_envs)

    def per_env(key: np.ndarray) -> np.ndarray:
        return R.split(key, total)[sl]

    rand_env = per_env(rand_rng)                 # apply_randomizations rng per env
    two = R.split(rand_env, 2)                   # (N, 2, 2): rand_rng, reset_rng
    rkey, reset_key = two[:, 0], two[:, 1]
    randomizations: dict[str, np.ndarray] = {}
    for name, randomizer in randomizers.items():
        pair = R.split(rkey, 2)                  # rng, randomization_rng = split(rng)
        rkey, sub = pair[:, 0], pair[:, 1]
        out = randomizer(model, sub)
        for k in out:
            if k in randomizations:
                raise ValueError(f"Found duplicate randomization keys: {k}")
        randomizations.update(out)
    if "geom_friction" in randomizations:
        # contact pairs carry their own (compile-time mixed) friction: only a no-op override is supported (FloorFrictionRandomizer)
        if not pair_friction_unaffected(model, randomizations["geom_friction"]):
            raise NotImplementedError("per-env geom_friction that changes a contact pair's friction is not supported by the CUDA engine")
        del randomizations["geom_friction"]
    if set(randomizations) != set(program.rand_slices):
        raise ValueError(f"randomizers produce {sorted(randomizations)}, the program expects {sorted(program.rand_slices)}")
    rand = np.zeros((num_envs, program.rand_size), dtype=np.float32)
    for fname, (off, cnt) in program.rand_slices.items():
        rand[:, off : off + cnt] = randomizations[fname].reshape(num_envs, cnt)
    init_keys = np.concatenate([reset_key, per_env(command), per_env(carry_obs), per_env(carry_obs_bias), per_env(rollout)],
                               axis=1).astype(np.uint32)
    return EnvInit(init_keys=np.ascontiguousarray(init_keys), rand=rand,
                   levels=np.full(num_envs, level, dtype=np.float32), model_carry_keys=per_env(carry_model),
                   reward_keys=per_env(reward), randomizations=randomizations)
