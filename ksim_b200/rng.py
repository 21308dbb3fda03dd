"""Host-side (numpy) restatement of ``jax.random``'s threefry2x32 key discipline.

Used only at initialisation time: per-env keys for ``init_envs`` (the split tree of
``RLTask._get_env_state``, ksim/task/rl.py:2175-2272) and the physics randomizers
(ksim/randomization.py), which run once per environment.  The device kernels carry their own
implementation of the same functions; tests check the two (and the oracle's) against each other and
against the Random123 known-answer vectors.

Counter layout is JAX's ``jax_threefry_partitionable`` one (the default in current JAX): element ``i`` of a
draw uses the 64-bit counter ``i`` split into (hi, lo) words; ``split(key, n)[i]`` is the two output words,
``bits`` is their xor.  [restated from memory of the un-vendored, unpinned ``jax`` package]
"""

from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key: np.ndarray, x0: np.ndarray, x1: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """key: (..., 2) uint32; x0, x1 broadcastable uint32 counters."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(key[..., 0], dtype=_U32)
        k1 = np.asarray(key[..., 1], dtype=_U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = (np.asarray(x0, dtype=_U32) + ks[0]).astype(_U32)
        x1 = (np.asarray(x1, dtype=_U32) + ks[1]).astype(_U32)
        for s in range(5):
            for r in _ROT[s % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(s + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(s + 2) % 3] + _U32(s + 1)).astype(_U32)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` for a non-negative 32/64-bit seed: (hi, lo) words."""
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key: np.ndarray, n: int = 2) -> np.ndarray:
    """key (..., 2) -> (..., n, 2)."""
    key = np.asarray(key, dtype=_U32)
    idx = np.arange(n, dtype=_U32)
    a, b = threefry2x32(key[..., None, :], np.zeros_like(idx), idx)
    return np.stack([a, b], axis=-1)


def bits(key: np.ndarray, n: int) -> np.ndarray:
    """key (..., 2) -> (..., n) uint32."""
    key = np.asarray(key, dtype=_U32)
    idx = np.arange(n, dtype=_U32)
    a, b = threefry2x32(key[..., None, :], np.zeros_like(idx), idx)
    return a ^ b


def _unit(b: np.ndarray) -> np.ndarray:
    f = ((b >> _U32(9)) | _U32(0x3F800000)).view(np.float32)
    return f - np.float32(1.0)


def uniform(key: np.ndarray, n: int, minval: float | np.ndarray = 0.0, maxval: float | np.ndarray = 1.0) -> np.ndarray:
    """``jax.random.uniform(key, (n,), minval, maxval)`` in float32."""
    lo = np.asarray(minval, dtype=np.float32)
    hi = np.asarray(maxval, dtype=np.float32)
    u = _unit(bits(key, n))
    return np.maximum(lo, u * (hi - lo) + lo).astype(np.float32)


def _erfinv_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float32)
    w = -np.log((np.float32(1) - x) * (np.float32(1) + x)).astype(np.float32)
    small = w < 5
    ws = np.where(small, w - np.float32(2.5), np.sqrt(np.maximum(w, 0)).astype(np.float32) - np.float32(3))
    cs = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941]
    cl = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682]
    p = np.where(small, np.float32(cs[0]), np.float32(cl[0])).astype(np.float32)
    for a, b in zip(cs[1:], cl[1:]):
        p = (np.where(small, np.float32(a), np.float32(b)) + p * ws).astype(np.float32)
    return (p * x).astype(np.float32)


def normal(key: np.ndarray, n: int) -> np.ndarray:
    lo = np.nextafter(np.float32(-1), np.float32(0))
    u = uniform(key, n, lo, 1.0)
    return (np.float32(np.sqrt(2)) * _erfinv_f32(u)).astype(np.float32)


def bernoulli(key: np.ndarray, n: int, p: float) -> np.ndarray:
    return _unit(bits(key, n)) < np.float32(p)
