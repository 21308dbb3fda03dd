"""threefry2x32 known-answer vectors (the Random123 / jax.random test vectors) for all three restatements."""

import ctypes

import numpy as np

from ksim_b200 import rng as R

KAT = [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6B200159, 0x99BA4EFE)),
    ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
    ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0)),
]


def test_numpy_threefry_kat() -> None:
    for key, ctr, want in KAT:
        a, b = R.threefry2x32(np.array(key, dtype=np.uint32), np.uint32(ctr[0]), np.uint32(ctr[1]))
        assert (int(a), int(b)) == want


def test_oracle_threefry_kat(oracle_built: None) -> None:
    import oracle

    lib = oracle.lib("f32")
    for key, ctr, want in KAT:
        o0, o1 = ctypes.c_uint32(), ctypes.c_uint32()
        lib.o_threefry2x32(ctypes.c_uint32(key[0]), ctypes.c_uint32(key[1]), ctypes.c_uint32(ctr[0]), ctypes.c_uint32(ctr[1]),
                           ctypes.byref(o0), ctypes.byref(o1))
        assert (o0.value, o1.value) == want


def test_oracle_matches_numpy_draws(oracle_built: None) -> None:
    import oracle

    lib = oracle.lib("f32")
    key = np.array([123456789, 987654321], dtype=np.uint32)
    n = 64
    out = np.zeros((n, 2), dtype=np.uint32)
    lib.o_rng_split(ctypes.c_void_p(key.ctypes.data), n, ctypes.c_void_p(out.ctypes.data))
    np.testing.assert_array_equal(out, R.split(key, n))
    u = np.zeros(n, dtype=np.float64)
    lib.o_rng_uniform(ctypes.c_void_p(key.ctypes.data), n, ctypes.c_double(-0.25), ctypes.c_double(1.5), ctypes.c_void_p(u.ctypes.data))
    np.testing.assert_array_equal(u.astype(np.float32), R.uniform(key, n, -0.25, 1.5))
    g = np.zeros(n, dtype=np.float64)
    lib.o_rng_normal(ctypes.c_void_p(key.ctypes.data), n, ctypes.c_void_p(g.ctypes.data))
    np.testing.assert_allclose(g.astype(np.float32), R.normal(key, n), rtol=2e-6, atol=1e-7)


def test_uniform_and_normal_statistics() -> None:
    key = R.prng_key(7)
    u = R.uniform(key, 20000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    g = R.normal(key, 20000)
    assert abs(g.mean()) < 0.03 and abs(g.std() - 1.0) < 0.03
    assert R.bernoulli(key, 20000, 0.25).mean() == np.float64((u < np.float32(0.25)).mean())
